set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 200 --warmup 10 2>gpurun_out/bench_err.log | tee gpurun_out/bench1.json
tail -5 gpurun_out/bench_err.log
EHYB_PDL=0 python scripts/sweep.py --iters 200 --configs 148:0:1:0:-1:2:8 2>&1 | grep "^cfg" | sed "s/^/pdl=0 /"
EHYB_PDL=1 python scripts/sweep.py --iters 200 --configs 148:0:1:0:-1:2:8,148:0:1:0:-1:2:4,296:0:1:0:-1:2:8 2>&1 | grep "^cfg" | sed "s/^/pdl=1 /"
python bench.py --impl reference --steps 20 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_ref1.json
