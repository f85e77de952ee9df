mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "overflow or power_law or general" ) > gpurun_out/r2_pytest_ovf.log 2>&1; tail -3 gpurun_out/r2_pytest_ovf.log
timeout 600 python scripts/rmat_variants.py --scale 24 --iters 30 --variants default,slots3,slots4,hubs4k --no-cusparse > gpurun_out/r2_rmat24_v7_vote.log 2>&1; tail -5 gpurun_out/r2_rmat24_v7_vote.log
