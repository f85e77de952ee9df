set -x
for pdl in 1 0; do for g in "" "--no-graph"; do echo "PDL=$pdl graph=$g"; EHYB_PDL=$pdl timeout 300 python scripts/run_rmat.py --scale 20 --blocks --iters 50 $g 2>&1 | grep "us per product\|main kernel"; done; done
