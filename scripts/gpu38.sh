set -x
for cfg in "1 1" "0 1" "1 0" "0 0"; do set -- $cfg; echo "LATE_TRIGGER=$1 PDL_OVF=$2"; EHYB_OVF_LATE_TRIGGER=$1 EHYB_PDL_OVF=$2 timeout 300 python scripts/run_rmat.py --scale 20 --blocks --iters 50 2>&1 | grep "us per product"; done
