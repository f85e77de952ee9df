for s in 2 3 4 5 6 7; do timeout 300 python scripts/stress_persistent.py $s 2>&1 | tail -1; done
