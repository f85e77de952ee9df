timeout 400 python scripts/stress_persistent.py 1 2>&1 | tail -22
