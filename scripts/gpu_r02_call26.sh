python scripts/csv_probe.py 2>&1 | tail -60
