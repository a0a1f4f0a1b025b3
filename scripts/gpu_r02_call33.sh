timeout 900 python -m pytest tests -m gpu -x -q -k "stream or csv or limit or coalesc" 2>&1 | tail -4
python scripts/csv_probe.py 2>&1 | grep -v "parser alone\|host\] csv" | tail -10
