python - <<'PY'
import numpy as np, os
n = 1_000_000
rng = np.random.default_rng(7)
ids = rng.integers(-10 ** 9, 10 ** 9, n); xs = rng.normal(size=n) * 1000.0; sl = rng.integers(0, 100000, n); fl = rng.integers(0, 2, n); nul = rng.random((n, 4)) < 0.1
lines = ["id,x,s,flag"]
for i in range(n):
    lines.append("%s,%s,%s,%s" % ("" if nul[i, 0] else ids[i], "" if nul[i, 1] else repr(float(xs[i])), "null" if nul[i, 2] else "name_%d" % sl[i], "" if nul[i, 3] else ("true" if fl[i] else "false")))
open("/tmp/b.csv", "w").write("\n".join(lines) + "\n")
PY
g++ -O2 -std=c++17 -Irivulus_b200/host -Iinclude scripts/csv_parse_speed.cpp -o /tmp/csv_parse_speed -Lrivulus_b200/lib -lrivulus_host -lrivulus_gpu -Wl,-rpath,$PWD/rivulus_b200/lib -pthread
uname -r
for t in 0 2 4 8 16; do echo "threads $t"; /tmp/csv_parse_speed /tmp/b.csv $t 2>&1 | tail -3; done
python scripts/csv_probe.py 2>&1 | grep -v "parser alone" | tail -40
