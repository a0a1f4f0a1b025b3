"""Generates tests/golden/synth_checksums.json: survivor counts + order-sensitive per-column checksums of the full-size synthetic
queries (BASELINE.json configs[1] at 10^9 rows, configs[4] at 4 x 10^9 rows), computed by the CPU oracle (oracle/: the restatement of
the reference's eager predicate, `eval_cmp` over AnyValues, physical_plan/plan.rs:114-120) straight from the counter-based generator
(include/rivulus_synth.h).  The oracle takes ~25 s per 10^9 rows on 8 cores, too slow to repeat inside every GPU test or bench run, so
the expected values are computed once here and committed; tests/test_synth_checksums.py and bench.py compare the CUDA results with them.

    python scripts/gen_golden_checksums.py            # (re)writes tests/golden/synth_checksums.json
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from rivulus_b200 import capi  # noqa: E402

C2 = [(capi.SYNTH_I64, 1), (capi.SYNTH_F64, 2), (capi.SYNTH_I64, 3), (capi.SYNTH_F64, 4)]   # configs[1]: a, b, c, d
C5 = [(capi.SYNTH_KEY1000, 0), (capi.SYNTH_F64, 1), (capi.SYNTH_BOOL, 2)]                   # configs[4]: k, v, f
CASES = ([("configs[1]", 1_000_000_000, 0, thr, C2) for thr in (998, 899, 499, 99)] +
         [("configs[4]", 4_000_000_000, 0, thr, C5) for thr in (899, 499)])


def main():
    out = []
    for name, n, row0, thr, proj in CASES:
        t0 = time.time()
        count, sums = O.synth_filter_checksums(n, row0, capi.SYNTH_KEY1000, 0, ">", thr, proj)
        out.append({"workload": name, "rows": n, "row0": row0, "pred": {"kind": capi.SYNTH_KEY1000, "col_id": 0, "op": ">", "literal": thr},
                    "proj": [list(p) for p in proj], "limit": -1, "count": count, "checksums": [str(s) for s in sums],
                    "oracle_seconds": round(time.time() - t0, 1)})
        print(out[-1], flush=True)
    path = os.path.join(ROOT, "tests", "golden", "synth_checksums.json")
    with open(path, "w") as f:
        json.dump({"generator": "scripts/gen_golden_checksums.py", "seed": 42, "cases": out}, f, indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()
