set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "chunk_plan or two_pass" > gpurun_out/r02_pytest_chunk.log 2>&1; echo "pytest-chunk rc=$?"; tail -15 gpurun_out/r02_pytest_chunk.log
timeout 900 python -m pytest tests/test_synth_checksums.py -m gpu -x -q > gpurun_out/r02_pytest_chunk2.log 2>&1; echo "pytest-synth rc=$?"; tail -6 gpurun_out/r02_pytest_chunk2.log
timeout 300 python - <<'PY'
import sys
sys.path.insert(0, '.')
from rivulus_b200 import capi
ctx = capi.Context(0)
spec = [(capi.SYNTH_KEY1000, 0, 0), (capi.SYNTH_F64, 1, 0), (capi.SYNTH_BOOL, 2, 0)]
t = ctx.gen_batch(spec, 500_000_000, 3_500_000_000)
ctx.profile_enable(True)
for chunk in (0, 1):
    ctx.set_option(capi.OPT_CHUNK_PLAN, chunk)
    for thr in (998, 899, 799, 499, 99):
        for r in range(3):
            o = ctx.filter_project(t, capi.predicate(0, ">", thr), [0, 1, 2])
            ms = ctx.profile_read_launches()
            n = o.num_rows(); cs = o.checksum(0) if r == 0 else None; o.release()
        print("c5 shard 500M chunk_plan", chunk, thr, n, "device ms", [round(x, 3) for x in ms], flush=True)
PY
for k in 1 3; do python scripts/c3_one.py --kernel $k --reps 3 | tail -2; done
python scripts/profile_one.py --rows 1000000000 --reps 2 | tail -4
