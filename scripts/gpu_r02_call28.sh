nproc; taskset -p $$; cat /sys/fs/cgroup/cpu.max 2>/dev/null; cat /sys/fs/cgroup/cpu.stat 2>/dev/null | head -6
python - <<'PY'
import time, multiprocessing as mp
def burn(_):
    t0=time.perf_counter(); x=0
    for i in range(20_000_000): x+=i
    return time.perf_counter()-t0
for n in (1,2,4,8,16):
    t0=time.perf_counter()
    with mp.Pool(n) as p: r=p.map(burn, range(n))
    print(n, "procs: wall %.2f s, per-proc %.2f-%.2f s" % (time.perf_counter()-t0, min(r), max(r)))
PY
