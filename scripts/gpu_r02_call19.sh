RVL_TRACE_ALLOC=1 python scripts/outlier_probe.py all 2>&1 | tail -30
echo ---- nosmall
RVL_TRACE_ALLOC=1 python scripts/outlier_probe.py nosmall 2>&1 | tail -30
