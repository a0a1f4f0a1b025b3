"""Summarise an .ncu-rep: headline metrics per launch + the most-stalled SASS instructions (source page)."""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 12
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"]
for r in rows[2:]:
    print("----", r[hdr.index("Kernel Name")][:70], "id", r[hdr.index("ID")])
    for k in keys:
        if k in hdr:
            print(f"  {k:88s} {r[hdr.index(k)]} {units[hdr.index(k)]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
i, k = 0, 0
while i < len(rows):
    if rows[i] and rows[i][0] == "Kernel Name":
        k += 1
        hdr = rows[i + 1]
        ia, isrc, isamp, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
        cols = {n: hdr.index(n) for n in ["stall_barrier", "stall_long_sb", "stall_branch_resolving", "stall_wait", "stall_short_sb", "stall_lg", "stall_mio", "stall_no_inst", "stall_math", "stall_not_selected"] if n in hdr}
        j, data = i + 2, []
        while j < len(rows) and not (rows[j] and rows[j][0] == "Kernel Name"):
            r = rows[j]
            if len(r) > isamp and r[isamp] != "":
                try:
                    data.append((int(r[isamp]), r[ia][-5:], r[isrc], int(r[iex] or 0), {n: r[c] for n, c in cols.items()}))
                except ValueError:
                    pass
            j += 1
        tot = max(sum(d[0] for d in data), 1)
        if k % 2 == 1 or True:
            print(f"=== launch {k}: samples {tot}, SASS instructions {len(data)}")
            for d in sorted(data, key=lambda x: -x[0])[:top]:
                st = {n[6:]: v for n, v in d[4].items() if v not in ("0", "")}
                print(f"{100*d[0]/tot:5.1f}% ex={d[3]:9d} {d[1]} {d[2][:62]:62s} {st}")
        i = j
    else:
        i += 1
