#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed): key metrics + top stall sites.
    python tools/ncu_summary.py gpurun_out/x.ncu-rep [--top 25] [--stream out.txt]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__cycles_elapsed.avg", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
        "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_fmaheavy.sum",
        "sm__inst_executed_pipe_lsu.sum", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
        "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct",
        "smsp__warp_issue_stalled_wait_per_warp_active.pct", "smsp__warp_issue_stalled_not_selected_per_warp_active.pct",
        "smsp__warp_issue_stalled_dispatch_stall_per_warp_active.pct", "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct",
        "smsp__warp_issue_stalled_membar_per_warp_active.pct", "smsp__warp_issue_stalled_sleeping_per_warp_active.pct",
        "smsp__warp_issue_stalled_no_instruction_per_warp_active.pct", "smsp__warp_issue_stalled_branch_resolving_per_warp_active.pct"]
for i, h in enumerate(hdr):
    if h in want:
        print(f"{h} [{units[i]}]: {[r[i] for r in data]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hidx = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
if hidx:
    h = rows[hidx[0]]
    end = hidx[1] - 1 if len(hidx) > 1 else len(rows)
    body = [r for r in rows[hidx[0] + 1:end] if len(r) > 6]
    ci = {n: i for i, n in enumerate(h)}
    stall_cols = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
    tot = sum(int(r[ci["# Samples"]] or 0) for r in body)
    toti = sum(int(r[ci["Instructions Executed"]] or 0) for r in body)
    print(f"-- source: {len(body)} SASS lines, {tot} samples, {toti} warp-instructions")
    agg = {}
    for r in body:
        for n in stall_cols:
            v = int(r[ci[n]] or 0)
            if v:
                agg[n] = agg.get(n, 0) + v
    print("stall totals:", sorted(agg.items(), key=lambda kv: -kv[1])[:10])
    out = [(k, r[1].strip(), int(r[ci["# Samples"]] or 0), int(r[ci["Instructions Executed"]] or 0),
            {n: int(r[ci[n]] or 0) for n in stall_cols if int(r[ci[n]] or 0)}) for k, r in enumerate(body)]
    for k, s, n, ie, st in sorted(out, key=lambda t: -t[2])[:top]:
        print(f"{k:5d} {n:5d} {ie:8d}  {s[:72]:72s} {st}")
    if "--stream" in sys.argv:
        open(sys.argv[sys.argv.index("--stream") + 1], "w").write("\n".join(f"{k:5d} {ie:8d} {n:6d}  {s}" for k, s, n, ie, st in out))
