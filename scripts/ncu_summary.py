#!/usr/bin/env python
"""Summarise an .ncu-rep (read here with `ncu -i`): key raw metrics per kernel + executed-instruction
histogram by opcode.  Usage: python scripts/ncu_summary.py gpurun_out/x.ncu-rep [out.md]"""
import collections, csv, io, subprocess, sys

rep = sys.argv[1]
out = open(sys.argv[2], "w") if len(sys.argv) > 2 else sys.stdout
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
        "smsp__inst_executed.sum", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]
print(f"# ncu summary of {rep}\n", file=out)
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print(f"## {d['Kernel Name']}  grid {d.get('Grid Size')} block {d.get('Block Size')}\n", file=out)
    print("| metric | value | unit |\n|---|---|---|", file=out)
    for k in KEYS:
        if k in d and d[k] != "":
            print(f"| {k} | {d[k]} | {units[hdr.index(k)]} |", file=out)
    print(file=out)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
sections, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}; sections.append(cur); continue
    if cur is not None:
        cur["rows"].append(r)
for sec in sections:
    h = sec["rows"][0]
    iA, iE = h.index("Source"), h.index("Instructions Executed")
    tot, byop = 0, collections.Counter()
    for r in sec["rows"][1:]:
        if len(r) <= iE or not r[iE]:
            continue
        try:
            n = int(r[iE])
        except ValueError:
            continue
        toks = r[iA].split()
        op = toks[1] if toks[0].startswith("@") else toks[0]
        byop[op.split(".")[0]] += n; tot += n
    print(f"## executed warp instructions by opcode: {sec['name']} (total {tot})\n", file=out)
    print("| opcode | warp instructions | share |\n|---|---|---|", file=out)
    for op, n in byop.most_common(24):
        print(f"| {op} | {n} | {100 * n / max(tot, 1):.1f}% |", file=out)
    print(file=out)
