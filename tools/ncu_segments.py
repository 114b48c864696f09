#!/usr/bin/env python
"""Headline counters + hot straight-line SASS segments (runs of instructions with the same execution count) of one ncu
report.  Usage: ncu_segments.py <report.ncu-rep> [min share %]"""
import csv, io, re, subprocess, sys
rep = sys.argv[1]
minshare = float(sys.argv[2]) if len(sys.argv) > 2 else 0.4
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, u, v = rows[0], rows[1], rows[2]
keep = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_eligible.avg.per_cycle_active", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum", "smsp__thread_inst_executed.sum", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active")
for n, uu, vv in zip(h, u, v):
    if n in keep:
        print(f"{n},{uu},{vv}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hi]
ia, isrc, iex, ith, isamp = h.index("Address"), h.index("Source"), h.index("Instructions Executed"), h.index("Thread Instructions Executed"), h.index("# Samples")
data = [(int(r[ia], 16), r[isrc].strip(), int(r[iex] or 0), int(r[ith] or 0), int(r[isamp] or 0)) for r in rows[hi + 1:] if len(r) >= len(h)]
base = data[0][0]
tot = sum(d[2] for d in data); tots = sum(d[4] for d in data)
print(f"sass instructions {len(data)}, executed {tot}, samples {tots}")
seg = []
for i, (a, s, e, t, sm) in enumerate(data):
    if seg and seg[-1][2] == e:
        seg[-1][1] = i
    else:
        seg.append([i, i, e])
for i0, i1, e in seg:
    n = i1 - i0 + 1
    if e * n < tot * minshare / 100:
        continue
    ops = " ".join((d[1].split()[1] if d[1].startswith("@") else d[1].split()[0]) for d in data[i0:i1 + 1])[:130]
    thr = sum(d[3] for d in data[i0:i1 + 1]) / max(1, e * n)
    smp = sum(d[4] for d in data[i0:i1 + 1])
    print(f"+{data[i0][0] - base:6x} n={n:4d} exec={e:8d} inst={100 * e * n / tot:5.2f}% samples={100 * smp / max(tots, 1):5.2f}% thr={thr:4.1f} {ops}")
