#!/usr/bin/env python
"""Join an ncu SASS source page (--page source --csv) with nvdisasm line info so stall samples and
executed instructions can be read per CUDA source line.  Usage:
  ncu_lines.py <report.ncu-rep> <cubin> <mangled kernel name> [top N]"""
import csv, re, subprocess, sys, collections, io
rep, cubin, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(dis) if l.strip() == f".text.{kname}:")
addr2line, cur = {}, None
for l in dis[start + 1:]:
    if l.startswith("//---") or l.strip().startswith(".section"):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        # inlined-at chain: keep innermost
        continue
    m = re.match(r"\s*/\*([0-9a-f]+)\*/", l)
    if m and cur:
        addr2line[int(m.group(1), 16)] = cur
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hi]
ia, isamp, iex = h.index("Address"), h.index("# Samples"), h.index("Instructions Executed")
stallcols = [(i, n) for i, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n]
base = None
agg = collections.defaultdict(lambda: [0, 0, collections.Counter()])
tot_s = tot_e = 0
for r in rows[hi + 1:]:
    if len(r) < len(h): continue
    a = int(r[ia], 16)
    if base is None: base = a
    key = addr2line.get(a - base, ("?", 0))
    s, e = int(r[isamp] or 0), int(r[iex] or 0)
    agg[key][0] += s; agg[key][1] += e; tot_s += s; tot_e += e
    for i, n in stallcols:
        v = int(r[i] or 0)
        if v: agg[key][2][n] += v
print(f"total samples {tot_s}, warp instructions {tot_e}")
for key, (s, e, st) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    tops = ", ".join(f"{n[6:]}={v}" for n, v in st.most_common(3))
    print(f"{key[0]}:{key[1]:<5d} samples {100*s/tot_s:5.1f}%  inst {100*e/max(tot_e,1):5.1f}%  {tops}")
# ---- by coarse region (file + line bucket) ----
import os
regions = collections.defaultdict(lambda: [0, 0])
for key, (s, e, st) in agg.items():
    f, ln = key
    regions[(f, ln // 10 * 10)][0] += s; regions[(f, ln // 10 * 10)][1] += e
print("---- by file / 10-line bucket (top 30 by instructions)")
for key, (s, e) in sorted(regions.items(), key=lambda kv: -kv[1][1])[:30]:
    print(f"{key[0]}:{key[1]:<5d} samples {100*s/tot_s:5.1f}%  inst {100*e/max(tot_e,1):5.1f}%")
byfile = collections.defaultdict(lambda: [0, 0])
for key, (s, e, st) in agg.items():
    byfile[key[0]][0] += s; byfile[key[0]][1] += e
for f, (s, e) in byfile.items(): print(f"FILE {f}: samples {100*s/tot_s:.1f}% inst {100*e/tot_e:.1f}%")
tot = collections.Counter()
for key, (s, e, st) in agg.items(): tot.update(st)
print("---- stall totals:", ", ".join(f"{n[6:]}={100*v/tot_s:.1f}%" for n, v in tot.most_common(12)))
