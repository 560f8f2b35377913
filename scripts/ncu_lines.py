#!/usr/bin/env python
"""Attribute ncu warp-stall samples of one kernel to CUDA source lines.

usage: ncu_lines.py <report.ncu-rep> <cubin-disassembly from `nvdisasm -g -c`> <kernel substring> [top]
Joins the SASS page of the report (per-instruction samples) with nvdisasm's line markers by
instruction offset.
"""
import csv, re, subprocess, sys, collections

rep, sass, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ci = {n: hdr.index(n) for n in ("Address", "Source", "# Samples", "Instructions Executed")}
stall_cols = [i for i, n in enumerate(hdr) if n.startswith("stall_") and "Not Issued" not in n]
inst = []
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        continue
    inst.append((int(r[0], 16), r[ci["Source"]].strip(), int(r[ci["# Samples"]] or 0), int(r[ci["Instructions Executed"]] or 0),
                 [int(r[i] or 0) for i in stall_cols]))
base = inst[0][0]
# nvdisasm: line markers and instruction offsets inside the kernel's section
line_of = {}
cur = None
inside = False
for ln in open(sass):
    if ln.startswith("//-") and ".text." in ln:
        inside = kern in ln
        cur = None
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
    if m:
        f = m.group(1).split("/")[-1]
        cur = (f, int(m.group(2)), "inlined" in m.group(3))
        continue
    m = re.match(r"\s*/\*([0-9a-f]+)\*/", ln)
    if m and cur:
        line_of[int(m.group(1), 16)] = cur
agg = collections.defaultdict(lambda: [0, 0, collections.Counter()])
tot = 0
for addr, src, s, n, st in inst:
    key = line_of.get(addr - base, ("?", 0, False))[:2]
    a = agg[key]
    a[0] += s; a[1] += n
    for i, v in zip(stall_cols, st):
        a[2][hdr[i]] += v
    tot += s
print(f"total samples {tot}, instructions {sum(i[3] for i in inst)}")
for (f, l), (s, n, c) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    top3 = ", ".join(f"{k[6:]}={v}" for k, v in c.most_common(3))
    print(f"{100*s/tot:6.2f}%  inst={n:>11}  {f}:{l:<5} {top3}")
