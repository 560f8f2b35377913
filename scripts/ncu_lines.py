#!/usr/bin/env python
"""Attribute ncu warp-stall samples of one kernel to CUDA source lines.

usage: ncu_lines.py <report.ncu-rep> <cubin-disassembly from `nvdisasm -gi -c`> <kernel substring> [top] [outer-file]
Joins the SASS page of the report (per-instruction samples) with nvdisasm's line markers by
instruction offset.  With `outer-file`, samples of inlined code are also summed by the outermost call
site in that file (e.g. k3_szmap.cu), which separates the phases that share inlined FFT code.
"""
import csv, re, subprocess, sys, collections

rep, sass, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
outer_file = sys.argv[5] if len(sys.argv) > 5 else None
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
chain_open = False
line_of = {}
outer_of = {}
cur = None
chain = []
inside = False
for ln in open(sass):
    if ln.startswith("//-") and ".text." in ln:
        inside = kern in ln
        cur = None
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', ln)
    if m:
        f = m.group(1).split("/")[-1]
        if not chain_open:
            chain = []
            chain_open = True
        chain.append((f, int(m.group(2))))
        if m.group(3):
            chain.append((m.group(3).split("/")[-1], int(m.group(4))))
        continue
    m = re.match(r"\s*/\*([0-9a-f]+)\*/", ln)
    if m and chain:
        chain_open = False
        off = int(m.group(1), 16)
        line_of[off] = chain[0] + (False,)
        if outer_file:
            outs = [c for c in chain if c[0] == outer_file]
            outer_of[off] = outs[-1] if outs else chain[-1]
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

if outer_file:
    agg2 = collections.defaultdict(lambda: [0, 0])
    for addr, src, smp, n, st in inst:
        key = outer_of.get(addr - base, ("?", 0))
        agg2[key][0] += smp; agg2[key][1] += n
    print(f"\n# by outermost line in {outer_file}")
    for (f, l), (smp, n) in sorted(agg2.items(), key=lambda kv: kv[0][1]):
        if smp * 200 >= tot:
            print(f"{100*smp/tot:6.2f}%  inst={n:>11}  {f}:{l}")
