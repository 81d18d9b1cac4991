#!/usr/bin/env python
"""Join an ncu SASS source page (csv) with nvdisasm -g line info: executed warp instructions / thread
instructions / stall samples per source line.  usage: ncu_lines.py <sass.csv> <kernel.dis> [top]"""
import csv, re, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
ci = {h: i for i, h in enumerate(hdr)}
lines = []; cur = ("?", 0)
for ln in open(sys.argv[2]):
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    if re.match(r"\s*/\*[0-9a-f]{4,}\*/", ln): lines.append(cur)
print("sass rows", len(data), "dis instrs", len(lines), file=sys.stderr)
agg = collections.defaultdict(lambda: [0, 0, 0, 0])
for r, l in zip(data, lines):
    a = agg[l]
    a[0] += int(r[ci["Instructions Executed"]]); a[1] += int(r[ci["Thread Instructions Executed"]]); a[2] += int(r[ci["# Samples"]]); a[3] += 1
tot = [sum(a[k] for a in agg.values()) for k in range(3)]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
print(f"total warp-inst {tot[0]:.3e} thread-inst {tot[1]:.3e} samples {tot[2]}")
for l, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{l[0]:22s}:{l[1]:<5d} sass={a[3]:4d} warp-inst {100*a[0]/tot[0]:5.1f}%  thr/inst {a[1]/max(1,a[0]):5.1f}  samples {100*a[2]/tot[2]:5.1f}%")

if len(sys.argv) > 4:   # category roll-up: file:lo-hi=name,...
    cats = []
    for spec in sys.argv[4].split(","):
        rng, name = spec.split("="); f, lh = rng.split(":"); lo, hi2 = lh.split("-")
        cats.append((f, int(lo), int(hi2), name))
    roll = collections.defaultdict(lambda: [0, 0, 0])
    for l, a in agg.items():
        nm = "other"
        for f, lo, hi2, name in cats:
            if l[0].startswith(f) and lo <= l[1] <= hi2: nm = name; break
        for k in range(3): roll[nm][k] += a[k]
    print("---- categories")
    for nm, a in sorted(roll.items(), key=lambda kv: -kv[1][0]):
        print(f"{nm:28s} warp-inst {100*a[0]/tot[0]:5.1f}%  thr/inst {a[1]/max(1,a[0]):5.1f}  thread-inst {100*a[1]/tot[1]:5.1f}%  samples {100*a[2]/tot[2]:5.1f}%")
