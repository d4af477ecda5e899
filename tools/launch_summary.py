"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): per-launch times of the LAST `tail` launches
(or of launches [a, b)), short kernel names.  usage: python tools/launch_summary.py file.csv [tail | a:b]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[hi]
ki, vi, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
seq = [(r[ki], float(r[vi].replace(",", "")), r[gi]) for r in rows[hi + 1:] if len(r) > vi]
sel = sys.argv[2] if len(sys.argv) > 2 else "60"
if ":" in sel:
    a, b = sel.split(":"); part = seq[int(a):int(b)]; off = int(a)
else:
    part = seq[-int(sel):]; off = len(seq) - len(part)
tot = 0.0
for i, (k, v, g) in enumerate(part):
    name = k.split("(")[0].replace("void ", "").replace("kzg::", "")
    print(f"{off + i:5d} {name[:48]:48s} {g:>18s} {v / 1e3:9.1f} us")
    tot += v
print(f"total {tot / 1e3:.1f} us over {len(part)} launches")
