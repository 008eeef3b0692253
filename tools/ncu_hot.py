"""Hot SASS lines of one kernel from `ncu --page source --csv` output: python tools/ncu_hot.py file.csv [section] [min_pct]."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
want = int(sys.argv[2]) if len(sys.argv) > 2 else 0
thr = float(sys.argv[3]) if len(sys.argv) > 3 else 0.5
sections, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        sections.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
sec = sections[want]
h = sec["rows"][0]
data = [r for r in sec["rows"][1:] if len(r) == len(h)]
isrc, ie, iss = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
tot = sum(int(r[ie]) for r in data)
tots = sum(int(r[iss]) for r in data)
print(sec["name"], "total instr", tot, "samples", tots, "sass lines", len(data))
for n, r in enumerate(data):
    e, s = int(r[ie]), int(r[iss])
    if e > thr / 100 * tot or s > thr / 100 * tots:
        print(f"{n:5d} {r[isrc].strip()[:72]:72s} {e:10d} {100*e/tot:5.1f}% {s:7d} {100*s/max(tots,1):5.1f}%")
