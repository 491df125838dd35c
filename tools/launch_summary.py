"""Aggregate an ncu launch list (gpu__time_duration.sum, --csv) by kernel name."""
import collections
import csv
import re
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[hi]
kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
for r in rows[hi + 1:]:
    if len(r) <= mv:
        continue
    v = float(r[mv].replace(",", ""))
    if r[mu] == "ns":
        v /= 1e3
    elif r[mu] == "ms":
        v *= 1e3
    name = re.sub(r"\(.*", "", r[kn])[:90]
    agg[name][0] += 1
    agg[name][1] += v
    tot += v
print(f"total {tot:.1f} us over {sum(a[0] for a in agg.values())} launches")
for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:top]:
    print(f"{t:10.1f} us {100 * t / tot:5.1f}% n={c:4d}  {n}")
