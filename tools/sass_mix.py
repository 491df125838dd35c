"""Instruction mix + top stall lines from `ncu -i X.ncu-rep --page source --csv` output."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
# a file may hold several kernels: split on "Kernel Name" rows
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1][:100], "rows": []}
        blocks.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
for b in blocks:
    hdr = b["rows"][0]
    si, ix, ss = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    ops, tot = collections.Counter(), 0
    lines = []
    for r in b["rows"][1:]:
        try:
            n = int(r[ix])
        except Exception:
            continue
        t = r[si].split()
        if not t:
            continue
        op = t[1] if t[0].startswith("@") else t[0]
        ops[op.split(".")[0]] += n
        tot += n
        lines.append((int(r[ss] or 0), n, r[si].strip()))
    print("==", b["name"], "total warp-inst", tot)
    print("  " + "  ".join(f"{k}:{100 * v / tot:.1f}%" for k, v in ops.most_common(14)))
    for s, n, src in sorted(lines, reverse=True)[:8]:
        print(f"   samples {s:6d} exec {n:10d}  {src[:90]}")
