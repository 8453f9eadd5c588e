"""Top stall-sample SASS lines per kernel from `ncu --page source --csv` output (gzip ok).
usage: python tools/ncu_source_top.py <file.csv[.gz]> <kernel substring> [occurrence=0] [top=25] [context=2]"""
import csv
import gzip
import sys

path, pat = sys.argv[1], sys.argv[2]
occ = int(sys.argv[3]) if len(sys.argv) > 3 else 0
top = int(sys.argv[4]) if len(sys.argv) > 4 else 25
ctx = int(sys.argv[5]) if len(sys.argv) > 5 else 2
f = gzip.open(path, "rt") if path.endswith(".gz") else open(path)
kernels, cur = [], None
for row in csv.reader(f):
    if row and row[0] == "Kernel Name":
        cur = {"name": row[1], "hdr": None, "rows": []}
        kernels.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = row
    elif cur is not None and row:
        cur["rows"].append(row)
sel = [k for k in kernels if pat in k["name"]]
print(f"{len(kernels)} kernels, {len(sel)} match '{pat}'")
k = sel[occ]
h = {n: i for i, n in enumerate(k["hdr"])}
rows = k["rows"]
tot = sum(int(r[h["# Samples"]]) for r in rows)
print(k["name"][:150], "total samples", tot)
stall_cols = [n for n in k["hdr"] if n.startswith("stall_") and "Not Issued" not in n]
order = sorted(range(len(rows)), key=lambda i: -int(rows[i][h["# Samples"]]))[:top]
for i in sorted(order):
    r = rows[i]
    s = int(r[h["# Samples"]])
    st = sorted(((int(r[h[c]]), c[6:]) for c in stall_cols if int(r[h[c]]) > 0), reverse=True)[:3]
    print(f"--- line {i}: {s} samples ({100.0 * s / tot:.1f}%) exec {r[h['Instructions Executed']]}  {st}")
    for j in range(max(0, i - ctx), min(len(rows), i + ctx + 1)):
        print(("  >> " if j == i else "     ") + rows[j][h["Source"]].strip()[:110])
