"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump by CUDA source line: samples + instructions.
usage: ncu -i rep --page source --csv --print-source cuda,sass | python tools/ncu_lines.py [topN]"""
import collections, csv, sys
top = int(sys.argv[1]) if len(sys.argv) > 1 else 25
rows = list(csv.reader(sys.stdin))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
hdr = rows[hdr_i]
ci = {h: i for i, h in enumerate(hdr)}
# first 'Source' = cuda line text, second = SASS
src_cols = [i for i, h in enumerate(hdr) if h == "Source"]
samp = collections.Counter(); inst = collections.Counter(); text = {}
cur_file = ""
for r in rows:
    if r and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if len(r) < len(hdr) or r[0] == "Line No":
        continue
    try:
        ln = cur_file + ":" + r[ci["Line No"]]
        s = float(r[ci["# Samples"]] or 0); n = float(r[ci["Instructions Executed"]] or 0)
    except ValueError:
        continue
    samp[ln] += s; inst[ln] += n
    text.setdefault(ln, r[src_cols[0]].strip()[:110])
tot_s = sum(samp.values()) or 1; tot_i = sum(inst.values()) or 1
print(f"total samples {tot_s:.0f}, warp instructions {tot_i:.0f}")
for ln, s in samp.most_common(top):
    print(f"{ln:>22} {100*s/tot_s:5.1f}% smp {100*inst[ln]/tot_i:5.1f}% inst  {text[ln]}")
