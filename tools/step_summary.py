"""Per-kernel summary of ONE training step (the launches between two consecutive pairs of Adam launches) from an
ncu launch list; optional per-launch dump of a range.  usage: step_summary.py launches.csv [lo hi]"""
import collections, csv, re, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
R = []
for r in csv.DictReader(lines):
    n = re.sub(r"\(.*", "", r['Kernel Name']).replace('void ', '').replace('mpgan::', '').replace('__nv_bfloat16', 'bf')
    R.append((float(r['Metric Value']) / 1e3, n, r['Grid Size'], int(r['ID'])))
ad = [i for i, (t, n, g, _) in enumerate(R) if n.startswith('adam_kernel')]
if len(sys.argv) > 3:
    for t, n, g, i in R[int(sys.argv[2]):int(sys.argv[3])]:
        print(f"{i:5d} {t:8.1f} {n[:64]:64s} {g}")
    sys.exit()
if len(ad) >= 3:
    a, b = ad[0], ad[2]
    seg = R[a + 1:b + 1]
else:
    seg = R
print(len(seg), "launches", round(sum(t for t, _, _, _ in seg) / 1e3, 2), 'ms; adam at', ad)
tot = collections.defaultdict(float); cnt = collections.Counter()
for t, n, g, _ in seg:
    tot[n] += t; cnt[n] += 1
for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:45]:
    print(f"{v/1e3:7.3f} {cnt[k]:4d} {k[:80]}")
