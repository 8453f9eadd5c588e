"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel table (markdown)."""
import collections
import csv
import re
import sys


def main(path, out=None, title=""):
    lines = [l for l in open(path) if not l.startswith("==")]
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for row in csv.DictReader(lines):
        name, v = row.get("Kernel Name"), row.get("Metric Value")
        if not name or v is None:
            continue
        try:
            t = float(v.replace(",", ""))
        except ValueError:
            continue
        unit = row.get("Metric Unit", "")
        t *= {"us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "s": 1e9}.get(unit, 1.0)
        name = re.sub(r"\(.*", "", name).replace("void ", "")
        tot[name] += t
        cnt[name] += 1
    s = sum(tot.values())
    import os
    shown = os.path.relpath(path, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))) if os.path.isabs(path) else path
    rows = [f"# {title}", "", f"source: `{shown}` ({sum(cnt.values())} launches, {s / 1e6:.2f} ms summed device time; ncu "
            "serialises launches and runs them cold-cache, so read SHARES, not absolutes)", "",
            "| ms | share | launches | kernel |", "|---:|---:|---:|---|"]
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        if v / s < 0.002:
            continue
        rows.append(f"| {v / 1e6:.3f} | {100 * v / s:.1f}% | {cnt[k]} | `{k[:120]}` |")
    text = "\n".join(rows) + "\n"
    if out:
        open(out, "w").write(text)
    print(text)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None, sys.argv[3] if len(sys.argv) > 3 else "launch list")
