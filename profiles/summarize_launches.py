"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel shares.

    python profiles/summarize_launches.py gpurun_out/launches.csv > profiles/rNN_launches_summary.md
"""
import collections
import csv
import re
import sys


def short(name: str) -> str:
    name = re.sub(r"<.*", "", name)
    name = name.replace("void ", "")
    return name.split("(")[0][-70:]


def main(path):
    rows = []
    with open(path) as fh:
        lines = [ln for ln in fh if ln.startswith('"')]
    for rec in csv.DictReader(lines):
        if rec.get("Metric Name") == "gpu__time_duration.sum":
            rows.append((short(rec["Kernel Name"]), float(rec["Metric Value"].replace(",", "")) / 1e3))
    tot = sum(t for _, t in rows)
    agg = collections.OrderedDict()
    for k, t in rows:
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += t
    print(f"launches: {len(rows)}, total device time {tot / 1e3:.3f} ms (ncu: serialised, cold-cache -> compare shares)\n")
    print("| kernel | launches | total us | share |")
    print("|---|---:|---:|---:|")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {n} | {t:.1f} | {100 * t / tot:.1f}% |")


if __name__ == "__main__":
    main(sys.argv[1])
