"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel shares.

    python profiles/summarize_launches.py gpurun_out/launches.csv [step_index] > profiles/rNN_launches_summary.md

With ``step_index`` only the launches of that render step are kept: a step starts at a launch of the ray-generation
encode kernel (``encode_fwd_kernel<true>``, once per step) and ends before the next one.
"""
import collections
import csv
import re
import sys


def short(name: str) -> str:
    name = re.sub(r"<.*", "", name)
    name = name.replace("void ", "")
    return name.split("(")[0][-70:]


def main(path, step=None):
    rows = []
    with open(path) as fh:
        lines = [ln for ln in fh if ln.startswith('"')]
    for rec in csv.DictReader(lines):
        if rec.get("Metric Name") == "gpu__time_duration.sum":
            full = rec["Kernel Name"]
            name = short(full) + (" [raygen]" if re.search(r"encode_(fwd|bwd)_kernel<(\(bool\))?1", full) else "")
            rows.append((name, float(rec["Metric Value"].replace(",", "")) / 1e3))
    if step is not None:
        starts = [i for i, (k, _) in enumerate(rows) if k.startswith("avr::encode_fwd_kernel") and k.endswith("[raygen]")]
        rows = rows[starts[step]:starts[step + 1] if step + 1 < len(starts) else len(rows)]
        print(f"render step {step} of {len(starts)} in the capture")
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
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else None)
