"""Condense `ncu --set full` reports into the few numbers DESIGN.md / bench.py quote.

    ncu -i report.ncu-rep --page raw --csv > raw.csv
    python profiles/summarize_ncu_full.py raw.csv [label ...] > profiles/rNN/<name>_full_summary.md

One row per captured launch.  DRAM bytes are dram__bytes_read.sum + dram__bytes_write.sum (the `roofline.traffic`
figure of bench.py); "tensor pipe" is sm__pipe_tensor_cycles_active (share of cycles with a tcgen05.mma in flight).
"""
import csv
import sys

SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12,
         "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "nsecond": 1e-9, "usecond": 1e-6, "msecond": 1e-3, "second": 1.0}

COLS = [
    ("ms", "gpu__time_duration.sum", 1e3, "{:.3f}"),
    ("DRAM rd MB", "dram__bytes_read.sum", 1e-6, "{:.0f}"),
    ("DRAM wr MB", "dram__bytes_write.sum", 1e-6, "{:.0f}"),
    ("L2 GB", "lts__t_bytes.sum", 1e-9, "{:.2f}"),
    ("tensor pipe %", "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", 1, "{:.1f}"),
    ("issue act %", "sm__issue_active.avg.pct_of_peak_sustained_elapsed", 1, "{:.1f}"),
    ("L2 thru %", "lts__throughput.avg.pct_of_peak_sustained_elapsed", 1, "{:.1f}"),
    ("DRAM thru %", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 1, "{:.1f}"),
    ("warps/SM", "sm__warps_active.avg.per_cycle_active", 1, "{:.1f}"),
    ("regs", "launch__registers_per_thread", 1, "{:.0f}"),
    ("smem KB", "launch__shared_mem_per_block", 1e-3, "{:.1f}"),
]


def main(path, labels):
    rows = list(csv.reader(open(path)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    ci = {k: i for i, k in enumerate(hdr)}
    print("| # | kernel | grid | " + " | ".join(c[0] for c in COLS) + " | note |")
    print("|---|---|---|" + "---:|" * len(COLS) + "---|")
    for n, r in enumerate(body):
        cells = []
        for _, key, mul, fmt in COLS:
            if key not in ci or r[ci[key]] == "":
                cells.append("-")
                continue
            try:
                v = float(r[ci[key]].replace(",", "")) * SCALE.get(units[ci[key]], 1.0)
            except ValueError:                                  # "no data" for a metric of this launch
                cells.append("-")
                continue
            if units[ci[key]] in ("Kbyte", "Kbyte/block"):
                v = float(r[ci[key]].replace(",", "")) * 1e3
            cells.append(fmt.format(v * mul))
        name = r[ci["Kernel Name"]].split("(")[0].replace("void ", "")[:40]
        note = labels[n] if n < len(labels) else ""
        print(f"| {n} | `{name}` | {r[ci['Grid Size']]} | " + " | ".join(cells) + f" | {note} |")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2:])
