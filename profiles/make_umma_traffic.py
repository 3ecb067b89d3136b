"""ncu CSV (dram__bytes_read.sum, dram__bytes_write.sum[, lts__t_bytes.sum] of the GEMM launches of one bench.py step)
-> profiles/umma_traffic.json, the `roofline.traffic` figure bench.py quotes.
    python profiles/make_umma_traffic.py <csv> "<source description>" > profiles/umma_traffic.json"""
import csv, json, sys
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
lines = [ln for ln in open(sys.argv[1]) if ln.startswith('"')]
tot, ids = {}, set()
for rec in csv.DictReader(lines):
    m = rec.get("Metric Name")
    if m in ("dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum"):
        tot[m] = tot.get(m, 0.0) + float(rec["Metric Value"].replace(",", "")) * SCALE.get(rec["Metric Unit"], 1.0)
        ids.add(rec["ID"])
n = len(ids)
out = {"kernel": "umma_gemm", "launches": n, "dram_bytes_read": tot.get("dram__bytes_read.sum"),
       "dram_bytes_write": tot.get("dram__bytes_write.sum"),
       "dram_bytes_per_launch": (tot.get("dram__bytes_read.sum", 0) + tot.get("dram__bytes_write.sum", 0)) / max(1, n),
       "l2_bytes": tot.get("lts__t_bytes.sum"), "source": sys.argv[2]}
print(json.dumps(out, indent=1))
