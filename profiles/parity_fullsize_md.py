"""profiles/r2/parity_fullsize.jsonl (written by tests/test_gpu_fullsize.py on the GPU box) -> parity_fullsize.md

    python profiles/parity_fullsize_md.py profiles/r2/parity_fullsize.jsonl "<log the run belongs to>" > profiles/r2/parity_fullsize.md
"""
import json
import sys

rows = [json.loads(l) for l in open(sys.argv[1])]
log = sys.argv[2] if len(sys.argv) > 2 else ""


def worst(g):
    k = max(g, key=g.get)
    return g[k], k.replace(".params", "")


print("# Parity at full size and over all seeds (round 2, final build)\n")
print(f"Source: `tests/test_gpu_fullsize.py` on one B200 ({log}; raw numbers in `parity_fullsize.jsonl`; table made by")
print("`profiles/parity_fullsize_md.py`).  All numbers are relative L2 distances to the CPU oracle (`oracle/`: restatement of")
print("`renderer_cpu.py` + fp32 field).\n")
print("## Every BASELINE config at its real ray count\n")
print("`tc` = the shipped tensor-core path (deterministic table gradients), `tc[atomic]` = fp32-atomic table gradients, `simt` = the")
print("independent fp32 FMA path of this library.  `spread` = how far fp32 arithmetic itself leaves the answer open on that draw")
print("(`tests/helpers.py::oracle_conditioning`: summation orders + every ReLU decision within 1e-6 of zero flipped; measured only")
print("when the 1e-4 bar is exceeded); the bar is `max(1e-4, 2 x spread)`.\n")
print("| config | bs | CPU oracle s | tc IR | tc worst grad (param) | tc[atomic] worst | simt IR | simt worst grad | tc vs simt worst | spread |")
print("|---|---:|---:|---:|---|---:|---:|---:|---:|---:|")
for r in rows:
    if r["test"] != "full_size_vs_cpu_oracle":
        continue
    w, k = worst(r["tc"]["grads"])
    print(f"| {r['config']} | {r['bs']} | {r['cpu_oracle_seconds']:.1f} | {r['tc']['ir']:.1e} | {w:.1e} ({k}) | "
          f"{worst(r['tc_atomic']['grads'])[0]:.1e} | {r['simt']['ir']:.1e} | {worst(r['simt']['grads'])[0]:.1e} | "
          f"{worst(r['tc_vs_simt']['grads'])[0]:.1e} | {('%.1e' % r['oracle_fp32_noise']) if r['oracle_fp32_noise'] else '-'} |")
print("""
Reading: the rendered IR agrees to ~1e-6 everywhere.  The hash-table gradients agree to < 5e-5 on draws where no ReLU
decision lies within fp32 rounding noise of zero, and differ by 1e-4 .. 2e-3 between ANY two fp32 evaluations (oracle, simt,
tc -- whichever two happen to take the same decisions agree to ~1e-5) on draws where one does: see `parity_flips.md`.

## The real fields on reduced ray grids, every candidate seed

A seed is held to the plain 1e-4 bar when `2 x spread <= 1e-4`, else to `2 x spread` (marked "wide").
""")
print("| config | rays | seed | spread of the oracle | bar | tc IR | tc worst grad (param) | simt IR | simt worst grad |")
print("|---|---:|---:|---:|---|---:|---|---:|---:|")
for r in rows:
    if r["test"] != "real_fields_all_seeds":
        continue
    for s in r["seeds"]:
        bar = "1e-4" if s["well_conditioned"] else f"wide ({2 * s['oracle_noise']:.1e})"
        print(f"| {r['config']} | {r['rays']} | {s['seed']} | {s['oracle_noise']:.1e} | {bar} | {s['tc_ir']:.1e} | "
              f"{s['tc_worst_grad']:.1e} ({s['tc_worst_param'].replace('.params', '')}) | {s['simt_ir']:.1e} | {s['simt_worst_grad']:.1e} |")
n = sum(len(r["seeds"]) for r in rows if r["test"] == "real_fields_all_seeds")
held = sum(s["well_conditioned"] for r in rows if r["test"] == "real_fields_all_seeds" for s in r["seeds"])
under = sum(s["tc_worst_grad"] < 1e-4 for r in rows if r["test"] == "real_fields_all_seeds" for s in r["seeds"])
print(f"\n{held} of {n} seeds are held to 1e-4; the tensor-core path is within 1e-4 on {under} of {n} (on every seed whose spread allows it).")
