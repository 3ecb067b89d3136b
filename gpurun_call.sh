mkdir -p gpurun_out
python bench.py --bs 1 --steps 30 --warmup 5 --no-other-configs --no-cpu-baseline --no-alt > gpurun_out/bench_simu_bs1.json 2>/dev/null
python bench.py --bs 8 --steps 10 --warmup 3 --no-other-configs --no-cpu-baseline --no-alt > gpurun_out/bench_simu_bs8.json 2>/dev/null
python bench.py --mode infer --bs 1 --steps 50 --warmup 5 --no-other-configs --no-cpu-baseline --no-alt > gpurun_out/bench_simu_infer_bs1.json 2>/dev/null
python -c "
import json
for f in ('bench_simu_bs1','bench_simu_bs8','bench_simu_infer_bs1'):
    d=json.load(open('gpurun_out/%s.json'%f)); print(f, round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1))
"
