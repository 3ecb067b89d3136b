set -x
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_1gpu_g.json 2> gpurun_out/bench_1gpu_g.err
tail -3 gpurun_out/bench_1gpu_g.err
python bench.py --config meshrir --mode infer --receivers 512 --bs 8 > gpurun_out/bench_meshrir_512.json 2> gpurun_out/bench_meshrir_512.err; tail -2 gpurun_out/bench_meshrir_512.err
python -c "
import json
d=json.load(open('gpurun_out/bench_1gpu_g.json')); print(round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), round(d['e2e']['ms_per_step'],3))
for k,v in d['other_configs'].items(): print(k, round(v['value'],1), 'e2e', round(v['e2e']['value'],1))
d=json.load(open('gpurun_out/bench_meshrir_512.json')); print(d['value'], d['e2e'])
"
