mkdir -p gpurun_out
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_1gpu_h.json 2> gpurun_out/bench_1gpu_h.err; tail -2 gpurun_out/bench_1gpu_h.err
python -c "
import json
d=json.load(open('gpurun_out/bench_1gpu_h.json')); print(round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), round(d['e2e']['ms_per_step'],3), 'frac', round(d['roofline']['frac'],4), d['warmup'], d['warmup_note'][:60], d['grid_grad_alt']['value'])
"
