set -x
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --durations=8 ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1; tail -1 gpurun_out/smoke.txt
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2g_launches.csv python bench.py --steps 2 --warmup 3 --no-other-configs --no-cpu-baseline --no-alt > gpurun_out/ncu_launches.log 2>&1
python -c "
import json
d=json.load(open('gpurun_out/bench_1gpu.json')); print(round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'frac', round(d['roofline']['frac'],4), round(d['roofline']['tensor_pipe_frac'],3), d['clocks'])
print({k:round(v['ms_per_step'],3) for k,v in d['kernels'].items()}, d['grid_grad_alt']['value'])
for k,v in d['other_configs'].items(): print(k, round(v['value'],1), 'e2e', round(v['e2e']['value'],1))
r=json.load(open('gpurun_out/bench_reference.json')); print('ref', r['value'])
"
