mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_2gpu_h.json 2> gpurun_out/bench_2gpu_h.err
tail -3 gpurun_out/bench_2gpu_h.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_ref_2gpu.json 2> gpurun_out/bench_ref_2gpu.err; tail -2 gpurun_out/bench_ref_2gpu.err; head -c 300 gpurun_out/bench_ref_2gpu.json
python -c "
import json
d=json.load(open('gpurun_out/bench_2gpu_h.json')); print(d['n_gpus'], round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d['grad_checksums'], d['warmup_note'][:50])
for k,v in d['other_configs'].items(): print(k, round(v['value'],1))
"
