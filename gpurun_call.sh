mkdir -p gpurun_out
python bench.py --steps 40 --warmup 3 --no-other-configs --no-cpu-baseline --no-alt > gpurun_out/solo.json 2>/dev/null
CUDA_VISIBLE_DEVICES=0 python bench.py --steps 40 --warmup 3 --no-other-configs --no-cpu-baseline --no-alt > gpurun_out/conc0.json 2>/dev/null &
CUDA_VISIBLE_DEVICES=1 python bench.py --steps 40 --warmup 3 --no-other-configs --no-cpu-baseline --no-alt > gpurun_out/conc1.json 2>/dev/null
wait
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 40 --warmup 3 --no-other-configs --no-cpu-baseline --no-alt > gpurun_out/ddp2.json 2>/dev/null
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 40 --warmup 3 --no-other-configs --no-cpu-baseline --no-alt --flat-allreduce > gpurun_out/ddp2_flat.json 2>/dev/null
python -c "
import json
for f in ('solo','conc0','conc1','ddp2','ddp2_flat'):
    d=json.load(open('gpurun_out/%s.json'%f)); print(f, round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), d['clocks']['sm_mhz'])
"
