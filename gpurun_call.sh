set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline --no-other-configs --no-alt > gpurun_out/bench_8box_1gpu.json 2> gpurun_out/bench_8box_1gpu.err
for n in 2 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 3 --no-cpu-baseline --no-other-configs --no-alt > gpurun_out/bench_8box_${n}gpu.json 2> gpurun_out/bench_8box_${n}gpu.err
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_8box_8gpu.json 2> gpurun_out/bench_8box_8gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 20 --warmup 3 --no-cpu-baseline --no-other-configs --no-alt --flat-allreduce > gpurun_out/bench_8box_8gpu_flat.json 2> gpurun_out/bench_8box_8gpu_flat.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 8 --config meshrir --mode infer --receivers 3969 --bs 8 > gpurun_out/bench_meshrir_all_8gpu.json 2> gpurun_out/bench_meshrir_all_8gpu.err
ls -la gpurun_out
