set -x
mkdir -p gpurun_out
timeout 600 python profiles/diag_realexp_r2.py > gpurun_out/diag_realexp.txt 2>&1
( time timeout 900 python -m pytest tests/test_gpu_render.py tests/test_gpu_kernels.py tests/test_gpu_chain.py -m gpu -q 2>&1 | tail -30 ) > gpurun_out/pytest_render.log 2>&1
timeout 300 python bench.py --steps 20 --no-other-configs --no-cpu-baseline > gpurun_out/bench_chain.json 2> gpurun_out/bench_chain.err
tail -3 gpurun_out/pytest_render.log
