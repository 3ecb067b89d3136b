mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_umma.py tests/test_gpu_chain.py tests/test_gpu_render.py -m gpu -x -q ) > gpurun_out/pytest_umma.log 2>&1
tail -3 gpurun_out/pytest_umma.log
AVR_BENCH_DETAIL=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-other-configs --no-cpu-baseline --no-alt > gpurun_out/bench_detail.json 2> gpurun_out/bench_detail.err
grep umma_gemm gpurun_out/bench_detail.err | sed -n 8,14p
python -c "import json; d=json.load(open('gpurun_out/bench_detail.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['ms_per_step'])"
