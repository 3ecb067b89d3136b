set -x
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q --durations=5 ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
AVR_BENCH_DETAIL=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-other-configs --no-cpu-baseline --no-alt > gpurun_out/bench_detail.json 2> gpurun_out/bench_detail.err
grep "mlp_chain" gpurun_out/bench_detail.err
python -c "import json; d=json.load(open('gpurun_out/bench_detail.json')); print(d['value'], d['ms_per_step'], d['clocks'])"
