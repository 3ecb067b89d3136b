set -x
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_chain.py -m gpu -q -x 2>&1 | tail -4 ) > gpurun_out/pytest_chain.log 2>&1
timeout 120 python profiles/run_chain_once.py > gpurun_out/chain_timing.txt 2>&1
timeout 300 python bench.py --steps 20 --no-other-configs --no-cpu-baseline > gpurun_out/bench_chain.json 2> gpurun_out/bench_chain.err
cp avr_b200/libavr_b200.so /tmp/lib_backup.so
python -m avr_b200.build --experiments > gpurun_out/build_exp.log 2>&1
python profiles/trace_chain.py > gpurun_out/chain_trace.txt 2>&1
cp /tmp/lib_backup.so avr_b200/libavr_b200.so
cat gpurun_out/chain_timing.txt; cat gpurun_out/pytest_chain.log
