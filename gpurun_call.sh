set -x
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q -s --durations=10 ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
AVR_BENCH_DETAIL=1 python bench.py --steps 5 --warmup 3 --no-other-configs --no-cpu-baseline --no-alt > gpurun_out/bench_detail.json 2> gpurun_out/bench_detail.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2b_launches.csv python bench.py --steps 2 --warmup 3 --no-other-configs --no-cpu-baseline --no-alt > gpurun_out/ncu_launches.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --clock-control none -k regex:umma_gemm_kernel -s 57 -c 19 --csv --log-file gpurun_out/r2b_umma_dram_traffic.csv python bench.py --steps 2 --warmup 3 --no-other-configs --no-cpu-baseline --no-alt > gpurun_out/ncu_traffic.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:prefix_walk -c 1 -s 3 -o gpurun_out/prefix_r2b python bench.py --steps 2 --warmup 3 --no-other-configs --no-cpu-baseline --no-alt > gpurun_out/ncu_prefix.log 2>&1
tail -5 gpurun_out/pytest_gpu.log
