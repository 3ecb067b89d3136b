# What the round's records were made with (one B200): gpurun --timeout 2400 -- 'bash gpurun_call.sh'
set -x
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q --durations=8 ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-other-configs --no-cpu-baseline --no-alt > gpurun_out/ncu_launches.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --clock-control none -k regex:umma_gemm_kernel -s 57 -c 19 --csv --log-file gpurun_out/umma_dram_traffic.csv python bench.py --steps 2 --warmup 3 --no-other-configs --no-cpu-baseline --no-alt > gpurun_out/ncu_traffic.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:umma_gemm_kernel -s 57 -c 19 -o gpurun_out/umma_step -f python bench.py --steps 2 --warmup 3 --no-other-configs --no-cpu-baseline --no-alt > gpurun_out/ncu_umma_step.log 2>&1
tail -5 gpurun_out/pytest_gpu.log
