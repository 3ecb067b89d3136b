set -x
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2f_launches.csv python bench.py --steps 2 --warmup 3 --no-other-configs --no-cpu-baseline --no-alt > gpurun_out/ncu_launches.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --clock-control none -k regex:umma_gemm_kernel -s 57 -c 19 --csv --log-file gpurun_out/r2f_umma_dram_traffic.csv python bench.py --steps 2 --warmup 3 --no-other-configs --no-cpu-baseline --no-alt > gpurun_out/ncu_traffic.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mlp_chain -c 1 -s 3 -o gpurun_out/chain_r2f -f python bench.py --steps 2 --warmup 3 --no-other-configs --no-cpu-baseline --no-alt > gpurun_out/ncu_chain.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1; tail -2 gpurun_out/smoke.txt
head -c 600 gpurun_out/bench_1gpu.json
