mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:umma_gemm_kernel -s 57 -c 19 -o gpurun_out/umma_step_r2g -f python bench.py --steps 2 --warmup 3 --no-other-configs --no-cpu-baseline --no-alt > gpurun_out/ncu_umma_step.log 2>&1
tail -3 gpurun_out/ncu_umma_step.log
ls -la gpurun_out/umma_step_r2g.ncu-rep
