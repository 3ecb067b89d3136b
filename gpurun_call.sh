mkdir -p gpurun_out
timeout 300 python profiles/ab_tn_from_f16.py 2>&1 | grep -v Warn | tail -6 | tee gpurun_out/ab_tn_from_f16.txt
AVR_BENCH_DETAIL=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-other-configs --no-cpu-baseline --no-alt > gpurun_out/bench_detail.json 2> gpurun_out/bench_detail.err
grep umma_gemm gpurun_out/bench_detail.err | head -12
