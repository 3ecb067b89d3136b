mkdir -p gpurun_out
timeout 600 python profiles/diag_flips_f16_vs_bf16x3.py 2>&1 | grep -v Warn | tail -8 | tee gpurun_out/diag_flips_f16_vs_bf16x3.txt
