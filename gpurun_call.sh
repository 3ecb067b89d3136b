mkdir -p gpurun_out
timeout 300 python profiles/ab_chain_realexp.py 2>&1 | grep -v Warn | tail -6
