mkdir -p gpurun_out
python -m avr_b200.build --experiments --force > gpurun_out/build_exp.log 2>&1
for r in 1 2; do for d in 0 16 32 64 48 112 1 ; do AVR_CHAIN_DEBUG=$d timeout 100 python profiles/run_chain_once.py 2>&1 | tail -1; done; done
