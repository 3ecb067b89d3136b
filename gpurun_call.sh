set -x
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_umma.py -m gpu -q -k "cta_pairs" ) > gpurun_out/pytest_pairs.log 2>&1
tail -15 gpurun_out/pytest_pairs.log
