set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_umma.py tests/test_gpu_render.py tests/test_gpu_kernels.py tests/test_gpu_chain.py -m gpu -q 2>&1 | tail -30 ) > gpurun_out/pytest_subset.log 2>&1
timeout 300 python bench.py --steps 20 --no-other-configs --no-cpu-baseline > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err
timeout 300 python bench.py --config meshrir --mode infer --bs 8 --steps 10 --no-other-configs --no-cpu-baseline > gpurun_out/bench_meshrir_infer.json 2> gpurun_out/bench_meshrir_infer.err
timeout 300 python -c "
import torch, time, avr_b200
from avr_b200.configs import get_config
cfg=get_config('simu'); dev='cuda:0'
f=avr_b200.AVRModel(cfg['model']).to(dev); ren=avr_b200.AVRRender(f, **cfg['render'])
rx=torch.rand(1,3,device=dev); tx=torch.rand(1,3,device=dev)
def t(fn,n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize(); t0=time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter()-t0)/n*1e3
def eager():
    with torch.no_grad(): return ren(rx,tx)
g=ren.graphed_inference(1)
print('simu bs=1 inference: eager %.3f ms, CUDA graph %.3f ms per receiver'%(t(eager), t(lambda: g(rx,tx))))
g8=ren.graphed_inference(8); rx8=torch.rand(8,3,device=dev); tx8=torch.rand(8,3,device=dev)
def eager8():
    with torch.no_grad(): return ren(rx8,tx8)
print('simu bs=8 inference: eager %.3f ms, CUDA graph %.3f ms per call'%(t(eager8,20), t(lambda: g8(rx8,tx8),20)))
" > gpurun_out/graph_latency.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1
cat gpurun_out/graph_latency.txt gpurun_out/smoke.txt; tail -5 gpurun_out/pytest_subset.log
