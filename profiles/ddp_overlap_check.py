"""torchrun check: gradients reduced inside the backward pass (GradArena.attach) == one all-reduce of the flat arena after
it (deterministic hash-grid mode, so the comparison is bit-exact for 2 ranks), and the step time of both.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 profiles/ddp_overlap_check.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import avr_b200
from avr_b200.configs import get_config
rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
cfg = get_config("simu")
field = avr_b200.AVRModel(cfg["model"], seed=1337)
with torch.no_grad():
    g = torch.Generator().manual_seed(7)
    for m in field.modules():
        if isinstance(m, avr_b200.Encoding):
            m.params.copy_(torch.randn(m.params.shape, generator=g) * 0.1)
field = field.to(dev)
ren = avr_b200.AVRRender(field, **cfg["render"], grid_grad="deterministic")
arena = avr_b200.GradArena(ren.parameters())
gen = torch.Generator().manual_seed(100 + rank)
rx = ((torch.rand(4, 3, generator=gen) * 2 - 1) * 4).to(dev); tx = ((torch.rand(4, 3, generator=gen) * 2 - 1) * 4).to(dev)
azi = torch.rand(cfg["render"]["n_azi"], generator=torch.Generator().manual_seed(3))
def step(overlap):
    arena.zero_()
    out = ren(rx, tx, azi_rand=azi)
    out.square().sum().backward()
    if not overlap:
        arena.all_reduce_mean()
res = {}
for overlap in (False, True):
    if overlap:
        arena.attach(ren)
    for _ in range(3):
        step(overlap)
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        step(overlap)
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / 10], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    res[overlap] = (float(t), arena.flat.clone())
same = torch.equal(res[False][1], res[True][1])
diff = float((res[False][1] - res[True][1]).abs().max() / res[False][1].abs().max())
if rank == 0:
    print(f"world {world}: step after-backward all-reduce {res[False][0]:.3f} ms, overlapped {res[True][0]:.3f} ms; "
          f"gradients identical: {same} (max rel diff {diff:.2e}); nonzero: {bool(res[True][1].abs().max() > 0)}")
dist.destroy_process_group()
