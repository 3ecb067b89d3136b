"""How many ReLU decisions of the tensor-core forward disagree with float64 / with an fp32 GEMM, per layer, and how
large the pre-activation error is relative to the row scale.  Run on the GPU box."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import avr_b200
from avr_b200 import ops
from avr_b200.configs import get_config
from oracle import field_ref
DEV = "cuda:0"
torch.backends.cuda.matmul.allow_tf32 = False
name, n_azi, n_ele = (sys.argv[1] if len(sys.argv) > 1 else "simu"), 16, 8
seeds = [int(s) for s in sys.argv[2:]] or [45]
rec = []
orig = ops.umma_nt
GUARD = os.environ.get("DIAG_GUARD", "0") == "1"
def spy(a, b, flags=0, c=None, c2=None, mask=None, c_f32=None, bits_out=None, **kw):
    if not GUARD:
        kw.pop("guard", None)
    orig(a, b, flags, c, c2, mask, c_f32, bits_out, **kw)
    if bits_out is not None and "bias_rcv" not in kw and "bias_ray" not in kw:
        rec.append((a, b, c, bits_out))
ops.umma_nt = spy
for seed in seeds:
    cfg = get_config(name); cfg["render"]["n_azi"], cfg["render"]["n_ele"] = n_azi, n_ele
    ref_net = field_ref.trained_like_(field_ref.AVRModelRef(cfg["model"], seed=seed), seed=seed + 1)
    r = cfg["render"]; bs = 2
    gen = torch.Generator().manual_seed(11)
    c0 = (r["xyz_min"] + r["xyz_max"]) / 2
    rx = (c0 + (torch.rand(bs, 3, generator=gen) * 2 - 1) * 1.5).float(); tx = (c0 + (torch.rand(bs, 3, generator=gen) * 2 - 1) * 1.5).float()
    azi = torch.rand(n_azi, generator=gen)
    native = avr_b200.AVRModel(cfg["model"]); native.load_state_dict(ref_net.state_dict()); native.to(DEV)
    rec.clear()
    out = avr_b200.AVRRender(native, **r)(rx.to(DEV), tx.to(DEV), azi_rand=azi)
    torch.cuda.synchronize()
    print(f"seed {seed}: {len(rec)} masked forward layers")
    for k, (a, b, c, bits) in enumerate(rec):
        a32, w32 = ops.planes_merge(a), ops.planes_merge(b)
        p64 = a32.double() @ w32.double().t()
        p32 = a32 @ w32.t()
        N = b.rows
        sh = torch.arange(32, device=DEV, dtype=torch.int32)
        tcb = ((bits[:, :(N + 31) // 32].unsqueeze(-1) >> sh) & 1).reshape(bits.shape[0], -1)[:, :N].bool()
        ref = p64 > 0
        scale = p64.abs().mean(1, keepdim=True)
        y = ops.planes_merge(c).double()
        # where TC's output is positive it carries the pre-activation
        pos = (y > 0) & ref
        e_tc = (((y - p64).abs() / scale)[pos]).mean().item()
        e_32 = (((p32.double() - p64).abs() / scale)[pos]).mean().item()
        bias_tc = (((y - p64) / scale)[pos]).mean().item()
        near = lambda t: int(((p64.abs() / scale) < t).sum())
        print(f"  L{k} K={a.cols:3d} N={N:3d}: flips tc {int((tcb != ref).sum()):5d}  fp32 {int(((p32 > 0) != ref).sum()):4d}  of {ref.numel():.2e};"
              f" err/scale tc {e_tc:.2e} (bias {bias_tc:+.2e}) fp32 {e_32:.2e}; |pre|/scale<1e-5: {near(1e-5)}, <1e-4: {near(1e-4)}")
