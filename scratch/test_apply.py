import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ot_vae_lightning_b200 import kernels as K, _native as N
lib = N.load()
torch.manual_seed(0)
def run(n, d, cg, L=None, timing=False):
    lib.otkdbg_set_apply_cg(cg)
    shape = (n, d) if L is None else (L, n, d)
    lead = () if L is None else (L,)
    x = torch.randn(*shape, device='cuda') * 2 + 1
    T = torch.randn(*lead, d, d, device='cuda', dtype=torch.float64) / d ** 0.5
    ms = torch.randn(*lead, d, device='cuda', dtype=torch.float64); mt = torch.randn(*lead, d, device='cuda', dtype=torch.float64)
    y = K.apply_transport(x, ms, mt, T); torch.cuda.synchronize()
    nchk = min(n, 4096)
    xs = x[..., :nchk, :].double(); xe = x[..., n - nchk:, :].double()
    ref_s = (xs - ms.unsqueeze(-2)) @ T.transpose(-1, -2) + mt.unsqueeze(-2)
    ref_e = (xe - ms.unsqueeze(-2)) @ T.transpose(-1, -2) + mt.unsqueeze(-2)
    es = ((y[..., :nchk, :].double() - ref_s).norm() / ref_s.norm()).item()
    ee = ((y[..., n - nchk:, :].double() - ref_e).norm() / ref_e.norm()).item()
    msg = f"n={n} d={d} L={L} cg={cg}: rel err head {es:.3e} tail {ee:.3e}"
    if timing:
        for _ in range(2): K.apply_transport(x, ms, mt, T)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): K.apply_transport(x, ms, mt, T)
        e1.record(); torch.cuda.synchronize()
        ms_ = e0.elapsed_time(e1) / 5
        msg += f"  {ms_:.3f} ms  {2 * n * d * d * (L or 1) / ms_ / 1e9:.1f} TFLOP/s alg  {2 * n * d * 4 * (L or 1) / ms_ / 1e6:.0f} GB/s"
    print(msg, flush=True)
    assert es < 1e-5 and ee < 1e-5, msg
which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "cg1"):
    run(1000, 128, 1); run(4096, 512, 1); run(333, 96, 1); run(5000, 260, 1, L=3)
    run(1 << 18, 512, 1, timing=True); run(1 << 20, 128, 1, timing=True)
if which in ("all", "cg2"):
    run(1000, 256, 2); run(4096, 512, 2); run(333, 260, 2); run(5000, 384, 2, L=3)
    run(1 << 18, 512, 2, timing=True); run(1 << 20, 512, 2, timing=True); run(1 << 18, 1024, 2, timing=True)
print("OK")
