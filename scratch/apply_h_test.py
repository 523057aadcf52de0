import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ot_vae_lightning_b200 import kernels as K, _native as N
lib = N.load()
torch.manual_seed(0)
for d, n in [(512, 65536), (512, 1 << 20), (128, 1 << 20), (256, 100000), (1024, 65536), (96, 777), (520, 3000)]:
    feat = torch.logspace(-2, 2, d, device='cuda')
    x = torch.randn(n, d, device='cuda') * feat + 3.0 * feat
    T = torch.randn(d, d, device='cuda', dtype=torch.float64) / d ** 0.5 / feat.double()
    ms = (3.0 * feat).double() + 0.01 * torch.randn(d, device='cuda', dtype=torch.float64) * feat.double()
    mt = torch.randn(d, device='cuda', dtype=torch.float64)
    nref = min(n, 4096)
    ref = (x[:nref].double() - ms) @ T.T + mt
    for forced, name in [(2 if d >= 256 else 1, "tf32"), (0, "fp16")]:
        lib.otkdbg_set_apply_cg(forced)
        for _ in range(2): y = K.apply_transport(x, ms, mt, T)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): y = K.apply_transport(x, ms, mt, T)
        e1.record(); torch.cuda.synchronize()
        ms_t = e0.elapsed_time(e1) / 5
        err = ((y[:nref].double() - ref).norm() / ref.norm()).item()
        last = ((y[-64:].double() - ((x[-64:].double() - ms) @ T.T + mt)).norm() / ref[-64:].norm()).item()
        print(f"d={d} n={n} {name}: {ms_t*1e3:.1f} us  {2*n*d*d/ms_t/1e9:.0f} TFLOP/s alg  {2*n*d*4/ms_t/1e6:.0f} GB/s  rel err {err:.2e} (tail {last:.2e})", flush=True)
    # overflow: an outlier far outside the FP16 window after the head rows -> gated TF32 recompute
    lib.otkdbg_set_apply_cg(0)
    x2 = x.clone(); x2[n // 2, 5] = 1.0e9
    y2 = K.apply_transport(x2, ms, mt, T)
    r2 = (x2[n // 2 - 2:n // 2 + 2].double() - ms) @ T.T + mt
    print(f"   outlier row: rel err {((y2[n // 2 - 2:n // 2 + 2].double() - r2).norm() / r2.norm()).item():.2e}; others {((y2[:nref].double() - ref).norm() / ref.norm()).item():.2e}", flush=True)
lib.otkdbg_set_apply_cg(0)
