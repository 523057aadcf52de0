import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ot_vae_lightning_b200 import kernels as K, _native as N
lib = N.load()
for d, n in [(128, 1 << 20), (128, 65536), (64, 1 << 20), (96, 1 << 20), (128, 10000)]:
    x = torch.randn(n, d, device='cuda') + 1.0
    for cg in [1, 0]:   # 1 = force the old single-CTA path, 0 = automatic (shared-tile mode for d <= 128)
        lib.otkdbg_set_stats_cg(cg)
        n_obs = torch.zeros((), dtype=torch.float64, device='cuda'); s = torch.zeros(d, dtype=torch.float64, device='cuda'); ss = torch.zeros(d, d, dtype=torch.float64, device='cuda')
        for _ in range(2): K.stats_update(x, n_obs, s, ss, None)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): K.stats_update(x, n_obs, s, ss, None)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        xd = x.double()
        ref = xd.T @ xd * 12
        err = ((ss - ref).norm() / ref.norm()).item()
        print(f"d={d} n={n} force_cg={cg}: {ms*1e3:.1f} us  {n*d*4/ms/1e6:.0f} GB/s  rel err {err:.2e}", flush=True)
