import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ot_vae_lightning_b200 import kernels as K, _native as N
lib = N.load()
torch.manual_seed(0)
for d, n in [(512, 65536), (512, 1 << 20), (256, 1 << 20), (192, 100000), (1024, 65536), (384, 250)]:
    x = torch.randn(n, d, device='cuda') * torch.logspace(-1, 1, d, device='cuda') + 1.0
    xd = x.double(); ref = xd.T @ xd
    for cg in [2 if d >= 512 else 1, 0]:   # forced TF32 path vs automatic (FP16 split)
        lib.otkdbg_set_stats_cg(cg)
        n_obs = torch.zeros((), dtype=torch.float64, device='cuda'); s = torch.zeros(d, dtype=torch.float64, device='cuda'); ss = torch.zeros(d, d, dtype=torch.float64, device='cuda')
        for _ in range(2): K.stats_update(x, n_obs, s, ss, None)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(8): K.stats_update(x, n_obs, s, ss, None)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 8
        err = ((ss - ref * 10).norm() / (ref * 10).norm()).item()
        serr = ((s - xd.sum(0) * 10).norm() / (xd.sum(0) * 10).norm()).item()
        print(f"d={d} n={n} force_cg={cg}: {ms*1e3:.1f} us  {2*n*d*d/ms/1e9:.0f} TFLOP/s alg  ss rel err {err:.2e} sum err {serr:.2e} asym {float((ss-ss.T).abs().max()):.1e}", flush=True)
lib.otkdbg_set_stats_cg(0)
