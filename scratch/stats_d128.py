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
# outlier far outside the FP16 window -> device flag -> gated TF32 fallback recomputes
lib.otkdbg_set_stats_cg(0)
for d, n in [(128, 5000), (96, 777), (64, 70000)]:
    x = torch.randn(n, d, device='cuda') * torch.logspace(-2, 1, d, device='cuda') + 3.0
    for outlier in [False, True]:
        if outlier: x[n // 2, 3] = 1.0e7
        n_obs = torch.zeros((), dtype=torch.float64, device='cuda'); s = torch.zeros(d, dtype=torch.float64, device='cuda'); ss = torch.zeros(d, d, dtype=torch.float64, device='cuda')
        K.stats_update(x, n_obs, s, ss, None)
        xd = x.double(); ref = xd.T @ xd; xc = xd - xd.mean(0)
        cov = ss / n - torch.outer(s / n, s / n); covref = xc.T @ xc / n
        print(f"d={d} n={n} outlier={outlier}: ss rel err {((ss - ref).norm() / ref.norm()).item():.2e}  sum rel err {((s - xd.sum(0)).norm() / xd.sum(0).norm()).item():.2e}  cov rel err {((cov - covref).norm() / covref.norm()).item():.2e}  max elementwise cov err / sqrt(cii cjj) {((cov - covref).abs() / torch.sqrt(torch.outer(covref.diag(), covref.diag()))).max().item():.2e}", flush=True)
