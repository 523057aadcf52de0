import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ot_vae_lightning_b200 import kernels as K, _native as N
lib = N.load()
torch.manual_seed(0)
def run(n, d, cg, L=None, timing=False, offset=3.0):
    lib.otkdbg_set_stats_cg(cg)
    lead = () if L is None else (L,)
    g = torch.Generator(device='cuda'); g.manual_seed(n + d)
    A = torch.randn(*lead, d, d, device='cuda', generator=g) / d ** 0.5
    x = torch.randn(*lead, n, d, device='cuda', generator=g) @ A + offset * torch.randn(*lead, 1, d, device='cuda', generator=g)
    n_obs = torch.zeros(lead, dtype=torch.float64, device='cuda'); s = torch.zeros(*lead, d, dtype=torch.float64, device='cuda'); ss = torch.zeros(*lead, d, d, dtype=torch.float64, device='cuda')
    K.stats_update(x, n_obs, s, ss, None); torch.cuda.synchronize()
    xd = x.double()
    rs = xd.sum(-2); rss = xd.transpose(-1, -2) @ xd
    mean_r = rs / n; cov_r = rss / n - mean_r.unsqueeze(-1) * mean_r.unsqueeze(-2)
    mean = s / n; cov = ss / n - mean.unsqueeze(-1) * mean.unsqueeze(-2)
    e_n = (n_obs - n).abs().max().item()
    e_s = ((s - rs).norm() / rs.norm()).item(); e_ss = ((ss - rss).norm() / rss.norm()).item()
    e_cov = ((cov - cov_r).norm() / cov_r.norm()).item()
    asym = (ss - ss.transpose(-1, -2)).abs().max().item()
    msg = f"n={n} d={d} L={L} cg={cg}: n_err {e_n} sum {e_s:.2e} sumcov {e_ss:.2e} cov {e_cov:.2e} asym {asym:.1e}"
    if timing:
        for _ in range(2): K.stats_update(x, n_obs, s, ss, None)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): K.stats_update(x, n_obs, s, ss, None)
        e1.record(); torch.cuda.synchronize()
        ms_ = e0.elapsed_time(e1) / 5
        msg += f"  {ms_:.3f} ms  {2 * n * d * d * (L or 1) / ms_ / 1e9:.1f} TFLOP/s alg  {n * d * 4 * (L or 1) / ms_ / 1e6:.0f} GB/s"
    print(msg, flush=True)
    assert e_n == 0 and e_s < 1e-6 and e_ss < 1e-5 and e_cov < 1e-4 and asym == 0, msg
which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "cg1"):
    run(1000, 128, 1); run(4096, 512, 1); run(333, 96, 1); run(5000, 260, 1, L=3); run(100, 64, 1); run(70000, 384, 1)
    run(1 << 18, 512, 1, timing=True); run(1 << 20, 128, 1, timing=True); run(1 << 19, 256, 1, timing=True)
if which in ("all", "cg2"):
    run(1000, 256, 2); run(4096, 512, 2); run(333, 260, 2); run(5000, 384, 2, L=3); run(70000, 640, 2)
    run(1 << 18, 512, 2, timing=True); run(1 << 20, 512, 2, timing=True); run(1 << 18, 1024, 2, timing=True); run(1 << 19, 256, 2, timing=True)
print("OK")
