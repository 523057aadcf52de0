import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ot_vae_lightning_b200 import kernels as K, _native as N
lib = N.load()
dev = torch.device('cuda', 0); torch.manual_seed(0)
d, rows = 128, 300000
for name, scale, mean in [("homog", torch.ones(d, device=dev), 0.7), ("hetero", torch.logspace(-1, 1, d, device=dev), 0.7), ("hetero0", torch.logspace(-1, 1, d, device=dev), 0.0)]:
    x = torch.randn(rows, d, device=dev) * scale + mean
    x64 = x.double(); want = x64.T @ x64
    wantc = (x64.cpu().T @ x64.cpu()).to(dev)
    for cg in (0, 1):
        lib.otkdbg_set_stats_cg(cg)
        n = torch.zeros((), dtype=torch.float64, device=dev); s = torch.zeros(d, dtype=torch.float64, device=dev); ss = torch.zeros(d, d, dtype=torch.float64, device=dev)
        K.stats_update(x, n, s, ss, None)
        err = ss - want
        dg = torch.diagonal(want); sc = torch.sqrt(dg[:, None] * dg[None, :])
        rel_entry = (err.abs() / sc)
        print(f"{name} cg={cg}: frob {float(err.norm()/want.norm()):.2e} (vs cpu ref {float((ss-wantc).norm()/wantc.norm()):.2e}); gpu-vs-cpu ref {float((want-wantc).norm()/wantc.norm()):.1e}; "
              f"diag rel max {float((torch.diagonal(err).abs()/dg).max()):.2e}; offdiag entry/sqrt max {float(rel_entry.max()):.2e} median {float(rel_entry.median()):.2e}; "
              f"diag err small-feat {float((torch.diagonal(err)/dg)[:8].abs().mean()):.1e} big-feat {float((torch.diagonal(err)/dg)[-8:].abs().mean()):.1e}")
