import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ot_vae_lightning_b200 import kernels as Kn, _native as N
lib = N.load(); lib.otkdbg_set_stats_cg(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
L, rows, d = 10, 1500, 1024
g = torch.Generator(device="cuda").manual_seed(7 * rows + d)
feat = torch.logspace(-3, 3, d, device="cuda")
x = torch.randn(L, rows, d, device="cuda", generator=g) * feat + 2.0 * feat
for outlier in (False, True):
    if outlier: x[:, rows // 2 + 3, d // 3] = 3.0e8
    x64 = x.double(); want = x64.transpose(1, 2) @ x64
    dg = torch.diagonal(want, dim1=1, dim2=2); scale = torch.sqrt(dg.unsqueeze(-1) * dg.unsqueeze(-2))
    bad = 0
    for it in range(40):
        n = torch.zeros(L, dtype=torch.float64, device="cuda"); s = torch.zeros(L, d, dtype=torch.float64, device="cuda"); ss = torch.zeros(L, d, d, dtype=torch.float64, device="cuda")
        Kn.stats_update(x, n, s, ss, None)
        e = ((ss - want).abs() / scale)
        m = float(e.max())
        if m > 5e-6:
            bad += 1
            idx = torch.nonzero(e > 5e-6)
            ls = sorted(set(idx[:, 0].tolist())); bi = sorted(set((idx[:, 1] // 128).tolist())); bj = sorted(set((idx[:, 2] // 128).tolist()))
            print(f"outlier={outlier} iter {it}: max err {m:.2e}, {idx.shape[0]} bad entries, l in {ls}, row blocks {bi}, col blocks {bj}, sum err {float(((s - x64.sum(1)).abs() / x64.sum(1).abs()).max()):.1e}", flush=True)
    print(f"outlier={outlier}: {bad}/40 bad", flush=True)
