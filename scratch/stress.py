import sys, os, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ot_vae_lightning_b200 import kernels as K
dev = torch.device('cuda', 0)
random.seed(0); torch.manual_seed(0)
bad = 0
cases = [(1, 300000, 128), (1, 70000, 64), (4, 9000, 96), (1, 65536, 512), (2, 20000, 384), (1, 30000, 1024), (3, 4000, 256), (1, 500000, 256)]
for L, rows, d in cases:
    x = torch.randn(L, rows, d, device=dev) * torch.logspace(-1, 1, d, device=dev) + 0.7
    x64 = x.double(); want = x64.transpose(1, 2) @ x64; wsum = x64.sum(1)
    T = torch.randn(L, d, d, device=dev, dtype=torch.float64) / d ** 0.5; ms = x64.mean(1); mt = torch.randn(L, d, device=dev, dtype=torch.float64)
    nref = min(rows, 2048)
    yref = (x64[:, -nref:] - ms.unsqueeze(1)) @ T.transpose(1, 2) + mt.unsqueeze(1)
    first = None
    for it in range(25):
        n = torch.zeros(L, dtype=torch.float64, device=dev); s = torch.zeros(L, d, dtype=torch.float64, device=dev); ss = torch.zeros(L, d, d, dtype=torch.float64, device=dev)
        K.stats_update(x, n, s, ss, None)
        e1 = ((ss - want).flatten(1).norm(dim=1) / want.flatten(1).norm(dim=1)).max().item()
        e2 = ((s - wsum).norm() / wsum.norm()).item()
        y = K.apply_transport(x, ms, mt, T)
        e3 = ((y[:, -nref:].double() - yref).norm() / yref.norm()).item()
        if first is None: first = (ss.clone(), y.clone())
        same = bool((ss == first[0]).all()) , bool((y == first[1]).all())
        if e1 > 2e-5 or e2 > 1e-6 or e3 > 2e-5 or not same[1]:
            bad += 1; print(f"BAD L={L} rows={rows} d={d} iter {it}: ss {e1:.2e} sum {e2:.2e} y {e3:.2e} y-repro {same[1]}", flush=True)
    print(f"L={L} rows={rows} d={d}: ss {e1:.2e} sum {e2:.2e} y {e3:.2e}; ss bitwise reproducible: {same[0]}", flush=True)
print("bad:", bad)
