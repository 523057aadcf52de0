import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ot_vae_lightning_b200 import kernels as K, _native as N
lib = N.load()
d, L = 128, 148
torch.manual_seed(1)
for rows in [64, 128, 192, 256, 320, 384, 448, 512, 576, 640, 704, 768, 1024, 1088, 2048, 2112]:
    x = torch.randn(L, rows, d, device='cuda') + 1.0
    n_obs = torch.zeros(L, dtype=torch.float64, device='cuda'); s = torch.zeros(L, d, dtype=torch.float64, device='cuda'); ss = torch.zeros(L, d, d, dtype=torch.float64, device='cuda')
    K.stats_update(x, n_obs, s, ss, None)
    xd = x.double(); ref = xd.transpose(1, 2) @ xd
    err = ((ss - ref).flatten(1).norm(dim=1) / ref.flatten(1).norm(dim=1))
    print(f"rows={rows} (steps {rows // 64}): rel err max {err.max().item():.2e} median {err.median().item():.2e}", flush=True)
