import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ot_vae_lightning_b200 import _native as N
N.LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "var_GT", "libotk.so")
from ot_vae_lightning_b200 import kernels as K
for d in (128, 512):
    src = torch.randn(4 * d, d, device='cuda', dtype=torch.float64)
    cov = src.T @ src / (4 * d) + 0.05 * torch.eye(d, device='cuda', dtype=torch.float64)
    print("==== d", d, flush=True)
    K.sqrtm_pair(cov); torch.cuda.synchronize()
