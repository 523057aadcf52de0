import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ot_vae_lightning_b200 import kernels as K
d = int(sys.argv[1]) if len(sys.argv) > 1 else 512
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 18
x = torch.randn(n, d, device='cuda')
T = torch.randn(d, d, device='cuda', dtype=torch.float64) / d ** 0.5
ms = torch.randn(d, device='cuda', dtype=torch.float64); mt = torch.randn(d, device='cuda', dtype=torch.float64)
for _ in range(3):
    y = K.apply_transport(x, ms, mt, T)
torch.cuda.synchronize()
print("done")
