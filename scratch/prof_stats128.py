import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ot_vae_lightning_b200 import kernels as K
d, n = 128, 1 << 20
x = torch.randn(n, d, device='cuda') + 1.0
n_obs = torch.zeros((), dtype=torch.float64, device='cuda'); s = torch.zeros(d, dtype=torch.float64, device='cuda'); ss = torch.zeros(d, d, dtype=torch.float64, device='cuda')
for _ in range(2): K.stats_update(x, n_obs, s, ss, None)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
K.stats_update(x, n_obs, s, ss, None)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
