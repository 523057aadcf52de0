import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ot_vae_lightning_b200 import kernels as K
d = int(sys.argv[1]) if len(sys.argv) > 1 else 512
n = 1 << 18
x = torch.randn(n, d, device='cuda')
T = torch.randn(d, d, device='cuda', dtype=torch.float64) / d ** 0.5
ms = torch.randn(d, device='cuda', dtype=torch.float64); mt = torch.randn(d, device='cuda', dtype=torch.float64)
n_obs = torch.zeros((), dtype=torch.float64, device='cuda'); s = torch.zeros(d, dtype=torch.float64, device='cuda'); ss = torch.zeros(d, d, dtype=torch.float64, device='cuda')
for _ in range(2):
    y = K.apply_transport(x, ms, mt, T); K.stats_update(x, n_obs, s, ss, None)
torch.cuda.synchronize()
e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
e0.record(); y = K.apply_transport(x, ms, mt, T); e1.record(); K.stats_update(x, n_obs, s, ss, None); e2.record(); torch.cuda.synchronize()
print('apply ms', e0.elapsed_time(e1), 'TFLOP/s', 2 * n * d * d / e0.elapsed_time(e1) / 1e9, 'stats ms', e1.elapsed_time(e2), 'TFLOP/s', 2 * n * d * d / e1.elapsed_time(e2) / 1e9)
