import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ot_vae_lightning_b200 import kernels as K
from ot_vae_lightning_b200.synthetic import point_clouds
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
x, y = point_clouds(n, n, 128, seed=99, device='cuda')
a = torch.full((n,), 1.0 / n, device='cuda')
scale = 1.0 / float(K.cost_max(x, y, 0).item())
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
res = K.sinkhorn_points(x, y, a, a, reg=0.05, max_iter=iters, threshold=0.0, scale=scale, want_summary=False, want_iters=False)
e1.record(); torch.cuda.synchronize()
print('scale', scale, 'ms/iter', e0.elapsed_time(e1) / iters)
