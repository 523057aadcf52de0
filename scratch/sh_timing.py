import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ot_vae_lightning_b200 import _native as N
if len(sys.argv) > 1:
    N.LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "var_" + sys.argv[1], "libotk.so")
from ot_vae_lightning_b200 import kernels as K
for d in (128,):
    x = torch.randn(1 << 20, d, device='cuda') + 1.0
    n_obs = torch.zeros((), dtype=torch.float64, device='cuda'); s = torch.zeros(d, dtype=torch.float64, device='cuda'); ss = torch.zeros(d, d, dtype=torch.float64, device='cuda')
    K.stats_update(x, n_obs, s, ss, None)
    torch.cuda.synchronize()
    print("----", sys.argv[1:], d, flush=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); K.stats_update(x, n_obs, s, ss, None); e1.record(); torch.cuda.synchronize()
    print("update call us:", e0.elapsed_time(e1) * 1e3, flush=True)
