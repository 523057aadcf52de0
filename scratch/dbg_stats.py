import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ot_vae_lightning_b200 import kernels as K, _native as N
lib = N.load()
d = int(sys.argv[1]); n = int(sys.argv[2]); cg = int(sys.argv[3])
lib.otkdbg_set_stats_cg(cg)
x = torch.randn(n, d, device='cuda') + 1.0
n_obs = torch.zeros((), dtype=torch.float64, device='cuda'); s = torch.zeros(d, dtype=torch.float64, device='cuda'); ss = torch.zeros(d, d, dtype=torch.float64, device='cuda')
for mode in [0, 1, 2, 4, 6, 7, 8, 15, 9, 14]:
    lib.otkdbg_set_stats_dbg(mode)
    for _ in range(2): K.stats_update(x, n_obs, s, ss, None)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): K.stats_update(x, n_obs, s, ss, None)
    e1.record(); torch.cuda.synchronize()
    print(f"d={d} n={n} cg={cg} mode={mode:2d} (noTMA={mode&1} noA={(mode>>1)&1} noB={(mode>>2)&1} noMMA={(mode>>3)&1}): {e0.elapsed_time(e1)/5:.3f} ms", flush=True)
