import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ot_vae_lightning_b200 import kernels as K, _native as N
lib = N.load()
d = int(sys.argv[1]); n = int(sys.argv[2])
x = torch.randn(n, d, device='cuda')
T = torch.randn(d, d, device='cuda', dtype=torch.float64) / d ** 0.5
ms = torch.randn(d, device='cuda', dtype=torch.float64); mt = torch.randn(d, device='cuda', dtype=torch.float64)
for mode in [0, 1]:
    lib.otkdbg_set_apply_dbg(mode)
    for _ in range(2): K.apply_transport(x, ms, mt, T)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): K.apply_transport(x, ms, mt, T)
    e1.record(); torch.cuda.synchronize()
    print(f"d={d} n={n} mode={mode}: {e0.elapsed_time(e1)/5:.3f} ms", flush=True)
