import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ot_vae_lightning_b200 import _native as N
if len(sys.argv) > 1:
    N.LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'var_' + sys.argv[1], 'libotk.so')
from ot_vae_lightning_b200 import kernels as K
for d, n, reps in [(256, 65536, 16), (384, 65536, 16), (640, 65536, 8), (768, 65536, 8)]:
    x = torch.randn(n, d, device='cuda') + 1.0
    n_obs = torch.zeros((), dtype=torch.float64, device='cuda'); s = torch.zeros(d, dtype=torch.float64, device='cuda'); ss = torch.zeros(d, d, dtype=torch.float64, device='cuda')
    for _ in range(2): K.stats_update(x, n_obs, s, ss, None)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): K.stats_update(x, n_obs, s, ss, None)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    xd = x.double(); ref = xd.T @ xd * (reps + 2)
    print(f"d={d} n={n}: {ms*1e3:.1f} us/update  {n*d*4/ms/1e6:.0f} GB/s  {2*n*d*d/ms/1e9:.0f} TFLOP/s  rel err {((ss - ref).norm() / ref.norm()).item():.2e} asym {(ss-ss.T).abs().max().item():.1e}", flush=True)
