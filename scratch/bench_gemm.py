import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ot_vae_lightning_b200 import kernels as K
def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for (M, N, Kd) in [(32768, 512, 512), (32768, 128, 128), (8192, 1024, 1024)]:
    A = torch.randn(M, Kd, device='cuda'); B = torch.randn(N, Kd, device='cuda')
    for eng, name in [(3, '1xTF32'), (2, '3xTF32')]:
        ms = t(lambda: K.gemm(A, B, engine=eng))
        print(f"M={M} N={N} K={Kd} {name}: {ms:.3f} ms  {2*M*N*Kd/ms/1e9:.1f} TFLOP/s algorithmic")
