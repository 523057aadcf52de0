import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ot_vae_lightning_b200 import kernels as K
def t(fn, n=50):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for d in (128, 256, 512, 1024):
    A = torch.randn(d, d, device='cuda'); B = torch.randn(d, d, device='cuda')
    ref = A.double() @ B.double().T
    out = K.gemm(A, B, engine=2)
    err = ((out.double() - ref).norm() / ref.norm()).item()
    ms = t(lambda: K.gemm(A, B, engine=2))
    print(f"d={d} 3xTF32 (incl. operand split kernels): {ms*1e3:.1f} us  rel err {err:.2e}", flush=True)
    src = torch.randn(4 * d, d, device='cuda', dtype=torch.float64)
    cov = src.T @ src / (4 * d) + 0.05 * torch.eye(d, device='cuda', dtype=torch.float64)
    ms = t(lambda: K.sqrtm_pair(cov), 10)
    print(f"d={d} sqrtm_pair: {ms*1e3:.1f} us", flush=True)
