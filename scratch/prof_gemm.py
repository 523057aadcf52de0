import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ot_vae_lightning_b200 import kernels as K
A = torch.randn(4096, 512, device='cuda'); B = torch.randn(512, 512, device='cuda')
for _ in range(3):
    C = K.gemm(A, B, engine=2, nn=True)
torch.cuda.synchronize()
print('ok', float(C[0, 0]))
