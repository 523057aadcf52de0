import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ot_vae_lightning_b200 import kernels as K
d = int(sys.argv[1]) if len(sys.argv) > 1 else 512
src = torch.randn(4 * d, d, device='cuda', dtype=torch.float64)
cov = src.T @ src / (4 * d) + 0.05 * torch.eye(d, device='cuda', dtype=torch.float64)
os.environ.setdefault("OTK_NS_GRAPHS", "0")
K.sqrtm_pair(cov); torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
K.sqrtm_pair(cov); torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
