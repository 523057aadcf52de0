"""Two 65536-row source / target chunks of D-wide latents through update -> compute -> transport, the second pass inside a
cudaProfiler window (ncu --profile-from-start off):   python profiles/tools/stats_step.py [D]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from ot_vae_lightning_b200.ot import GaussianTransport
from ot_vae_lightning_b200.synthetic import gaussian_latents
d = int(sys.argv[1]) if len(sys.argv) > 1 else 512
n = 1 << 17
dev = torch.device('cuda', 0)
src = gaussian_latents(n, d, seed=1234, device=dev)
tgt = gaussian_latents(n, d, seed=4321, device=dev, shift=0.5, scale=1.5)
cfg = dict(dtype=torch.double, device=dev, reduce_on_update=False)
op = GaussianTransport(d, transport_cfg=dict(make_pd=True), source_cfg=dict(cfg), target_cfg=dict(cfg)).to(dev)
def step():
    op.reset()
    for lo in range(0, n, 65536):
        op.update(source_samples=src[lo:lo + 65536], target_samples=tgt[lo:lo + 65536])
    w2 = op.compute()
    out = op.transport(src[:65536])
    return w2
step(); torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
w2 = step(); torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print('w2', float(w2))
