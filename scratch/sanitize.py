import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ot_vae_lightning_b200 import kernels as K
from ot_vae_lightning_b200.ot import GaussianTransport
from ot_vae_lightning_b200.synthetic import gaussian_latents, point_clouds
dev = torch.device('cuda', 0)
for d, n in [(128, 3000), (512, 5000), (260, 1111)]:
    src = gaussian_latents(n, d, seed=1, device=dev); tgt = gaussian_latents(n, d, seed=2, device=dev, shift=0.5, scale=1.5)
    cfg = dict(dtype=torch.double, device=dev)
    op = GaussianTransport(d, transport_cfg=dict(make_pd=True), source_cfg=dict(cfg), target_cfg=dict(cfg)).to(dev)
    op.update(source_samples=src, target_samples=tgt); op.update(source_samples=src[:200], target_samples=tgt[:200])
    w2 = op.compute(); y = op.transport(src); y2 = K.apply_transport(src[:777], op.source_model.mean, op.target_model.mean, op.transport_operator)
    far = src.clone(); far[5, 3] = 1e9
    op.source_model.update(far); y3 = op.transport(far)
    torch.cuda.synchronize(); print(d, float(w2), float(y.abs().mean()))
x, yv = point_clouds(700, 520, 64, seed=3, device=dev)
a = torch.full((700,), 1 / 700, device=dev); b = torch.full((520,), 1 / 520, device=dev)
r = K.sinkhorn_points(x, yv, a, b, reg=0.05, max_iter=6, threshold=0.0)
torch.cuda.synchronize(); print("sinkhorn", r["summary"].tolist())
