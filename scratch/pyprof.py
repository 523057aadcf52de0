import sys, os, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ot_vae_lightning_b200.ot import GaussianTransport
from ot_vae_lightning_b200.synthetic import gaussian_latents
dev = torch.device('cuda', 0)
d, n, bs = 128, 10000, 250
src = gaussian_latents(n, d, seed=1, device=dev); tgt = gaussian_latents(n, d, seed=2, device=dev, shift=0.5, scale=1.5)
cfg = dict(dtype=torch.double, device=dev, reduce_on_update=False)
op = GaussianTransport(d, transport_cfg=dict(make_pd=True), source_cfg=dict(cfg), target_cfg=dict(cfg)).to(dev)
op.update(source_samples=src[:bs], target_samples=tgt[:bs]); op.compute(); op.transport(src[:bs]); torch.cuda.synchronize()
batches = [src[lo:lo + bs] for lo in range(0, n, bs)]
def upd():
    for _ in range(25):
        for b in batches: op.source_model.update(b)
    torch.cuda.synchronize()
def tr():
    for _ in range(25):
        for b in batches: op.transport(b)
    torch.cuda.synchronize()
for fn in (upd, tr):
    pr = cProfile.Profile(); pr.enable(); fn(); pr.disable()
    st = pstats.Stats(pr); st.sort_stats('tottime')
    print(f"==== {fn.__name__}: 1000 calls"); st.print_stats(14)
