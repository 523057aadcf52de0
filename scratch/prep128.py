import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ot_vae_lightning_b200 import kernels as K
from ot_vae_lightning_b200.ot import GaussianTransport
from ot_vae_lightning_b200.synthetic import gaussian_latents
dev = torch.device('cuda', 0)
def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for d in (128, 512):
    n = 1 << 20
    x = gaussian_latents(n, d, seed=77, device=dev)
    cfg = dict(dtype=torch.double, device=dev, reduce_on_update=False)
    for tgt_kind in ("affine", "other"):
        op = GaussianTransport(d, transport_cfg=dict(make_pd=True), source_cfg=dict(cfg), target_cfg=dict(cfg)).to(dev)
        tgt = x * 1.5 + 0.5 if tgt_kind == "affine" else gaussian_latents(n, d, seed=78, device=dev, shift=0.5, scale=1.5)
        op.update(source_samples=x, target_samples=tgt)
        op.compute()
        prep = op._prepared_operator()
        print(f"d={d} target={tgt_kind}: transport {t(lambda: op.transport(x)):.3f} ms; prepared.apply {t(lambda: prep.apply(x)):.3f} ms; functional {t(lambda: K.apply_transport(x, op.source_model.mean, op.target_model.mean, op.transport_operator)):.3f} ms", flush=True)
