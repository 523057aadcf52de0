import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ot_vae_lightning_b200.ot import GaussianTransport
from ot_vae_lightning_b200.synthetic import gaussian_latents, mixture_latents
dev = torch.device('cuda', 0)
def timed(fn, n=3):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): r = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3, r
cfg = dict(dtype=torch.double, device=dev, reduce_on_update=False)
# cfg1: N = 10k, d = 128, batches of 250
d, n, bs = 128, 10000, 250
src = gaussian_latents(n, d, seed=1, device=dev); tgt = gaussian_latents(n, d, seed=2, device=dev, shift=0.5, scale=1.5)
op = GaussianTransport(d, transport_cfg=dict(make_pd=True), source_cfg=dict(cfg), target_cfg=dict(cfg)).to(dev)
def step1():
    op.reset()
    for lo in range(0, n, bs): op.update(source_samples=src[lo:lo + bs], target_samples=tgt[lo:lo + bs])
    w2 = op.compute()
    outs = [op.transport(src[lo:lo + bs]) for lo in range(0, n, bs)]
    return w2
ms, w2 = timed(step1)
print(f"cfg1 (N=10k, d=128, batches of 250): {ms:.2f} ms/step -> {n / ms * 1e3 / 1e6:.2f} M latents/s, w2 {float(w2):.4f}")
def upd(): 
    op.source_model.update(src[:bs])
ms, _ = timed(upd, 20); print(f"   one update(250 x 128): {ms * 1e3:.0f} us")
ms, _ = timed(lambda: op.compute(), 5); print(f"   compute() d=128: {ms:.2f} ms")
ms, _ = timed(lambda: op.transport(src[:bs]), 20); print(f"   one transport(250 x 128): {ms * 1e3:.0f} us")
# cfg4: 10 classes x d = 1024
L, d, n = 10, 1024, 20000
xs = torch.stack([mixture_latents(n, d, seed=10 + c, device=dev) for c in range(L)])
xt = torch.stack([gaussian_latents(n, d, seed=50 + c, device=dev, shift=0.3, scale=1.2) for c in range(L)])
op4 = GaussianTransport(L, d, transport_cfg=dict(make_pd=True), source_cfg=dict(cfg), target_cfg=dict(cfg)).to(dev)
def step4():
    op4.reset()
    for lo in range(0, n, 4000): op4.update(source_samples=xs[:, lo:lo + 4000], target_samples=xt[:, lo:lo + 4000])
    w2 = op4.compute()
    y = op4.transport(xs)
    return w2
ms, w2 = timed(step4, 2)
print(f"cfg4 (10 classes x d=1024, 20k latents each): {ms:.1f} ms/step -> {L * n / ms * 1e3 / 1e6:.2f} M latents/s; w2[:3] {w2[:3].tolist()}")
ms, _ = timed(lambda: op4.compute(), 2); print(f"   compute() 10 x 1024: {ms:.1f} ms")
ms, _ = timed(lambda: op4.transport(xs), 3); print(f"   transport 10 x 20000 x 1024: {ms:.2f} ms")
ms, _ = timed(lambda: op4.source_model.update(xs), 3); print(f"   update 10 x 20000 x 1024: {ms:.2f} ms")
