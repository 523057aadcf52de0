"""One `GaussianTransport.compute()` (fit both models, W2^2, map, prepared operator) inside a cudaProfiler window, for
`ncu --profile-from-start off` launch lists, and its wall/CUDA-event time outside the profiler.
    python profiles/tools/compute_only.py D [ROWS [CLASSES]]      # OTK_NS_GRAPHS=0 so that ncu sees the individual launches
CLASSES > 1: `GaussianTransport(CLASSES, D)` - one operator per class, batched through the same launches (cfg4)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch  # noqa: E402
from ot_vae_lightning_b200.ot import GaussianTransport  # noqa: E402
from ot_vae_lightning_b200.synthetic import gaussian_latents  # noqa: E402

d = int(sys.argv[1]) if len(sys.argv) > 1 else 512
rows = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
classes = int(sys.argv[3]) if len(sys.argv) > 3 else 1
dev = torch.device("cuda", 0)
cfg = dict(dtype=torch.double, device=dev, reduce_on_update=False)
size = (classes, d) if classes > 1 else (d,)
op = GaussianTransport(*size, transport_cfg=dict(make_pd=True), source_cfg=dict(cfg), target_cfg=dict(cfg)).to(dev)
src = gaussian_latents(rows, d, seed=1234, device=dev)
tgt = gaussian_latents(rows, d, seed=4321, device=dev, shift=0.5, scale=1.5)
if classes > 1:
    src = torch.stack([gaussian_latents(rows, d, seed=1234 + k, device=dev) for k in range(classes)])
    tgt = torch.stack([gaussian_latents(rows, d, seed=4321 + k, device=dev, shift=0.5, scale=1.5) for k in range(classes)])
op.update(source_samples=src, target_samples=tgt)
for _ in range(3):
    op.compute()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
for _ in range(10):
    op.compute()
e1.record()
torch.cuda.synchronize()
print(f"d={d} classes={classes}: compute() {e0.elapsed_time(e1) / 10:.3f} ms (events), {(time.perf_counter() - t0) * 100:.3f} ms (host)")
torch.cuda.cudart().cudaProfilerStart()
op.compute()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
