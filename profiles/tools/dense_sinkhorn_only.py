"""`sinkhorn_log` on a materialised cost (the reference signature) at N = M = SIZE, fp32: a few iterations inside a
cudaProfiler window (ncu --profile-from-start off) and its CUDA-event time outside the profiler.
    python profiles/tools/dense_sinkhorn_only.py [SIZE=32768] [ITERS=3]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch  # noqa: E402
from ot_vae_lightning_b200 import kernels as K  # noqa: E402
from ot_vae_lightning_b200.synthetic import point_clouds  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda", 0)
x, y = point_clouds(n, n, 128, seed=99, device=dev)
a = torch.full((n,), 1.0 / n, device=dev)
scale = 1.0 / float(K.cost_max(x, y, 0).item())
C = K.cost_matrix(x, y, 0, scale)
run = lambda: K.sinkhorn_dense(a, a, C, 0.05, iters, 0.0, want_plan=False)
run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
run()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
print(f"N=M={n}: {ms:.3f} ms/iter, {2.0 * n * n * 4 / (ms * 1e-3) / 1e9:.0f} GB/s algorithmic (2 N M 4 bytes per iteration)")
torch.cuda.cudart().cudaProfilerStart()
run()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
