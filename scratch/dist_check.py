"""2+ GPU check of the sharded paths: packed statistics all-reduce and row-sharded Sinkhorn vs single-GPU results."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from ot_vae_lightning_b200 import kernels as K, parallel
from ot_vae_lightning_b200.ot import GaussianModel
from ot_vae_lightning_b200.synthetic import gaussian_latents, point_clouds
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
# statistics: every rank streams its shard; fit() does one packed all-reduce
n, d = 200_000, 256
x = gaussian_latents(n, d, seed=7, device=dev)
lo, hi = parallel.shard_rows(n, rank, world)
gm = GaussianModel(d, w2_cfg=dict(make_pd=True), reduce_on_update=False, dtype=torch.double, device=dev)
for s in range(lo, hi, 32768):
    gm.update(x[s:min(hi, s + 32768)])
gm.fit()
ref = GaussianModel(d, w2_cfg=dict(make_pd=True), reduce_on_update=False, dtype=torch.double, device=dev, ddp_reduce_func=None)
ref.update(x); ref.fit()
e_mean = ((gm.mean - ref.mean).norm() / ref.mean.norm()).item()
e_cov = ((gm.cov - ref.cov).norm() / ref.cov.norm()).item()
# Sinkhorn: rows sharded
N = M = 8192
xs, ys = point_clouds(N, M, 128, seed=5, device=dev)
a = torch.full((N,), 1.0 / N, device=dev); b = torch.full((M,), 1.0 / M, device=dev)
lo, hi = parallel.shard_rows(N, rank, world)
res = parallel.sharded_sinkhorn(xs[lo:hi].contiguous(), ys, a[lo:hi].contiguous(), b, reg=0.05, max_iter=30, threshold=0.0)
one = K.sinkhorn_points(xs, ys, a, b, reg=0.05, max_iter=30, threshold=0.0, scale=res["scale"])
e_u = (res["u_local"] - one["u"][lo:hi]).abs().max().item()
e_v = (res["v"] - one["v"]).abs().max().item()
out = torch.tensor([e_mean, e_cov, e_u, e_v], device=dev)
dist.all_reduce(out, op=dist.ReduceOp.MAX)
if rank == 0:
    print(f"world={world} n_obs={float(gm._n_obs)} mean_err={out[0]:.2e} cov_err={out[1]:.2e} sinkhorn |du|={out[2]:.2e} |dv|={out[3]:.2e}")
    assert float(gm._n_obs) == n and out[0] < 1e-6 and out[1] < 1e-5 and out[2] < 2e-3 and out[3] < 2e-3
    print("DIST OK")
dist.destroy_process_group()
