"""Row-sharded Sinkhorn alone (cfg3: N = M = 65536, d = 128, eps = 0.05, 100 iterations), one process per GPU:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        profiles/tools/sinkhorn_scale.py
Prints iterations/s (CUDA events, max over ranks) and the plan statistics of the sharded solution."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
from ot_vae_lightning_b200 import kernels as K  # noqa: E402
from ot_vae_lightning_b200 import parallel  # noqa: E402
from ot_vae_lightning_b200.synthetic import point_clouds  # noqa: E402

N, D, EPS, ITERS = 65536, 128, 0.05, int(os.environ.get("SK_ITERS", "100"))
x, y = point_clouds(N, N, D, seed=99, device=dev)
a = torch.full((N,), 1.0 / N, device=dev)
lo, hi = parallel.shard_rows(N, rank, world)
xl, al = x[lo:hi].contiguous(), a[lo:hi].contiguous()
scale = parallel.global_cost_scale(xl, y)
state = {}


def run():
    out = parallel.sharded_sinkhorn(xl, y, al, a, reg=EPS, max_iter=ITERS, threshold=0.0, scale=scale, plan=state.get("plan"))
    state["plan"] = out["plan"]
    return out


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


run()
best = 1e30
for _ in range(3):
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = run()
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    best = min(best, float(t))
chk = parallel.sharded_summary(xl, y, al, a, out["u_local"], out["v"], scale, EPS)
if rank == 0:
    peer = state["plan"].get("peer")
    how = "peer-memory exchange" if peer is not None else ("NCCL all-gather" if world > 1 else "single GPU")
    tmo = int(peer["ctrl"][2]) if peer is not None else 0
    print(f"n_gpus={world} iters/s={ITERS / (best * 1e-3):.1f} ms/iter={best / ITERS:.4f} exchange={how} timeouts={tmo} check={chk}", flush=True)
state.clear()
del out
import gc  # noqa: E402
gc.collect()
torch.cuda.synchronize()
if world > 1:
    dist.destroy_process_group()
