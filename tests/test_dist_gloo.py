"""N>1 host logic on CPU: world_size-2 gloo process group, kernels swapped for the oracle-backed stand-ins.
Covers (a) the packed one-shot all-reduce of the sufficient statistics through the DDPMixin seam and
(b) the row-sharded Sinkhorn driver (all-gather of column LSE partials + scalar all-reduce for the stop rule)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    from oracle import ot_oracle as O
    from tests import fake_kernels
    import ot_vae_lightning_b200.ot as ot
    from ot_vae_lightning_b200 import parallel
    try:
        with fake_kernels.installed():
            # ---- (a) statistics: each rank streams its shard, fit() reduces once
            g = torch.Generator().manual_seed(0)
            x = torch.randn(600, 12, generator=g) * 2 + 1
            lo, hi = parallel.shard_rows(600, rank, world)
            gm = ot.GaussianModel(12, w2_cfg=dict(make_pd=True), reduce_on_update=False, dtype=torch.double, device="cpu")
            for s in range(lo, hi, 50):
                gm.update(x[s:min(hi, s + 50)])
            gm.fit()
            mean = x.double().mean(0)
            cov = (x.double() - mean).T @ (x.double() - mean) / 600
            assert torch.allclose(gm.mean.double(), mean, atol=1e-10) and float(gm._n_obs) == 600.0
            assert torch.allclose(gm.cov.double(), cov + 1e-8 * torch.eye(12, dtype=torch.double), atol=1e-9)
            # default reduce_on_update=True: per-batch reduce + reduce again in fit -> ratios unchanged
            gm2 = ot.GaussianModel(12, w2_cfg=dict(make_pd=True), dtype=torch.double, device="cpu")
            for s in range(lo, hi, 50):
                gm2.update(x[s:min(hi, s + 50)])
            gm2.fit()
            assert float(gm2._n_obs) == 1200.0 and torch.allclose(gm2.mean.double(), mean, atol=1e-10)
            # ---- (a') GaussianTransport: ONE joint all-reduce of both models' statistics inside compute()
            y = torch.randn(600, 12, generator=torch.Generator().manual_seed(5)) * 0.7 - 0.5
            op = ot.GaussianTransport(12, transport_cfg=dict(make_pd=True), source_cfg=dict(dtype=torch.double, device="cpu",
                                                                                         reduce_on_update=False),
                                      target_cfg=dict(dtype=torch.double, device="cpu", reduce_on_update=False))
            calls = []
            plain = op.source_model.reduce
            counting = lambda t: (calls.append(t.numel()), plain(t))[1]
            op.source_model.reduce = op.target_model.reduce = counting
            lo, hi = parallel.shard_rows(600, rank, world)
            op.update(source_samples=x[lo:hi], target_samples=y[lo:hi])
            w2 = op.compute()
            assert calls == [2 * (1 + 12 + 144)], calls                       # one reduction, both models packed
            want = O.gaussian_transport_pipeline(x, y, 600)
            assert torch.allclose(op.source_model.mean.double(), want["mean_s"], atol=1e-10)
            assert torch.allclose(op.target_model.cov.double(), want["cov_t"], atol=1e-9)
            assert abs(float(w2) - float(want["w2"])) < 1e-8 * float(want["w2"]) and float(op.source_model._n_obs) == 600.0
            # ---- (b) row-sharded Sinkhorn == single-process oracle
            xs = torch.randn(64, 6, generator=g)
            ys = torch.randn(48, 6, generator=g) + 0.5
            a = torch.rand(64, generator=g) + 0.1
            a /= a.sum()
            b = torch.full((48,), 1 / 48)
            lo, hi = parallel.shard_rows(64, rank, world)
            res = parallel.sharded_sinkhorn(xs[lo:hi], ys, a[lo:hi], b, reg=0.05, max_iter=300, threshold=1e-5,
                                            poll_every=4, kernels=fake_kernels)
            C = O.sqeuclidean_cost(xs.double(), ys.double())
            C = C / C.max()
            plan, u, v, iters = O.sinkhorn_log(a.double(), b.double(), C, 0.05, 300, 1e-5, return_potentials=True)
            assert abs(res["scale"] - 1.0 / O.sqeuclidean_cost(xs.double(), ys.double()).max().item()) < 1e-6
            assert iters <= res["iters"] < iters + 4          # stop rule polled every 4 iterations
            u2, v2 = O.sinkhorn_log(a.double(), b.double(), C, 0.05, res["iters"], 0.0, return_potentials=True)[1:3]
            assert torch.allclose(res["u_local"].double(), u2[lo:hi], atol=2e-4)
            assert torch.allclose(res["v"].double(), v2, atol=2e-4)
            # plan statistics of the sharded solution: one packed SUM all-reduce (+ MAX of the row error), no plan in memory
            chk = parallel.sharded_summary(xs[lo:hi], ys, a[lo:hi], b, res["u_local"], res["v"], res["scale"], 0.05,
                                           kernels=fake_kernels)
            pi = torch.exp(u2[:, None] + v2[None, :] - C / 0.05)
            assert abs(chk["cost"] - float((C * pi).sum())) < 1e-4 and abs(chk["mass"] - float(pi.sum())) < 1e-4
            assert abs(chk["max_col_err"] - float((pi.sum(0) - b.double()).abs().max())) < 1e-4 and chk["max_row_err"] < 1e-4
            # the caller-owned plan (buffers [+ graph on CUDA]) can be handed back: same problem, same answer; a plan
            # of another problem is ignored
            xl, al = xs[lo:hi].contiguous(), a[lo:hi].contiguous()
            first = parallel.sharded_sinkhorn(xl, ys, al, b, reg=0.05, max_iter=40, threshold=0.0, kernels=fake_kernels)
            again = parallel.sharded_sinkhorn(xl, ys, al, b, reg=0.05, max_iter=40, threshold=0.0, kernels=fake_kernels,
                                              plan=first["plan"])
            assert again["plan"] is first["plan"]
            assert torch.equal(first["u_local"], again["u_local"]) and torch.equal(first["v"], again["v"])
            other = parallel.sharded_sinkhorn(xl, ys, al, b, reg=0.07, max_iter=5, threshold=0.0, kernels=fake_kernels,
                                              plan=first["plan"])
            assert other["plan"] is not first["plan"]
        open(os.path.join(out_dir, f"ok{rank}"), "w").close()
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_world_size_2_gloo(tmp_path):
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok0") and os.path.exists(tmp_path / "ok1")
