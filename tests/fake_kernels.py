"""CPU stand-ins for `ot_vae_lightning_b200.kernels`, backed by the oracle.  TEST INFRASTRUCTURE ONLY: it lets the
`-m "not gpu"` suite exercise the Python host logic (validation, broadcasting, state handling, the distributed
drivers under gloo) in a container without a GPU.  The product never imports this."""
import contextlib

import torch

from oracle import ot_oracle as O


def stats_update(x, n_obs, run_sum, run_cov, decay):
    x = x.to(run_sum.dtype)
    x = x.expand(*run_sum.shape[:-1], *x.shape[-2:])
    n, s, ss = O.batch_stats(x)
    keep, gain = (1.0, 1.0) if decay is None else (decay, 1.0 - decay)
    n_obs.mul_(keep).add_(gain * float(x.shape[-2]))
    run_sum.mul_(keep).add_(gain * s)
    run_cov.mul_(keep).add_(gain * ss)


def mean_cov(run_sum, run_cov, n_obs):
    return O.mean_cov(run_sum, run_cov, n_obs.to(run_sum.dtype).expand(run_sum.shape[:-1]))


def symmetrize_shift(a, shift):
    out = a.triu() + a.triu(1).transpose(-1, -2)
    if shift is not None:
        out = out + torch.eye(a.shape[-1], dtype=a.dtype) * shift[..., None, None]
    return out


def asymmetry(a):
    skew = a.double() - a.double().transpose(-1, -2)
    return (skew * skew).sum(dim=(-1, -2))


def min_eig(a, steps=0):
    low = a.double().tril()
    return O.min_eig(low + low.tril(-1).transpose(-1, -2))


def sqrtm_pair(a, want_root=True, want_iroot=True, ridge=0.0, iters=0):
    a64 = a.double() + ridge * torch.eye(a.shape[-1], dtype=torch.double)
    return (O.sqrtm(a64).to(a.dtype) if want_root else None, O.invsqrtm(a64).to(a.dtype) if want_iroot else None)


def w2_gaussian(ms, mt, cs, ct, iters=0):
    return O.w2_gaussian(ms, mt, cs, ct)


def transport_operator(cs, ct, pg_star=0.0, mean_s=None, mean_t=None, iters=0):
    T, _ = O.transport_operator_full(cs, ct, pg_star)
    w2 = None if mean_s is None else O.w2_gaussian(mean_s, mean_t, cs, ct)
    return T.to(cs.dtype), w2


def transport_operator_stochastic(cs, ct, pg_star=0.0):
    return O.transport_operator_full_stochastic(cs, ct, pg_star)


def apply_transport(x, ms, mt, T):
    return O.apply_transport(x, ms, mt, T).float()


def sinkhorn_dense(a, b, C, reg, max_iter, threshold, want_plan=True, poll_every=16):
    dt = torch.float64 if C.dtype == torch.float64 else torch.float32
    plan, u, v, n = O.sinkhorn_log(a.to(dt), b.to(dt), C.to(dt), reg, max_iter, threshold, return_potentials=True)
    return (plan if want_plan else None), u, v, n


def cost_matrix(x, y, cost, scale=1.0):
    c = O.sqeuclidean_cost(x.double(), y.double()) if cost == 0 else O.inverse_distance_energy(x.double(), y.double())
    return (c * scale).float()


def cost_max(x, y, cost):
    return cost_matrix(x, y, cost).max().reshape(1)


def global_cost_scale(x_local, y):
    import torch.distributed as dist
    mx = cost_max(x_local, y, 0)
    if dist.is_initialized():
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    return 1.0 / float(mx)


def colstep(x_local, y, u_local, scale, reg, cost=0, precision=0, out=None, ws=None, reuse=False):
    t = u_local.double()[:, None] - cost_matrix(x_local, y, cost, scale).double() / reg
    m = t.max(0).values
    cm, cs = m.float(), torch.exp(t - m).sum(0).float()
    if out is not None:
        out[0].copy_(cm)
        out[1].copy_(cs)
        return out[0], out[1]
    return cm, cs


def lse_combine(pm, ps, b, v, diff):
    m = pm.double().max(0).values
    s = (ps.double() * torch.exp(pm.double() - m)).sum(0)
    new = torch.log(b.double() + 1e-8) - (m + torch.log(s))
    if diff is not None:
        diff += (new - v.double()).abs().sum().float()
    v.copy_(new.float())


def rowstep(x_local, y, a_local, v, u_local, diff, scale, reg, cost=0, precision=0, ws=None, reuse=False):
    t = v.double()[None, :] - cost_matrix(x_local, y, cost, scale).double() / reg
    new = torch.log(a_local.double() + 1e-8) - torch.logsumexp(t, dim=1)
    if diff is not None:
        diff += (new - u_local.double()).abs().sum().float()
    u_local.copy_(new.float())


def points_summary(x_local, y, a_local, b, u_local, v, scale, reg, cost=0, precision=0, ws=None, reuse=False):
    C = cost_matrix(x_local, y, cost, scale).double()
    pi = torch.exp(u_local.double()[:, None] + v.double()[None, :] - C / reg)
    rows, cols = pi.sum(1), pi.sum(0)
    part = torch.stack([(C * pi).sum(), pi.sum(), (rows - a_local.double()).abs().max(), (cols - b.double()).abs().max()])
    return part, rows.float(), cols.float()


NAMES = ["stats_update", "mean_cov", "symmetrize_shift", "asymmetry", "min_eig", "sqrtm_pair", "w2_gaussian",
         "transport_operator", "transport_operator_stochastic", "apply_transport", "sinkhorn_dense", "cost_matrix", "cost_max", "colstep", "lse_combine",
         "rowstep", "points_summary"]


@contextlib.contextmanager
def installed():
    """Swap the kernel wrappers for the CPU stand-ins (and lift the CUDA-buffer requirement) inside the block."""
    from ot_vae_lightning_b200 import kernels as K
    from ot_vae_lightning_b200.ot.distribution_models import gaussian_model as gm
    saved = {n: getattr(K, n) for n in NAMES}
    saved_req = gm.GaussianModel._require_cuda_buffers
    for n in NAMES:
        setattr(K, n, globals()[n])
    gm.GaussianModel._require_cuda_buffers = lambda self: None
    try:
        yield
    finally:
        for n, f in saved.items():
            setattr(K, n, f)
        gm.GaussianModel._require_cuda_buffers = saved_req
