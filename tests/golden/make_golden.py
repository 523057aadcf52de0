"""Generate tests/golden/*.npz by running the UNMODIFIED reference (imported from /root/reference).

Run in the authoring container only:   python tests/golden/make_golden.py
The fixtures are small, committed, and are what pins `oracle/ot_oracle.py` (tests/test_oracle_golden.py)
and - through the oracle and directly - the CUDA path (tests/test_gpu_parity.py).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from _load_reference import load_reference  # noqa: E402

load_reference()
from ot_vae_lightning.ot import matrix_utils as ref_mu  # noqa: E402
from ot_vae_lightning.ot import w2_utils as ref_w2  # noqa: E402
from ot_vae_lightning.ot.distribution_models.codebook_model import CodebookModel  # noqa: E402
from ot_vae_lightning.ot.distribution_models.gaussian_model import GaussianModel  # noqa: E402
from ot_vae_lightning.ot.transport.gaussian_transport import GaussianTransport  # noqa: E402


def npy(t):
    return t.detach().cpu().numpy()


def spd(gen, *shape, kappa=50.0):
    d = shape[-1]
    q, _ = torch.linalg.qr(torch.randn(*shape, d, generator=gen, dtype=torch.double))
    lam = torch.logspace(-np.log10(kappa), 0.0, d, dtype=torch.double)
    return (q * lam) @ q.transpose(-1, -2)


def latents(gen, n, d, lead=(), shift=0.0, scale=1.0):
    mix = torch.randn(*lead, d, d, generator=gen, dtype=torch.double) / np.sqrt(d)
    z = torch.randn(*lead, n, d, generator=gen, dtype=torch.double)
    mu = torch.randn(*lead, 1, d, generator=gen, dtype=torch.double) + shift
    return (scale * (z @ mix.transpose(-1, -2)) + mu).float()


def case_matrix():
    g = torch.Generator().manual_seed(101)
    A = spd(g, 3, 12)
    raw = torch.randn(2, 12, 12, generator=g, dtype=torch.double)
    indef = raw + raw.transpose(-1, -2)  # symmetric, indefinite
    repaired, shift = ref_mu.make_psd(indef, strict=True, return_correction=True)
    s = torch.randn(2, 12, generator=g, dtype=torch.double) * 40
    ss = spd(g, 2, 12) * 500 + s.unsqueeze(-1) * s.unsqueeze(-2) / 30
    n = torch.tensor([30.0, 41.0], dtype=torch.double)
    m, c = ref_mu.mean_cov(s.clone(), ss.clone(), n)
    asym = A.clone()
    asym[0, 1, 2] += 1e-3
    np.savez(os.path.join(HERE, "matrix.npz"),
             A=npy(A), sqrtm=npy(ref_mu.sqrtm(A)), invsqrtm=npy(ref_mu.invsqrtm(A)),
             min_eig=npy(ref_mu.min_eig(A)), indef=npy(indef), indef_min_eig=npy(ref_mu.min_eig(indef)),
             repaired=npy(repaired), shift=npy(shift),
             sum=npy(s), sum_cov=npy(ss), n=npy(n), mean=npy(m), cov=npy(c),
             asym=npy(asym), asym_is_symmetric=npy(ref_mu.is_symmetric(asym)),
             A_is_spd=npy(ref_mu.is_spd(A)), indef_is_pd=npy(ref_mu.is_pd(indef)))


def run_transport(src, tgt, lead, d, batch, decay=None, pg_star=0.0):
    cfg = dict(dtype=torch.double)
    if decay is not None:
        cfg["update_decay"] = decay
    op = GaussianTransport(*lead, d,
                           transport_cfg=dict(diag=False, stochastic=False, make_pd=True, pg_star=pg_star,
                                              dtype=torch.double),
                           source_cfg=dict(cfg), target_cfg=dict(cfg))
    for lo in range(0, src.shape[-2], batch):
        op.update(source_samples=src[..., lo:lo + batch, :])
    for lo in range(0, tgt.shape[-2], batch):
        op.update(target_samples=tgt[..., lo:lo + batch, :])
    w2 = op.compute()
    moved = op.transport(src)
    sm, tm = op.source_model, op.target_model
    return dict(n_s=npy(sm._n_obs), sum_s=npy(sm._running_sum), sumcov_s=npy(sm._running_sum_cov),
                mean_s=npy(sm.mean), cov_s=npy(sm.cov), mean_t=npy(tm.mean), cov_t=npy(tm.cov),
                w2=npy(w2), T=npy(op.transport_operator), Cw=npy(op.cov_stochastic_noise), moved=npy(moved))


def case_gaussian():
    g = torch.Generator().manual_seed(202)
    # single operator, d=16, ragged last batch (650 = 6*100 + 50)
    src = latents(g, 650, 16, shift=0.5)
    tgt = latents(g, 700, 16, shift=-1.0, scale=1.7)
    out = run_transport(src, tgt, (), 16, 100)
    np.savez(os.path.join(HERE, "gaussian_d16.npz"), src=npy(src), tgt=npy(tgt), batch=100, **out)
    # two operators at once (leading shape (2,)), d=8, pg_star blend
    src = latents(g, 300, 8, lead=(2,))
    tgt = latents(g, 300, 8, lead=(2,), shift=2.0, scale=0.6)
    out = run_transport(src, tgt, (2,), 8, 75, pg_star=0.25)
    np.savez(os.path.join(HERE, "gaussian_lead2_d8.npz"), src=npy(src), tgt=npy(tgt), batch=75, pg_star=0.25, **out)
    # EMA accumulation (update_decay), d=8
    src = latents(g, 400, 8)
    tgt = latents(g, 400, 8, shift=1.0)
    out = run_transport(src, tgt, (), 8, 50, decay=0.9)
    np.savez(os.path.join(HERE, "gaussian_ema_d8.npz"), src=npy(src), tgt=npy(tgt), batch=50, decay=0.9, **out)
    # NOTE: the reference's *default* buffer dtype (fp32 buffers + fp64 `_n_obs`) cannot be pinned: `fit()` raises
    # "Index put requires the source and destination dtypes match" at gaussian_model.py:182 (probed here), so every
    # reference test passes dtype=torch.double.  The fixtures above therefore use fp64 buffers.


def case_w2_functions():
    g = torch.Generator().manual_seed(303)
    lead, d = (2, 3), 3  # the reference tests' own shape (tests/test_w2_utils.py:23-24)
    m1 = torch.randn(*lead, d, generator=g)
    m2 = torch.randn(*lead, d, generator=g)
    r1 = torch.randn(*lead, d, d, generator=g)
    r2 = torch.randn(*lead, d, d, generator=g)
    c1 = r1 @ r1.transpose(-1, -2) + torch.eye(d) * 1e-5
    c2 = r2 @ r2.transpose(-1, -2) + torch.eye(d) * 1e-5
    w2 = ref_w2.w2_gaussian(m1, m2, c1, c2)
    T, Cw = ref_w2.compute_transport_operators(c1, c2, stochastic=False, diag=False, make_pd=True)
    x = torch.randn(*lead, 7, d, generator=g)
    y = ref_w2.apply_transport(x, m1.unsqueeze(-2), m2.unsqueeze(-2), T.unsqueeze(-3), Cw.unsqueeze(-3))
    np.savez(os.path.join(HERE, "w2_d3.npz"), m1=npy(m1), m2=npy(m2), c1=npy(c1), c2=npy(c2),
             w2=npy(w2), T=npy(T), x=npy(x), y=npy(y))
    # d = 24, conditioned covariances
    A, B = spd(g, 24, kappa=100.0), spd(g, 24, kappa=30.0) * 2.5
    m1, m2 = torch.randn(24, generator=g, dtype=torch.double), torch.randn(24, generator=g, dtype=torch.double)
    w2 = ref_w2.w2_gaussian(m1, m2, A, B)
    T, _ = ref_w2.compute_transport_operators(A, B, stochastic=False, diag=False)
    np.savez(os.path.join(HERE, "w2_d24.npz"), m1=npy(m1), m2=npy(m2), c1=npy(A), c2=npy(B), w2=npy(w2), T=npy(T))


def case_sinkhorn():
    g = torch.Generator().manual_seed(404)
    # points in d=8, squared-euclidean cost normalised by its max (w2_utils.py:265-266), eps = 0.05
    x = torch.randn(2, 40, 8, generator=g, dtype=torch.double)
    y = torch.randn(2, 56, 8, generator=g, dtype=torch.double) * 1.3 + 0.4
    C = (x * x).sum(-1, keepdim=True) + (y * y).sum(-1).unsqueeze(-2) - 2 * x @ y.transpose(-1, -2)
    C = C / C.amax(dim=(-1, -2), keepdim=True)
    a = torch.rand(2, 40, generator=g, dtype=torch.double); a /= a.sum(-1, keepdim=True)
    b = torch.rand(2, 56, generator=g, dtype=torch.double); b /= b.sum(-1, keepdim=True)
    plan_fixed = ref_w2.sinkhorn_log(a, b, C, reg=0.05, max_iter=25, threshold=0.0)
    plan_conv = ref_w2.sinkhorn_log(a, b, C, reg=0.05, max_iter=1000, threshold=1e-6)
    np.savez(os.path.join(HERE, "sinkhorn_points.npz"), x=npy(x), y=npy(y), a=npy(a), b=npy(b), C=npy(C),
             reg=0.05, plan_fixed25=npy(plan_fixed), plan_thr1e6=npy(plan_conv))
    # the reference test's own setting: 3x3 symmetric costs, reg = 1e-5 (tests/test_w2_utils.py:236-247)
    cost = torch.randn(2, 3, 3, 3, generator=g).abs()
    cost = (cost + cost.transpose(-1, -2)).double()
    a3 = torch.randn(2, 3, 3, generator=g).abs(); a3 = (a3 / a3.sum(-1, keepdim=True)).double()
    b3 = torch.randn(2, 3, 3, generator=g).abs(); b3 = (b3 / b3.sum(-1, keepdim=True)).double()
    plan3 = ref_w2.sinkhorn_log(a3, b3, cost, reg=1e-5, max_iter=1000, threshold=1e-8)
    np.savez(os.path.join(HERE, "sinkhorn_3x3.npz"), a=npy(a3), b=npy(b3), C=npy(cost), plan=npy(plan3))
    # CodebookModel.energy cost (inverse distance), the DiscreteTransport cost producer
    cb = CodebookModel(1, 8, mixture_cfg=dict(n_components=24), dtype=torch.double)
    pts = torch.randn(1, 32, 8, generator=g, dtype=torch.double)
    with torch.no_grad():
        cb.codebook.copy_(torch.randn(1, 24, 8, generator=g, dtype=torch.double))
    np.savez(os.path.join(HERE, "energy.npz"), pts=npy(pts), codebook=npy(cb.codebook), energy=npy(cb.energy(pts)))


def gmm_clouds(g, n, d, centres, spread):
    """n points around len(centres) well separated centres (so that hard assignments do not depend on round-off)"""
    k = len(centres)
    which = torch.arange(n) % k
    mix = torch.randn(k, d, d, generator=g, dtype=torch.double) * spread / np.sqrt(d)
    z = torch.randn(n, d, generator=g, dtype=torch.double)
    pts = torch.einsum("nij,nj->ni", mix[which], z) + torch.tensor(centres, dtype=torch.double)[which]
    return pts[torch.randperm(n, generator=g)].float()


def run_gmm(src, tgt, probe, d, ks, kt, batch, diag, transport_type, source_mode="argmax"):
    from ot_vae_lightning.ot.transport.gmm_transport import GMMTransport
    op = GMMTransport(d, transport_type=transport_type,
                      transport_cfg=dict(diag=diag, stochastic=False, make_pd=True, dtype=torch.double),
                      source_cfg=dict(mixture_cfg=dict(n_components=ks, training_mode=source_mode,
                                                       inference_mode=source_mode), dtype=torch.double),
                      target_cfg=dict(mixture_cfg=dict(n_components=kt), dtype=torch.double))
    # the models draw their initial means with the global generator on their first update: one seed per model
    torch.manual_seed(11)
    for lo in range(0, src.shape[0], batch):
        op.update(source_samples=src[lo:lo + batch])
    torch.manual_seed(12)
    for lo in range(0, tgt.shape[0], batch):
        op.update(target_samples=tgt[lo:lo + batch])
    cost = op.compute()
    torch.manual_seed(13)
    moved = op.transport(probe)
    out = dict(cost=npy(cost), coupling=npy(op.transport_matrix), moved=npy(moved))
    for tag, m in (("s", op.source_model), ("t", op.target_model)):
        out.update({f"n_{tag}": npy(m._n_obs), f"sum_{tag}": npy(m._running_sum), f"sumcov_{tag}": npy(m._running_sum_cov),
                    f"mean_{tag}": npy(m.mean), f"var_{tag}": npy(m.variances), f"w_{tag}": npy(m.weights)})
    out["energy_s"] = npy(op.source_model.energy(probe.double()))
    return out


def case_gmm(only_soft=None):
    """GaussianMixtureModel + GMMTransport on separated clusters, full and diagonal covariances."""
    g = torch.Generator().manual_seed(505)
    d = 6
    cs = [[6.0 * (i == j) - 3.0 * (i == (j + 1) % 3) for j in range(d)] for i in range(3)]
    ct = [[-5.0 if j % 2 == i else 4.0 for j in range(d)] for i in range(2)]
    src, tgt = gmm_clouds(g, 600, d, cs, 0.8), gmm_clouds(g, 500, d, ct, 0.5)
    probe = src[:40].clone()
    out = run_gmm(src, tgt, probe, d, 3, 2, 150, diag=False, transport_type="argmax")
    np.savez(os.path.join(HERE, "gmm_full_argmax.npz"), src=npy(src), tgt=npy(tgt), probe=npy(probe), batch=150, **out)
    # ('barycenter' cannot be pinned: the reference feeds `assignments @ coupling` - rows summing to the source weight,
    # not to 1 - to gaussian_barycenter, whose validation raises "`weights` is expected to be a valid probability
    # vector" (gmm_transport.py:106-111 -> w2_utils.py:648; probed here))
    out = run_gmm(src, tgt, probe, d, 3, 2, 200, diag=True, transport_type="argmax")
    np.savez(os.path.join(HERE, "gmm_diag_argmax.npz"), src=npy(src), tgt=npy(tgt), probe=npy(probe), batch=200, **out)
    if only_soft is not False:
        # soft ('mean') source assignments: weighted statistics with fractional weights, one operator per input
        out = run_gmm(src, tgt, probe[:12], d, 3, 2, 150, diag=False, transport_type="argmax", source_mode="mean")
        np.savez(os.path.join(HERE, "gmm_full_soft.npz"), src=npy(src), tgt=npy(tgt), probe=npy(probe[:12]), batch=150, **out)


def case_barycenter():
    """gaussian_barycenter (w2_utils.py:325-385): closed form for variances, the Alvarez-Esteban fixed point for full
    covariances (the fixed point does not depend on the randomly drawn start, so n_iter = 100 pins it)."""
    g = torch.Generator().manual_seed(808)
    d, n = 8, 4
    mean = torch.randn(2, n, d, generator=g, dtype=torch.double)
    cov = spd(g, 2, n, d, kappa=25.0) * (1.0 + torch.arange(n, dtype=torch.double).view(1, n, 1, 1))
    var = torch.rand(2, n, d, generator=g, dtype=torch.double) + 0.1
    w = torch.rand(2, n, generator=g, dtype=torch.double) + 0.05
    w = w / w.sum(-1, keepdim=True)
    torch.manual_seed(31)
    mb, cb = ref_w2.gaussian_barycenter(mean, cov, w, diag=False, n_iter=100)
    mbd, vbd = ref_w2.gaussian_barycenter(mean, var, w, diag=True)
    np.savez(os.path.join(HERE, "barycenter.npz"), mean=npy(mean), cov=npy(cov), var=npy(var), w=npy(w),
             mean_b=npy(mb), cov_b=npy(cb), mean_b_diag=npy(mbd), var_b_diag=npy(vbd))


def case_discrete():
    """DiscreteTransport end to end: streaming k-means codebooks, inverse-distance cost, Sinkhorn plan,
    argmax / mean routing (discrete_transport.py:27-98, codebook_model.py:122-214)."""
    from ot_vae_lightning.ot.transport.discrete_transport import DiscreteTransport
    g = torch.Generator().manual_seed(707)
    d = 8
    src = gmm_clouds(g, 400, d, [[4.0 * ((i + j) % 3 == 0) for j in range(d)] for i in range(6)], 0.4)
    # (equal codebook sizes: the reference builds the cost as source.energy(target codebook) = [k, n] and hands it to
    # sinkhorn_log with a [n], b [k] - discrete_transport.py:58-66 - so n != k raises in its own logsumexp; probed here)
    tgt = gmm_clouds(g, 400, d, [[-3.0 * ((i + 2 * j) % 4 == 0) + 1.0 for j in range(d)] for i in range(6)], 0.3)
    probe = src[:32].clone()
    out = dict(src=npy(src), tgt=npy(tgt), probe=npy(probe), batch=100)
    for kind in ("argmax", "mean"):
        op = DiscreteTransport(d, transport_type=kind, sinkhorn_reg=0.05, sinkhorn_max_iter=300, sinkhorn_threshold=1e-9,
                               source_cfg=dict(mixture_cfg=dict(n_components=6), dtype=torch.double),
                               target_cfg=dict(mixture_cfg=dict(n_components=6), dtype=torch.double))
        torch.manual_seed(21)                      # the codebooks start from host-side randperm picks: one seed per model
        for lo in range(0, 400, 100):
            op.update(source_samples=src[lo:lo + 100])
        torch.manual_seed(22)
        for lo in range(0, 400, 100):
            op.update(target_samples=tgt[lo:lo + 100])
        cost = op.compute()
        torch.manual_seed(23)
        moved = op.transport(probe)
        out.update({f"cost_{kind}": npy(cost), f"plan_{kind}": npy(op.transport_matrix), f"moved_{kind}": npy(moved)})
    sm, tm = op.source_model, op.target_model
    out.update(codebook_s=npy(sm.codebook), codebook_t=npy(tm.codebook), n_s=npy(sm._n_obs), n_t=npy(tm._n_obs),
               sum_s=npy(sm._running_sum), w_s=npy(sm.weights), w_t=npy(tm.weights))
    np.savez(os.path.join(HERE, "discrete.npz"), **out)


def case_operator_variants():
    """compute_transport_operators beyond the deterministic full-matrix branch: stochastic (eq. 19,
    with a rank-deficient source), diagonal, diagonal stochastic, each with and without a pg_star blend."""
    g = torch.Generator().manual_seed(606)
    d = 10
    cs, ct = spd(g, 2, d, kappa=40.0), spd(g, 2, d, kappa=15.0) * 1.8
    low = torch.randn(2, d, 4, generator=g, dtype=torch.double)
    cs_low = low @ low.transpose(-1, -2)                      # rank 4 source: pinv + stochastic completion
    vs = torch.rand(3, d, generator=g, dtype=torch.double) + 0.2
    vt = torch.rand(3, d, generator=g, dtype=torch.double) * 2 + 0.1
    out = dict(cs=npy(cs), ct=npy(ct), cs_low=npy(cs_low), vs=npy(vs), vt=npy(vt))
    for tag, pg in (("p0", 0.0), ("p3", 0.3)):
        T, Cw = ref_w2.compute_transport_operators(cs.clone(), ct.clone(), stochastic=True, diag=False, pg_star=pg, make_pd=True)
        out[f"full_st_T_{tag}"], out[f"full_st_Cw_{tag}"] = npy(T), npy(Cw)
        # (a rank-deficient source, the case eq. 19 exists for, cannot be pinned: the reference's own `is_spd(Cw)` check
        # dies in `eigh` on the non-finite Cw it produces for cs_low - w2_utils.py:453, probed here)
        T, Cw = ref_w2.compute_transport_operators(vs.clone(), vt.clone(), stochastic=False, diag=True, pg_star=pg)
        out[f"diag_T_{tag}"], out[f"diag_Cw_{tag}"] = npy(T), npy(Cw)
        T, Cw = ref_w2.compute_transport_operators(vs.clone(), vt.clone(), stochastic=True, diag=True, pg_star=pg)
        out[f"diag_st_T_{tag}"], out[f"diag_st_Cw_{tag}"] = npy(T), npy(Cw)
    x = torch.randn(3, 7, d, generator=g, dtype=torch.double)
    ms, mt = torch.randn(3, 1, d, generator=g, dtype=torch.double), torch.randn(3, 1, d, generator=g, dtype=torch.double)
    Td, Cwd = ref_w2.compute_transport_operators(vs.clone(), vt.clone(), stochastic=False, diag=True)
    out.update(x=npy(x), ms=npy(ms), mt=npy(mt),
               y_diag=npy(ref_w2.apply_transport(x, ms, mt, Td.unsqueeze(-2), Cwd.unsqueeze(-2), diag=True)))
    np.savez(os.path.join(HERE, "operator_variants.npz"), **out)


def case_fid():
    """The UNMODIFIED reference `FrechetInceptionDistance` (metrics/fid.py:27-131) behind the torchmetrics stub of
    `_load_reference.py`, fed by the stand-in feature net / images of tests/fid_stub.py.  The 2048 x 2048 correlations are
    too large to commit: the fixture keeps the score, the sums, the counts, the traces and a strided 32 x 32 sample."""
    sys.path.insert(0, os.path.dirname(HERE))
    from fid_stub import CASES, FeatureNet, images
    from ot_vae_lightning.metrics.fid import FrechetInceptionDistance
    out = {}
    for name, (fsize, n_gen, n_smp, batch) in CASES.items():
        fid = FrechetInceptionDistance(net=FeatureNet(fsize), feature_size=fsize)
        gen, smp = images(11, n_gen), images(12, n_smp, gain=0.85, offset=0.05)
        for lo in range(0, n_gen, batch):
            fid.update(generated=gen[lo:lo + batch])
        for lo in range(0, n_smp, batch):
            fid.update(samples=smp[lo:lo + batch])
        step = max(1, fsize // 32)
        out.update({f"{name}_score": npy(fid.compute()), f"{name}_real_sum": npy(fid.real_sum),
                    f"{name}_fake_sum": npy(fid.fake_sum), f"{name}_num_real": npy(fid.num_real_obs),
                    f"{name}_num_fake": npy(fid.num_fake_obs),
                    f"{name}_real_trace": npy(fid.real_correlation.trace()),
                    f"{name}_fake_trace": npy(fid.fake_correlation.trace()),
                    f"{name}_real_corr_sample": npy(fid.real_correlation[::step, ::step]),
                    f"{name}_fake_corr_sample": npy(fid.fake_correlation[::step, ::step])})
        print(name, "FID", float(fid.compute()))
    few = FrechetInceptionDistance(net=FeatureNet(64), feature_size=64)
    few.update(generated=images(13, 999), samples=images(14, 1200))
    out["few_score"] = npy(few.compute())                      # < 1000 observations: +inf (fid.py:126)
    np.savez(os.path.join(HERE, "fid.npz"), **out)


# NOTE (probed here, seed 909, d=10, rank-4 `low @ low.T`): exactly singular 'spsd' arguments cannot be pinned - the
# reference's `sqrtm` returns NaN for them (eigh yields eigenvalues like -1.8e-15 and matrix_utils.py:37-46 takes their
# sqrt), and so does `compute_transport_operators(cs, ct_low, stochastic=False, make_pd=True/False)`.  The GPU tests
# compare against the oracle with those eigenvalues clamped at 0 (tests/test_gpu_parity.py::test_singular_spsd_*).


if __name__ == "__main__":
    torch.manual_seed(0)
    if len(sys.argv) > 1 and sys.argv[1] == "fid":
        case_fid()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "gmm":      # only the fixtures added later (the others stay byte-identical)
        case_gmm()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "barycenter":
        case_barycenter()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "discrete":
        case_discrete()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "operators":
        case_operator_variants()
        sys.exit(0)
    case_matrix()
    case_gaussian()
    case_w2_functions()
    case_sinkhorn()
    case_gmm()
    case_operator_variants()
    case_discrete()
    case_barycenter()
    case_fid()
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))
