"""Parity of the CUDA path (through the Python mirror -> C ABI -> sm_100a kernels) against
 (1) the committed golden outputs of the unmodified reference (tests/golden/*.npz),
 (2) the CPU oracle on seeded synthetic inputs,
 (3) size-independent properties at larger sizes.
Tolerances are BASELINE.json's: rel <= 1e-4 mean/cov, <= 1e-3 sqrtm / W2 / transported latents,
<= 1e-4 Sinkhorn marginals and cost (relative Frobenius unless stated)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

TOL_STATS, TOL_MATFUN, TOL_SINKHORN = 1e-4, 1e-3, 1e-4


def T(a, dev="cuda"):
    return torch.from_numpy(np.asarray(a)).to(dev)


def rel(got, want):
    got, want = got.detach().double().cpu(), torch.as_tensor(np.asarray(want)).double()
    return ((got - want).norm() / want.norm().clamp_min(1e-300)).item()


@pytest.fixture(scope="module")
def api():
    import ot_vae_lightning_b200.ot as ot
    return ot


@pytest.fixture(scope="module")
def oracle():
    from oracle import ot_oracle
    return ot_oracle


# ------------------------------------------------------------------------------------------------- golden fixtures

def test_golden_matrix_primitives(api, golden):
    g = golden("matrix")
    A = T(g["A"])
    assert rel(api.sqrtm(A), g["sqrtm"]) < TOL_MATFUN
    assert rel(api.invsqrtm(A), g["invsqrtm"]) < TOL_MATFUN
    assert api.sqrtm(A).dtype == torch.float64 and api.sqrtm(A.float()).dtype == torch.float32
    assert np.allclose(api.min_eig(A).cpu().numpy(), g["min_eig"], rtol=1e-6, atol=1e-9)
    assert np.allclose(api.min_eig(T(g["indef"])).cpu().numpy(), g["indef_min_eig"], rtol=1e-6)
    fixed, shift = api.make_psd(T(g["indef"]), strict=True, return_correction=True)
    assert rel(fixed, g["repaired"]) < 1e-6 and rel(shift, g["shift"]) < 1e-6
    mean, cov = api.mean_cov(T(g["sum"]), T(g["sum_cov"]), T(g["n"]))
    assert rel(mean, g["mean"]) < 1e-12 and rel(cov, g["cov"]) < 1e-12
    assert api.is_symmetric(T(g["asym"])).cpu().tolist() == g["asym_is_symmetric"].tolist()
    assert api.is_spd(A).cpu().tolist() == g["A_is_spd"].tolist()
    assert api.is_pd(T(g["indef"])).cpu().tolist() == g["indef_is_pd"].tolist()


def _run_transport(api, g, lead, d, decay=None, pg_star=0.0):
    cfg = dict(dtype=torch.double)
    if decay is not None:
        cfg["update_decay"] = decay
    op = api.GaussianTransport(*lead, d, transport_cfg=dict(diag=False, stochastic=False, make_pd=True,
                                                            pg_star=pg_star, dtype=torch.double),
                               source_cfg=dict(cfg), target_cfg=dict(cfg)).cuda()
    src, tgt, bs = T(g["src"]), T(g["tgt"]), int(g["batch"])
    for lo in range(0, src.shape[-2], bs):
        op.update(source_samples=src[..., lo:lo + bs, :])
    for lo in range(0, tgt.shape[-2], bs):
        op.update(target_samples=tgt[..., lo:lo + bs, :])
    w2 = op.compute()
    moved = op.transport(src)
    sm, tm = op.source_model, op.target_model
    assert rel(sm._n_obs, g["n_s"]) < 1e-12
    assert rel(sm._running_sum, g["sum_s"]) < TOL_STATS and rel(sm._running_sum_cov, g["sumcov_s"]) < TOL_STATS
    assert rel(sm.mean, g["mean_s"]) < TOL_STATS and rel(tm.mean, g["mean_t"]) < TOL_STATS
    assert rel(sm.cov, g["cov_s"]) < TOL_STATS and rel(tm.cov, g["cov_t"]) < TOL_STATS
    assert rel(w2, g["w2"]) < TOL_MATFUN
    assert rel(op.transport_operator, g["T"]) < TOL_MATFUN
    assert moved.dtype == src.dtype and moved.device == src.device and rel(moved, g["moved"]) < TOL_MATFUN
    assert float(op.cov_stochastic_noise.abs().max()) == 0.0
    return op


def test_golden_gaussian_transport_d16_ragged_batches(api, golden):
    op = _run_transport(api, golden("gaussian_d16"), (), 16)
    assert sorted(op.source_model.state_dict().keys()) == sorted(
        ["mean", "vec_init", "mat_init", "cov_init", "_running_sum", "_running_sum_cov", "_n_obs",
         "parametrizations.cov.original"])
    assert all(v.dtype == torch.float64 for v in op.source_model.state_dict().values())


def test_golden_gaussian_transport_leading_dims_pgstar(api, golden):
    g = golden("gaussian_lead2_d8")
    _run_transport(api, g, (2,), 8, pg_star=float(g["pg_star"]))


def test_golden_gaussian_transport_ema(api, golden):
    g = golden("gaussian_ema_d8")
    _run_transport(api, g, (), 8, decay=float(g["decay"]))


def test_golden_w2_functions(api, golden):
    for name in ("w2_d3", "w2_d24"):
        g = golden(name)
        m1, m2, c1, c2 = (T(g[k]) for k in ("m1", "m2", "c1", "c2"))
        w2 = api.w2_gaussian(m1, m2, c1, c2)
        assert w2.dtype == torch.float64 and w2.shape == torch.Size(g["w2"].shape)
        assert rel(w2, g["w2"]) < TOL_MATFUN
        Top, Cw = api.compute_transport_operators(c1, c2, stochastic=False, diag=False, make_pd=True)
        assert rel(Top, g["T"]) < TOL_MATFUN and float(Cw.abs().max()) == 0.0
        if "x" in g:
            y = api.apply_transport(T(g["x"]), m1.unsqueeze(-2), m2.unsqueeze(-2), Top.unsqueeze(-3), Cw.unsqueeze(-3))
            assert y.dtype == torch.float64 and rel(y, g["y"]) < TOL_MATFUN


def test_golden_sinkhorn(api, golden):
    g = golden("sinkhorn_points")
    a, b, C = T(g["a"]), T(g["b"]), T(g["C"])
    for key, kw in (("plan_fixed25", dict(max_iter=25, threshold=0.0)), ("plan_thr1e6", dict(max_iter=1000, threshold=1e-6))):
        plan = api.sinkhorn_log(a, b, C, reg=float(g["reg"]), **kw)            # fp64 in -> fp64 solve
        assert plan.dtype == torch.float64 and rel(plan, g[key]) < 1e-9
        plan32 = api.sinkhorn_log(a.float(), b.float(), C.float(), reg=float(g["reg"]), **kw)
        want = torch.from_numpy(g[key])
        assert rel(plan32.sum(-1), want.sum(-1)) < TOL_SINKHORN and rel(plan32.sum(-2), want.sum(-2)) < TOL_SINKHORN
        assert rel((plan32.double().cpu() * C.cpu()).sum((-1, -2)), (want * C.cpu()).sum((-1, -2))) < TOL_SINKHORN
    g3 = golden("sinkhorn_3x3")  # the reference test's own case: reg = 1e-5, allclose(rtol 1e-5, atol 1e-8)
    plan = api.sinkhorn_log(T(g3["a"]), T(g3["b"]), T(g3["C"]), reg=1e-5, max_iter=1000, threshold=1e-8)
    assert torch.allclose(plan.cpu(), torch.from_numpy(g3["plan"]), rtol=1e-5, atol=1e-8)


def test_golden_energy_cost(golden):
    from ot_vae_lightning_b200 import kernels as K
    g = golden("energy")
    got = K.cost_matrix(T(g["pts"][0]).float(), T(g["codebook"][0]).float(), cost=1)
    assert rel(got, g["energy"][0]) < 1e-5


# ------------------------------------------------------------------------------------------------- oracle, synthetic

@pytest.mark.parametrize("d,n,bs", [(64, 10000, 100), (128, 10000, 250), (256, 20000, 1000), (512, 65536, 8192)])
def test_streaming_stats_vs_oracle(api, oracle, d, n, bs):
    """the reference's test_empirical_cov protocol (tests/test_empirical_cov.py:47-72): batched accumulation then
    mean_cov must equal the one-shot mean / biased covariance."""
    from ot_vae_lightning_b200.synthetic import gaussian_latents
    x = gaussian_latents(n, d, seed=10 + d, device="cuda")
    gm = api.GaussianModel(d, w2_cfg=dict(make_pd=True), dtype=torch.double).cuda()
    for lo in range(0, n, bs):
        gm.update(x[lo:lo + bs])
    gm.fit()
    x64 = x.double().cpu()
    mean = x64.mean(0)
    cov = (x64 - mean).T @ (x64 - mean) / n
    assert rel(gm.mean, mean) < TOL_STATS and rel(gm.cov, cov) < TOL_STATS
    st = oracle.GaussianStats(d)
    st.update(x64)
    assert rel(gm._running_sum_cov, st.sum_cov) < 1e-5 and rel(gm._running_sum, st.sum) < 1e-5


@pytest.mark.parametrize("d,kappa", [(64, 1e2), (128, 1e2), (128, 1e4), (512, 1e2), (1024, 1e2)])
def test_sqrtm_w2_operator_vs_oracle(api, oracle, d, kappa):
    from ot_vae_lightning_b200.synthetic import gaussian_spec
    _, hs = gaussian_spec(d, seed=3, kappa=kappa)
    _, ht = gaussian_spec(d, seed=4, kappa=kappa)
    cs, ct = hs @ hs.T, 1.7 * (ht @ ht.T)
    ms, mt = torch.randn(d, dtype=torch.double), torch.randn(d, dtype=torch.double)
    root = api.sqrtm(cs.cuda())
    assert rel(root, oracle.sqrtm(cs)) < TOL_MATFUN and rel(api.invsqrtm(cs.cuda()), oracle.invsqrtm(cs)) < TOL_MATFUN
    assert rel(root @ root, cs) < TOL_MATFUN
    assert rel(api.w2_gaussian(ms.cuda(), mt.cuda(), cs.cuda(), ct.cuda()), oracle.w2_gaussian(ms, mt, cs, ct)) < TOL_MATFUN
    Top, _ = api.compute_transport_operators(cs.cuda(), ct.cuda(), stochastic=False, diag=False)
    want, _ = oracle.transport_operator_full(cs, ct)
    assert rel(Top, want) < TOL_MATFUN
    # defining property of the Monge map: T Cs T = Ct
    assert rel(Top @ cs.cuda() @ Top, ct) < 5 * TOL_MATFUN


def test_conditional_batch_of_operators(api, oracle):
    """cfg4 shape: one operator per class (leading shape (10,)), d reduced to keep the oracle fast."""
    from ot_vae_lightning_b200.synthetic import gaussian_latents
    L, d, n = 10, 96, 1024
    src = torch.stack([gaussian_latents(n, d, seed=20 + k, device="cuda") for k in range(L)])
    tgt = torch.stack([gaussian_latents(n, d, seed=40 + k, device="cuda", shift=1.0, scale=0.8) for k in range(L)])
    op = api.GaussianTransport(L, d, transport_cfg=dict(make_pd=True), source_cfg=dict(dtype=torch.double),
                               target_cfg=dict(dtype=torch.double)).cuda()
    for lo in range(0, n, 256):
        op.update(src[:, lo:lo + 256], tgt[:, lo:lo + 256])
    w2 = op.compute()
    moved = op.transport(src)
    want = oracle.gaussian_transport_pipeline(src.cpu(), tgt.cpu(), 256)
    assert w2.shape == (L,) and rel(w2, want["w2"]) < TOL_MATFUN
    assert rel(op.transport_operator, want["T"]) < TOL_MATFUN and rel(moved, want["moved"]) < TOL_MATFUN


@pytest.mark.parametrize("n,m,d", [(256, 192, 16), (1024, 1024, 128), (2048, 1536, 128), (1000, 1333, 72)])
def test_sinkhorn_points_vs_oracle(oracle, n, m, d):
    from ot_vae_lightning_b200 import kernels as K
    from ot_vae_lightning_b200.synthetic import point_clouds
    x, y = point_clouds(n, m, d, seed=5, device="cuda")
    a = torch.rand(n, device="cuda") + 0.5
    a /= a.sum()
    b = torch.full((m,), 1.0 / m, device="cuda")
    res = K.sinkhorn_points(x, y, a, b, reg=0.05, max_iter=40, threshold=0.0)
    C = oracle.sqeuclidean_cost(x.double().cpu(), y.double().cpu())
    scale = 1.0 / C.max()
    plan, u, v, _ = oracle.sinkhorn_log(a.double().cpu(), b.double().cpu(), C * scale, reg=0.05, max_iter=40,
                                        threshold=0.0, return_potentials=True)
    s = res["summary"].cpu()
    assert abs(s[0].item() - (C * scale * plan).sum().item()) / (C * scale * plan).sum().item() < TOL_SINKHORN
    assert abs(s[1].item() - plan.sum().item()) < TOL_SINKHORN
    # marginals of the plan the kernels hold (never materialised) against the reference plan's marginals
    assert rel(res["row_marginal"], plan.sum(1)) < TOL_SINKHORN and rel(res["col_marginal"], plan.sum(0)) < TOL_SINKHORN
    # the potentials themselves, rebuilt on the oracle's exact fp64 cost: the TF32 rounding of the points (2^-12
    # relative) perturbs individual plan entries by ~1e-3, it must average out of the marginals to ~1e-3
    got = torch.exp(res["u"].double().cpu()[:, None] + res["v"].double().cpu()[None, :] - C * scale / 0.05)
    assert rel(got.sum(1), plan.sum(1)) < 2e-3 and rel(got.sum(0), plan.sum(0)) < 2e-3
    assert res["iters"] == 40
    # `precision = 1` (exact fp32 cost tiles, what DiscreteTransport asks for): the same reconstruction on the exact cost now
    # holds at the Sinkhorn tolerance, and so does the plan `otk_sinkhorn_points_plan` materialises
    ex = K.sinkhorn_points(x, y, a, b, reg=0.05, max_iter=40, threshold=0.0, precision=1)
    got = torch.exp(ex["u"].double().cpu()[:, None] + ex["v"].double().cpu()[None, :] - C * scale / 0.05)
    assert rel(got.sum(1), plan.sum(1)) < TOL_SINKHORN and rel(got.sum(0), plan.sum(0)) < TOL_SINKHORN
    native = K.points_plan(x, y, ex["u"], ex["v"], float(scale), 0.05)
    assert rel(native.sum(1), plan.sum(1)) < TOL_SINKHORN and rel(native.sum(0), plan.sum(0)) < TOL_SINKHORN
    assert rel(native, plan) < 1e-3


def test_sinkhorn_stop_rule_matches_oracle(api, oracle):
    """the MIN-over-batch stop rule (reference w2_utils.py:314-315): same iteration count as the oracle.
    N == M on purpose: with N != M the `+1e-8` inside the logs makes the two marginals' masses differ, the potentials
    drift by a constant every iteration and sum|du|+sum|dv| never drops below ~(N+M)|N-M|1e-8 (reference quirk)."""
    from ot_vae_lightning_b200 import kernels as K
    g = torch.Generator().manual_seed(9)
    C = torch.rand(3, 50, 50, generator=g, dtype=torch.double)
    C[1] *= 0.3                                                     # batch element 1 converges first and stops all
    a = torch.full((3, 50), 1 / 50, dtype=torch.double)
    b = torch.rand(3, 50, generator=g, dtype=torch.double)
    b /= b.sum(-1, keepdim=True)
    for thr, poll in ((1e-5, 1), (1e-7, 3), (1e-9, 16)):
        plan, u, v, iters = oracle.sinkhorn_log(a, b, C, reg=0.1, max_iter=500, threshold=thr, return_potentials=True)
        got_plan, gu, gv, got_iters = K.sinkhorn_dense(a.cuda(), b.cuda(), C.cuda(), 0.1, 500, thr, poll_every=poll)
        assert got_iters == iters and iters < 500, (got_iters, iters)
        assert rel(got_plan, plan) < 1e-9


# ------------------------------------------------------------------------------------------------- properties / edges

def test_stats_linearity_and_empty_batch(api):
    from ot_vae_lightning_b200.synthetic import gaussian_latents
    d, n = 384, 200_000
    x = gaussian_latents(n, d, seed=77, device="cuda")
    whole = api.GaussianModel(d, dtype=torch.double).cuda()
    whole.update(x)
    parts = api.GaussianModel(d, dtype=torch.double).cuda()
    parts.update(x[:1]); parts.update(x[1:70_001]); parts.update(x[70_001:70_001]); parts.update(x[70_001:])
    assert float(parts._n_obs) == n == float(whole._n_obs)
    assert rel(parts._running_sum_cov, whole._running_sum_cov.cpu()) < 1e-6
    assert rel(parts._running_sum, whole._running_sum.cpu()) < 1e-6
    sc = whole._running_sum_cov
    assert float((sc - sc.T).abs().max()) == 0.0  # exactly symmetric


def test_transport_round_trip_at_scale(api):
    """encode -> decode property at a size the oracle cannot reach: T_{t->s}(T_{s->t}(x)) = x."""
    from ot_vae_lightning_b200.synthetic import gaussian_latents
    d, n = 256, 500_000
    src = gaussian_latents(n, d, seed=5, device="cuda")
    tgt = gaussian_latents(n, d, seed=6, device="cuda", shift=-0.3, scale=2.0)
    fwd = api.GaussianTransport(d, transport_cfg=dict(make_pd=True), source_cfg=dict(dtype=torch.double),
                                target_cfg=dict(dtype=torch.double)).cuda()
    bwd = api.GaussianTransport(d, transport_cfg=dict(make_pd=True), source_cfg=dict(dtype=torch.double),
                                target_cfg=dict(dtype=torch.double)).cuda()
    for lo in range(0, n, 65536):
        fwd.update(src[lo:lo + 65536], tgt[lo:lo + 65536])
        bwd.update(tgt[lo:lo + 65536], src[lo:lo + 65536])
    w_f, w_b = fwd.compute(), bwd.compute()
    assert abs(float(w_f) - float(w_b)) / float(w_f) < TOL_MATFUN          # W2 is symmetric
    back = bwd.transport(fwd.transport(src))
    assert ((back - src).norm() / src.norm()).item() < TOL_MATFUN
    moved = fwd.transport(src[:200_000]).double()
    mean = moved.mean(0)
    cov = (moved - mean).T @ (moved - mean) / moved.shape[0]
    assert rel(mean, fwd.target_model.mean.cpu()) < 1e-2 and rel(cov, fwd.target_model.cov.cpu()) < 2e-2


def test_small_and_odd_dims(api, oracle):
    """d = 3 is the reference test-suite's own dimension (tests/test_w2_utils.py:24); 130 is not a multiple of 4."""
    for d in (3, 5, 130):
        g = torch.Generator().manual_seed(d)
        r = torch.randn(2, d, d, generator=g, dtype=torch.double)
        c1 = r @ r.transpose(-1, -2) + 0.1 * torch.eye(d, dtype=torch.double)
        r = torch.randn(2, d, d, generator=g, dtype=torch.double)
        c2 = r @ r.transpose(-1, -2) + 0.1 * torch.eye(d, dtype=torch.double)
        m1, m2 = torch.randn(2, d, generator=g, dtype=torch.double), torch.randn(2, d, generator=g, dtype=torch.double)
        assert rel(api.w2_gaussian(m1.cuda(), m2.cuda(), c1.cuda(), c2.cuda()), oracle.w2_gaussian(m1, m2, c1, c2)) < TOL_MATFUN
        x = torch.randn(2, 33, d, generator=g)
        Top, Cw = api.compute_transport_operators(c1.cuda(), c2.cuda(), stochastic=False, diag=False)
        y = api.apply_transport(x.cuda(), m1.cuda().unsqueeze(-2), m2.cuda().unsqueeze(-2), Top.unsqueeze(-3), Cw.unsqueeze(-3))
        want = oracle.apply_transport(x, m1, m2, oracle.transport_operator_full(c1, c2)[0])
        assert rel(y, want) < TOL_MATFUN


def test_w2_self_distance_is_small(api):
    """reference tests/test_w2_utils.py:35-41 asks |W2(x,x)| <= 1e-8 d in fp64; the fp32-accurate kernels give
    ~1e-6 relative to tr(C) (documented deviation)."""
    g = torch.Generator().manual_seed(1)
    r = torch.randn(2, 3, 8, 8, generator=g)
    cov = (r @ r.transpose(-1, -2) + 1e-5 * torch.eye(8)).cuda()
    mean = torch.randn(2, 3, 8, generator=g).cuda()
    w = api.w2_gaussian(mean, mean, cov, cov)
    assert w.shape == (2, 3)
    tr = cov.diagonal(dim1=-1, dim2=-2).sum(-1).double()
    assert float((w.abs() / tr).max()) < 1e-5


def test_validation_errors_match_reference_conditions(api):
    eye = torch.eye(4, dtype=torch.double, device="cuda")
    v = torch.zeros(4, dtype=torch.double, device="cuda")
    with pytest.raises(ValueError):
        api.w2_gaussian(v, v, eye, [[1.0]])                                  # not a tensor
    with pytest.raises(ValueError):
        api.w2_gaussian(v, v, eye, -eye)                                     # not PD, make_pd=False
    with pytest.raises(ValueError):
        api.w2_gaussian(v, v, eye, eye + torch.triu(torch.ones_like(eye), 1))  # asymmetric
    with pytest.raises(ValueError):
        api.w2_gaussian(v, v[:3], eye, eye)                                  # dims mismatch
    with pytest.raises(ValueError):
        api.apply_transport(v, v, v, eye, None)                              # Cw=None is rejected (reference :619)
    assert float(api.w2_gaussian(v, v, eye, -eye, make_pd=True)) > 0        # repaired instead
    with pytest.raises((ValueError, AttributeError, TypeError)):
        api.mean_cov(v, eye, 3)                                              # python int crashes the reference too


def test_compute_falls_back_to_eigenvalue_repair_for_indefinite_covariance(api, oracle, golden):
    """`compute()` skips the smallest-eigenvalue solve when Newton-Schulz certifies a PD covariance; an indefinite
    matrix written into `.cov` must still get the reference's make_psd(strict) repair (gaussian_model.py:204-217)."""
    d = 12
    indef = T(golden("matrix")["indef"][0])                      # symmetric, lambda_min ~ -10
    op = api.GaussianTransport(d, transport_cfg=dict(make_pd=True), source_cfg=dict(dtype=torch.double),
                               target_cfg=dict(dtype=torch.double)).cuda()
    g = torch.Generator().manual_seed(3)
    x = (torch.randn(500, d, generator=g) * 1.5 + 0.3).cuda()
    op.update(target_samples=x)
    op.target_model.fit()
    with torch.no_grad():
        op.source_model.cov = indef.clone()
        op.source_model.mean.copy_(torch.zeros(d, dtype=torch.double))
    op.fit_models = lambda: None                                  # keep the hand-written source statistics
    w2 = op.compute()
    cov_s = oracle.parametrized_cov(indef.cpu())
    cov_t = op.target_model.cov.cpu()
    want_w2 = oracle.w2_gaussian(torch.zeros(d, dtype=torch.double), op.target_model.mean.cpu(), cov_s, cov_t)
    # the repaired matrix has lambda_min = 1e-8 exactly, so T itself is ill-posed (it depends on that eigenvalue to
    # 1e-9); the distance and the repaired covariance are the robust observables
    assert rel(w2, want_w2) < TOL_MATFUN and bool(torch.isfinite(op.transport_operator).all())
    assert float((op.source_model.cov.cpu() - cov_s).abs().max()) < 1e-4


def test_streaming_helpers_match_direct_calls(api):
    """host-buffer pipelines (streaming.py: H2D | kernels | D2H on side streams) == plain device-resident calls"""
    from ot_vae_lightning_b200.streaming import stream_transport, stream_update
    from ot_vae_lightning_b200.synthetic import gaussian_latents
    d, n, chunk = 128, 50_000, 8192
    src = gaussian_latents(n, d, seed=31, device="cuda")
    tgt = gaussian_latents(n, d, seed=32, device="cuda", shift=1.0, scale=0.7)
    mk = lambda: api.GaussianTransport(d, transport_cfg=dict(make_pd=True), source_cfg=dict(dtype=torch.double),
                                       target_cfg=dict(dtype=torch.double)).cuda()
    direct, piped = mk(), mk()
    for lo in range(0, n, chunk):
        direct.update(src[lo:lo + chunk], tgt[lo:lo + chunk])
    h_src, h_tgt = src.cpu().pin_memory(), tgt.cpu().pin_memory()
    stream_update(piped, h_src, h_tgt, chunk)
    assert rel(piped.source_model._running_sum_cov, direct.source_model._running_sum_cov.cpu()) < 1e-12
    assert rel(piped.target_model._running_sum, direct.target_model._running_sum.cpu()) < 1e-12
    w_d, w_p = direct.compute(), piped.compute()
    assert abs(float(w_d) - float(w_p)) < 1e-9 * abs(float(w_d))
    h_out = torch.empty_like(h_src).pin_memory()
    stream_transport(piped, h_src, h_out, chunk)
    torch.cuda.synchronize()
    want = torch.cat([direct.transport(src[lo:lo + chunk]) for lo in range(0, n, chunk)]).cpu()
    assert torch.equal(h_out, want)


def test_row_sharded_halfsteps_match_fused_solver(api):
    """The row-sharded half-steps (two emulated ranks on one GPU, operands prepared once and reused from a dedicated
    workspace) reproduce the potentials of the single-GPU fused solver."""
    from ot_vae_lightning_b200 import kernels as K
    from ot_vae_lightning_b200.synthetic import point_clouds
    dev = torch.device("cuda", 0)
    n, m, d, reg, iters = 1536, 1024, 128, 0.05, 12
    x, y = point_clouds(n, m, d, seed=21, device=dev)
    a = torch.full((n,), 1.0 / n, device=dev)
    b = torch.full((m,), 1.0 / m, device=dev)
    scale = 1.0 / float(K.cost_max(x, y, 0).item())
    want = K.sinkhorn_points(x, y, a, b, reg=reg, max_iter=iters, threshold=0.0, scale=scale)
    shards = [(0, 640), (640, n)]
    us = [torch.zeros(hi - lo, device=dev) for lo, hi in shards]
    wss = [K.points_workspace(hi - lo, m, d, 0, dev) for lo, hi in shards]
    v = torch.zeros(m, device=dev)
    parts = torch.empty(len(shards), 2, m, device=dev)
    for it in range(iters):
        for r, (lo, hi) in enumerate(shards):
            K.colstep(x[lo:hi], y, us[r], scale, reg, out=parts[r], ws=wss[r], reuse=it > 0)
        K.lse_combine(parts[:, 0], parts[:, 1], b, v, None)
        for r, (lo, hi) in enumerate(shards):
            K.rowstep(x[lo:hi], y, a[lo:hi], v, us[r], None, scale, reg, ws=wss[r], reuse=True)
    u = torch.cat(us)
    # every shard rounds its points to FP16 with its own sigma = max |coordinate|, so potentials agree to the
    # operand-rounding level, not to fp32 round-off
    assert (u - want["u"]).abs().max().item() < 2e-3
    assert (v - want["v"]).abs().max().item() < 2e-3
    assert (u - want["u"]).abs().mean().item() < 2e-4


def test_sharded_sinkhorn_driver_graph_replay_matches_fused_solver(api):
    """`parallel.sharded_sinkhorn` (world size 1 here: eager first iterations, then CUDA-graph replays) against the
    fused single-GPU solver, with and without the graph."""
    from ot_vae_lightning_b200 import kernels as K, parallel
    from ot_vae_lightning_b200.synthetic import point_clouds
    dev = torch.device("cuda", 0)
    n, m, d, reg, iters = 1024, 1280, 64, 0.05, 15
    x, y = point_clouds(n, m, d, seed=5, device=dev)
    a = torch.full((n,), 1.0 / n, device=dev)
    b = torch.full((m,), 1.0 / m, device=dev)
    scale = 1.0 / float(K.cost_max(x, y, 0).item())
    want = K.sinkhorn_points(x, y, a, b, reg=reg, max_iter=iters, threshold=0.0, scale=scale)
    for use_graph in (False, True):
        got = parallel.sharded_sinkhorn(x, y, a, b, reg=reg, max_iter=iters, threshold=0.0, scale=scale, use_graph=use_graph)
        assert got["iters"] == iters
        assert (got["u_local"] - want["u"]).abs().max().item() < 1e-4
        assert (got["v"] - want["v"]).abs().max().item() < 1e-4


def test_prepared_transport_matches_functional_path_and_tracks_operator_changes(api):
    """`GaussianTransport.transport` applies a prepared operator (built once per map): same result as the functional
    `apply_transport`, rebuilt when the map or a mean is replaced, exact fallback for values outside the FP16 window."""
    from ot_vae_lightning_b200 import kernels as K
    from ot_vae_lightning_b200.synthetic import gaussian_latents
    dev = torch.device("cuda", 0)
    d, n = 192, 6000
    src = gaussian_latents(n, d, seed=3, device=dev)
    tgt = gaussian_latents(n, d, seed=4, device=dev, shift=0.5, scale=1.5)
    cfg = dict(dtype=torch.double, device=dev)
    op = api.GaussianTransport(d, transport_cfg=dict(make_pd=True), source_cfg=dict(cfg), target_cfg=dict(cfg)).to(dev)
    op.update(source_samples=src, target_samples=tgt)
    op.compute()
    want = (src.double() - op.source_model.mean) @ op.transport_operator.T + op.target_model.mean
    got = op.transport(src)
    assert got.dtype == src.dtype and rel(got, want.cpu().numpy()) < 1e-5
    assert rel(K.apply_transport(src, op.source_model.mean, op.target_model.mean, op.transport_operator), want.cpu().numpy()) < 1e-5
    # a latent far outside the FP16 window of the prepared scales: the device-gated TF32 kernel recomputes
    far = src.clone()
    far[17, 5] = 4.0e7
    want_far = (far.double() - op.source_model.mean) @ op.transport_operator.T + op.target_model.mean
    assert rel(op.transport(far), want_far.cpu().numpy()) < 1e-5
    # replacing the map must not reuse the stale preparation
    op.transport_operator = 2.0 * op.transport_operator
    assert rel(op.transport(src), (2.0 * (want - op.target_model.mean) + op.target_model.mean).cpu().numpy()) < 1e-5
    # ragged batch, leading dims and a second compute()
    op.update(source_samples=tgt[:1000], target_samples=src[:1000])
    op.compute()
    want2 = (src[:777].double() - op.source_model.mean) @ op.transport_operator.T + op.target_model.mean
    assert rel(op.transport(src[:777]), want2.cpu().numpy()) < 1e-5


# ------------------------------------------------------------------------------------------------- GMM

def _gmm_on_gpu(api, g, diag, n_rep=1):
    d, bs = int(g["src"].shape[1]), int(g["batch"])
    cfg = dict(dtype=torch.double, device="cuda")
    op = api.GMMTransport(d, transport_type="argmax",
                          transport_cfg=dict(diag=diag, stochastic=False, make_pd=True, dtype=torch.double),
                          source_cfg=dict(mixture_cfg=dict(n_components=int(g["n_s"].shape[0])), **cfg),
                          target_cfg=dict(mixture_cfg=dict(n_components=int(g["n_t"].shape[0])), **cfg)).cuda()
    src, tgt = T(g["src"]), T(g["tgt"])
    torch.manual_seed(11)                       # the seeds of tests/golden/make_golden.py::run_gmm (host-side randperm)
    for lo in range(0, src.shape[0], bs):
        op.update(source_samples=src[lo:lo + bs])
    torch.manual_seed(12)
    for lo in range(0, tgt.shape[0], bs):
        op.update(target_samples=tgt[lo:lo + bs])
    return op


@pytest.mark.parametrize("name,diag", [("gmm_full_argmax", False), ("gmm_diag_argmax", True)])
def test_golden_gmm_transport(api, golden, name, diag):
    """GaussianMixtureModel streaming fit (per-component weighted SYRK through otk_stats_update), component OT through
    the Sinkhorn kernel and the per-pair Gaussian maps, against the unmodified reference's outputs."""
    g = golden(name)
    op = _gmm_on_gpu(api, g, diag)
    cost = op.compute()
    moved = op.transport(T(g["probe"]))
    for tag, m in (("s", op.source_model), ("t", op.target_model)):
        assert np.allclose(m._n_obs.cpu().numpy(), g[f"n_{tag}"])                      # same hard assignments
        assert rel(m._running_sum, g[f"sum_{tag}"]) < TOL_STATS
        assert rel(m._running_sum_cov, g[f"sumcov_{tag}"]) < TOL_STATS
        assert rel(m.mean, g[f"mean_{tag}"]) < TOL_STATS and rel(m.variances, g[f"var_{tag}"]) < TOL_STATS
        assert rel(m.weights, g[f"w_{tag}"]) < 1e-6
    assert rel(cost, g["cost"]) < TOL_MATFUN
    assert np.abs(op.transport_matrix.cpu().numpy() - g["coupling"]).max() < TOL_SINKHORN
    assert rel(op.source_model.energy(T(g["probe"]).double()), g["energy_s"]) < TOL_MATFUN
    assert moved.dtype == torch.float32 and moved.is_cuda and rel(moved, g["moved"]) < TOL_MATFUN


def test_gmm_weighted_syrk_soft_assignments_vs_dense(api):
    """soft ('mean' mode) assignments: sum_b w_bk x_b x_b^T from the kernel path == the reference's dense expression
    (gassian_mixture_model.py:109-115) on a size the dense form can still hold (B x d^2 = 2000 x 48^2)."""
    torch.manual_seed(5)
    d, k, b = 48, 5, 2000
    m = api.GaussianMixtureModel(d, mixture_cfg=dict(n_components=k, inference_mode="mean"), dtype=torch.double,
                                 device="cuda").cuda().eval()
    x = torch.randn(b, d, device="cuda", dtype=torch.double) * 0.7 + torch.randn(1, d, device="cuda", dtype=torch.double)
    with torch.no_grad():
        m.mean.copy_(x[:k] * 0.5)
    n, s, ss = m.kmean_iteration(x)
    w, _, _ = m.assign(x)
    dense = (w.transpose(-1, -2) @ (x.unsqueeze(-1) @ x.unsqueeze(-2)).flatten(-2)).unflatten(-1, (d, d))
    assert rel(n, w.sum(-2).cpu()) < 1e-12 and rel(s, (w.transpose(-1, -2) @ x).cpu()) < 1e-12
    assert rel(ss, dense.cpu()) < TOL_STATS


# ------------------------------------------------------------------------------- caller-side layout

def test_strided_token_latents_are_read_in_place(api):
    """ViT tokens [B, T, D] with one operator per token (`permute_and_flatten(batch_first=False)`,
    transport_callback.py:36-43): the view [T, B, D] with strides (D, T D, 1) goes to the kernel without the
    reference's `.contiguous()` copy and gives the same statistics as the copied layout."""
    from ot_vae_lightning_b200 import kernels as K
    from ot_vae_lightning_b200.utils import permute_and_flatten, unflatten_and_unpermute
    torch.manual_seed(3)
    for B, Tk, D in [(700, 5, 128), (300, 3, 512), (180, 4, 96), (64, 2, 40)]:
        lat = torch.randn(B, Tk, D, device="cuda") * 0.8 + torch.randn(1, Tk, D, device="cuda")
        view = permute_and_flatten(lat, (2,), batch_first=False)
        assert view.shape == (Tk, B, D) and view.data_ptr() == lat.data_ptr() and not view.is_contiguous()
        back = unflatten_and_unpermute(view, lat.shape, (2,), batch_first=False)
        assert torch.equal(back, lat)
        out = []
        for x in (view, view.contiguous()):
            n = torch.zeros(Tk, dtype=torch.float64, device="cuda")
            s = torch.zeros(Tk, D, dtype=torch.float64, device="cuda")
            ss = torch.zeros(Tk, D, D, dtype=torch.float64, device="cuda")
            K.stats_update(x, n, s, ss, None)
            K.stats_update(x[:, :37], n, s, ss, None)                     # ragged tail batch, still a strided view
            out.append((n, s, ss))
        x64 = torch.cat([view, view[:, :37]], dim=1).double()
        assert torch.equal(out[0][0], out[1][0]) and float(out[0][0][0]) == B + 37
        assert rel(out[0][1], x64.sum(1).cpu()) < 1e-6 and rel(out[0][2], (x64.transpose(1, 2) @ x64).cpu()) < TOL_STATS
        assert rel(out[0][2], out[1][2].cpu()) < 1e-6
    # the flatten-everything case needs no copy either (MNIST32 CNN latents [B, 128, 1, 1] -> [B, 128])
    lat = torch.randn(250, 128, 1, 1, device="cuda")
    assert permute_and_flatten(lat, (1, 2, 3)).data_ptr() == lat.data_ptr()


def test_golden_operator_variants(api, golden):
    """stochastic (eq. 19) / diagonal / pg_star-blended operators against the unmodified reference's outputs"""
    from tests.test_host_logic import check_operator_variants
    check_operator_variants(api, golden("operator_variants"), "cuda", rtol=TOL_MATFUN, cw_atol=1e-4)



def test_golden_gmm_soft_assignments(api, golden):
    """soft ('mean') source assignments through the sqrt(w)-scaled weighted SYRK and the per-input operator path"""
    from tests.test_host_logic import run_gmm_case
    g = golden("gmm_full_soft")
    op, cost, moved = run_gmm_case(api, g, False, device="cuda", source_mode="mean", rtol=TOL_STATS)
    assert rel(cost, g["cost"]) < TOL_MATFUN
    assert np.abs(op.transport_matrix.cpu().numpy() - g["coupling"]).max() < TOL_SINKHORN
    assert moved.is_cuda and rel(moved, g["moved"]) < TOL_MATFUN


@pytest.mark.parametrize("kind", ["argmax", "mean"])
def test_golden_discrete_transport(api, golden, kind):
    """DiscreteTransport end to end (streaming k-means codebooks -> inverse-distance cost kernel -> Sinkhorn kernel ->
    routing) against the unmodified reference's outputs"""
    from tests.test_host_logic import run_discrete_case
    g = golden("discrete")
    op, cost, moved = run_discrete_case(api, g, kind, device="cuda")
    assert np.allclose(op.source_model._n_obs.cpu().numpy(), g["n_s"]) and np.allclose(op.target_model._n_obs.cpu().numpy(), g["n_t"])
    assert rel(op.source_model.codebook, g["codebook_s"]) < TOL_STATS and rel(op.target_model.codebook, g["codebook_t"]) < TOL_STATS
    assert rel(cost, g[f"cost_{kind}"]) < TOL_SINKHORN
    assert np.abs(op.transport_matrix.cpu().numpy() - g[f"plan_{kind}"]).max() < TOL_SINKHORN
    assert moved.is_cuda and moved.dtype == torch.float32 and rel(moved, g[f"moved_{kind}"]) < TOL_MATFUN


def test_golden_gaussian_barycenter(api, golden):
    """Alvarez-Esteban fixed point (100 iterations of batched Newton-Schulz roots) and the diagonal closed form"""
    from tests.test_host_logic import check_barycenter
    check_barycenter(api, golden("barycenter"), "cuda", TOL_MATFUN)


# ------------------------------------------------------------------------------------------------- FID

def _fid_features(name):
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from fid_stub import CASES, FeatureNet, images
    fsize, n_gen, n_smp, batch = CASES[name]
    net = FeatureNet(fsize)
    gen, smp = images(11, n_gen), images(12, n_smp, gain=0.85, offset=0.05)
    return fsize, batch, net, gen, smp


@pytest.mark.parametrize("name", ["f64", "f2048_deficient", "f2048_full"])
def test_golden_fid_statistics_and_score(golden, oracle, name):
    """`FrechetInceptionDistance` (reference metrics/fid.py:99-130) at feature_size 64 / 2048, with fewer and with more
    observations than features, against the golden of the unmodified reference class and against the oracle.  The
    features are computed once on the CPU (bit-identical to the fixture's) and fed through an identity `net`."""
    from ot_vae_lightning_b200.metrics.fid import FrechetInceptionDistance
    g = golden("fid")
    fsize, batch, net, gen, smp = _fid_features(name)
    fid = FrechetInceptionDistance(net=torch.nn.Identity(), feature_size=fsize, device="cuda")
    st = oracle.FidStats(fsize)
    for lo in range(0, gen.shape[0], batch):
        f = net(gen[lo:lo + batch])
        fid.update(generated=f.cuda())
        st.update(generated_features=f)
    for lo in range(0, smp.shape[0], batch):
        f = net(smp[lo:lo + batch])
        fid.update(samples=f.cuda())
        st.update(sample_features=f)
    assert fid.num_real_obs.cpu().tolist() == g[f"{name}_num_real"].tolist() and fid.num_real_obs.dtype == torch.long
    assert fid.real_sum.dtype == torch.double and fid.real_correlation.shape == (fsize, fsize)
    assert rel(fid.real_sum, g[f"{name}_real_sum"]) < 1e-6 and rel(fid.fake_sum, g[f"{name}_fake_sum"]) < 1e-6
    step = max(1, fsize // 32)
    assert rel(fid.real_correlation[::step, ::step], g[f"{name}_real_corr_sample"]) < 1e-5
    assert rel(fid.fake_correlation, st.corr["fake"]) < 1e-5
    assert abs(float(fid.fake_correlation.trace()) - float(g[f"{name}_fake_trace"])) < 1e-6 * float(g[f"{name}_fake_trace"])
    score = fid.compute()
    assert score.shape == () and score.dtype == torch.double
    want, want_oracle = float(g[f"{name}_score"]), float(st.compute())
    assert abs(float(score) - want) < TOL_MATFUN * want and abs(float(score) - want_oracle) < TOL_MATFUN * want_oracle


def test_fid_with_a_device_feature_net_and_too_few_observations(golden):
    """the `net` forward on the GPU (grey images are replicated to 3 channels, fid.py:100) and the < 1000 rule (:126)"""
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from fid_stub import FeatureNet, images
    from ot_vae_lightning_b200.metrics.fid import FrechetInceptionDistance
    g = golden("fid")
    fid = FrechetInceptionDistance(net=FeatureNet(64).cuda(), feature_size=64, device="cuda")
    gen, smp = images(11, 1100), images(12, 1200, gain=0.85, offset=0.05)
    for lo in range(0, 1100, 275):
        fid.update(generated=gen[lo:lo + 275].cuda())
    fid.update(samples=smp[:999].cuda())
    assert torch.isinf(fid.compute()).all()
    fid.update(samples=smp[999:].cuda())
    assert abs(float(fid.compute()) - float(g["f64_score"])) < TOL_MATFUN * float(g["f64_score"])
    fid.reset()
    assert int(fid.num_real_obs) == 0 and float(fid.real_correlation.abs().sum()) == 0.0


# ------------------------------------------------------------------------------------------------- a6 at d >= 128

def _spectrum_matrix(d, lam, seed):
    g = torch.Generator().manual_seed(seed)
    q, _ = torch.linalg.qr(torch.randn(d, d, generator=g, dtype=torch.double))
    a = (q * lam) @ q.T
    return (a + a.T) / 2


@pytest.mark.parametrize("d", [128, 512, 1024])
def test_min_eig_is_exact_beyond_the_krylov_cap(api, d):
    """min_eig / is_pd / make_psd (reference matrix_utils.py:91-142) on indefinite, near-singular (lambda_min = +-1e-7)
    and rank-deficient matrices at d >= 128, against LAPACK's eigvalsh: the sign decides `is_pd`, and the repaired matrix
    must be positive definite (Cholesky succeeds), which an upper bound on lambda_min does not guarantee."""
    body = torch.logspace(-3, 0, d - 1, dtype=torch.double)
    cases = {
        "indefinite": torch.cat([torch.tensor([-0.37], dtype=torch.double), body]),
        "barely_negative": torch.cat([torch.tensor([-1e-7], dtype=torch.double), body]),
        "barely_positive": torch.cat([torch.tensor([1e-7], dtype=torch.double), body]),
        "two_negative_close": torch.cat([torch.tensor([-2e-7, -1e-7], dtype=torch.double), body[1:]]),
    }
    mats = torch.stack([_spectrum_matrix(d, lam, 70 + i) for i, lam in enumerate(cases.values())])
    g = torch.Generator().manual_seed(5)
    low = torch.randn(d, d // 4, generator=g, dtype=torch.double)
    deficient = low @ low.T / d                                         # rank d/4: a cluster of ~0 eigenvalues
    mats = torch.cat([mats, deficient[None]])
    want = torch.linalg.eigvalsh(mats).amin(-1)
    scale = mats.flatten(1).norm(dim=1)
    got = api.min_eig(mats.cuda()).cpu()
    assert got.dtype == torch.double
    assert float(((got - want).abs() / scale).max()) < 1e-11, (got, want)
    assert api.is_pd(mats[:4].cuda()).cpu().tolist() == (want[:4] > 0).tolist() == [False, False, True, False]
    fixed, shift = api.make_psd(mats.cuda(), strict=True, return_correction=True)
    assert float((shift.cpu() - (want.clamp(max=0).abs() + 1e-8)).abs().max()) < 1e-11 * float(scale.max())
    _, info = torch.linalg.cholesky_ex(fixed.cpu())
    assert info.tolist() == [0] * mats.shape[0]
    # float32 input follows the same path (cast to fp64 inside the kernel)
    assert abs(float(api.min_eig(mats[0].float().cuda())) + 0.37) < 1e-5


# ------------------------------------------------------------------------------------------------- singular 'spsd' arguments

@pytest.mark.parametrize("d,rank", [(10, 4), (256, 100), (512, 500)])
def test_singular_spsd_sqrtm_and_target_covariance(api, oracle, d, rank):
    """Exactly singular SPSD input (a covariance of fewer samples than dimensions).  The reference is NaN there whenever
    `eigh` returns a round-off-negative eigenvalue (probed: tests/golden/make_golden.py note); the oracle clamps those at
    0, and the CUDA path must return that finite root / operator (fp64 Newton-Schulz with a 1e-15 relative ridge)."""
    g = torch.Generator().manual_seed(909 + d)
    low = torch.randn(2, d, rank, generator=g, dtype=torch.double)
    ct_low = low @ low.transpose(-1, -2) / rank
    clamp_sqrt = lambda lam: lam.clamp_min(0).sqrt()
    root = api.sqrtm(ct_low.cuda())
    assert bool(torch.isfinite(root).all()) and rel(root, oracle.spectral_apply(ct_low, clamp_sqrt)) < TOL_MATFUN
    assert rel(root @ root, ct_low) < TOL_MATFUN
    cs = torch.stack([_spectrum_matrix(d, torch.logspace(-1.5, 0, d, dtype=torch.double), 3 + i) for i in range(2)])
    Top, Cw = api.compute_transport_operators(cs.cuda(), ct_low.cuda(), stochastic=False, diag=False, make_pd=True)
    s, si = oracle.sqrtm(cs), oracle.invsqrtm(cs + 1e-8 * torch.eye(d, dtype=torch.double))
    mid = s @ ct_low @ s
    want = si @ oracle.spectral_apply((mid + mid.transpose(-1, -2)) / 2, clamp_sqrt) @ si
    assert rel(Top, want) < TOL_MATFUN and float(Cw.abs().max()) == 0.0
    assert rel(Top @ cs.cuda() @ Top, ct_low) < 5 * TOL_MATFUN


# ------------------------------------------------------------------------------------------------- cfg4 at full size

def test_cfg4_conditional_transport_10x1024_vs_oracle(api, oracle):
    """BASELINE.json configs[3]: `GaussianTransport(10, 1024)` (one operator per class, transport_callback.py:388-453),
    4 d samples per class, against the oracle pipeline (update -> fit -> W2 -> T -> transport) at full width."""
    from ot_vae_lightning_b200.synthetic import gaussian_latents
    L, d, n, bs = 10, 1024, 4096, 1024
    src = torch.stack([gaussian_latents(n, d, seed=120 + k, device="cuda") for k in range(L)])
    tgt = torch.stack([gaussian_latents(n, d, seed=140 + k, device="cuda", shift=1.0, scale=0.8) for k in range(L)])
    op = api.GaussianTransport(L, d, transport_cfg=dict(make_pd=True), source_cfg=dict(dtype=torch.double),
                               target_cfg=dict(dtype=torch.double)).cuda()
    for lo in range(0, n, bs):
        op.update(src[:, lo:lo + bs], tgt[:, lo:lo + bs])
    w2 = op.compute()
    moved = op.transport(src[:, :512])
    s_st, t_st = oracle.GaussianStats(L, d), oracle.GaussianStats(L, d)
    for lo in range(0, n, bs):
        s_st.update(src[:, lo:lo + bs].cpu())
        t_st.update(tgt[:, lo:lo + bs].cpu())
    (mean_s, cov_s), (mean_t, cov_t) = s_st.fit(), t_st.fit()
    want_w2 = oracle.w2_gaussian(mean_s, mean_t, cov_s, cov_t)
    want_T, _ = oracle.transport_operator_full(cov_s, cov_t, 0.0)
    # oracle.apply_transport broadcasts one d x d operator per latent (the reference's bmm, w2_utils.py:515-520): the same
    # product written as a GEMM per class, so that the 10 x 512 x 1024 x 1024 broadcast is never materialised
    want_moved = (src[:, :512].cpu().double() - mean_s[:, None]) @ want_T.transpose(-1, -2) + mean_t[:, None]
    assert w2.shape == (L,) and rel(w2, want_w2) < TOL_MATFUN
    assert rel(op.source_model.mean, mean_s) < TOL_STATS and rel(op.source_model.cov, cov_s) < TOL_STATS
    assert rel(op.transport_operator, want_T) < TOL_MATFUN
    assert moved.dtype == torch.float32 and rel(moved, want_moved) < TOL_MATFUN


@pytest.mark.gpu
@pytest.mark.parametrize("L,d", [(3, 776), (2, 1000), (24, 256)])
def test_batched_operator_on_the_cta_pair_gemm_vs_oracle(L, d, oracle):
    """Batched maps whose products run on the persistent CTA-pair tcgen05 GEMM (256 x 256 tiles; > 200 tiles of 128 x 128 per
    launch), with widths that leave partial row / column tiles and a peer CTA whose B half is entirely out of range, against
    the oracle operator `T = Cs^-1/2 (Cs^1/2 Ct Cs^1/2)^1/2 Cs^-1/2` (w2_utils.py:756-768) and W2^2 (w2_utils.py:40-80)."""
    from ot_vae_lightning_b200 import kernels as K
    g = torch.Generator().manual_seed(31 * L + d)

    def spd(scale):
        q, _ = torch.linalg.qr(torch.randn(L, d, d, generator=g, dtype=torch.float64))
        ev = torch.logspace(0, -2, d, dtype=torch.float64) * scale
        return (q * ev.unsqueeze(-2)) @ q.transpose(-1, -2)

    cs, ct = spd(1.0), spd(1.7)
    ms, mt = torch.randn(L, d, generator=g, dtype=torch.float64), torch.randn(L, d, generator=g, dtype=torch.float64)
    T, w2 = K.transport_operator(cs.cuda(), ct.cuda(), mean_s=ms.cuda(), mean_t=mt.cuda())
    want_T, _ = oracle.transport_operator_full(cs, ct, 0.0)
    assert T.shape == (L, d, d) and rel(T, want_T) < TOL_MATFUN
    assert rel(w2, oracle.w2_gaussian(ms, mt, cs, ct)) < TOL_MATFUN


# ------------------------------------------------------------------------------------------------- dense Sinkhorn, all kernel variants

@pytest.mark.parametrize("n,m,dt", [(1500, 4100, torch.float32), (700, 4099, torch.float32), (2048, 1024, torch.float32),
                                    (5, 8200, torch.float32), (333, 130, torch.float64), (64, 4098, torch.float64)])
def test_sinkhorn_dense_vector_and_scalar_variants_vs_oracle(api, oracle, n, m, dt):
    """`sinkhorn_log` on a materialised cost (reference w2_utils.py:276-319) through every variant of the streaming kernels:
    16-byte and scalar loads (M % 4), warp-per-row and block-per-row (M > 4096), several column blocks and row slabs."""
    g = torch.Generator().manual_seed(n + m)
    x, y = torch.randn(n, 16, generator=g, dtype=torch.double), torch.randn(m, 16, generator=g, dtype=torch.double) + 0.3
    C = oracle.sqeuclidean_cost(x, y)
    C = (C / C.max()).to(dt)
    a = torch.rand(n, generator=g, dtype=torch.double) + 0.2
    a = (a / a.sum()).to(dt)
    b = torch.full((m,), 1.0 / m, dtype=dt)
    plan = api.sinkhorn_log(a.cuda(), b.cuda(), C.cuda(), reg=0.05, max_iter=30, threshold=0.0)
    want = oracle.sinkhorn_log(a.double(), b.double(), C.double(), reg=0.05, max_iter=30, threshold=0.0)
    assert plan.dtype == dt and plan.shape == (n, m)
    tol = 1e-9 if dt == torch.float64 else TOL_SINKHORN
    cost, want_cost = float((plan.double().cpu() * C.double()).sum()), float((want * C.double()).sum())
    assert abs(cost - want_cost) < tol * want_cost
    assert float((plan.double().cpu().sum(1) - want.sum(1)).abs().max()) < tol * float(a.max())
    assert float((plan.double().cpu().sum(0) - want.sum(0)).abs().max()) < tol * float(b.max()) * (1 if dt == torch.float64 else 10)


def test_sharded_summary_matches_fused_solver_summary(api):
    """the plan statistics entry point used by the multi-GPU `check` (world size 1 here) == the solver's own summary"""
    from ot_vae_lightning_b200 import kernels as K
    from ot_vae_lightning_b200 import parallel
    from ot_vae_lightning_b200.synthetic import point_clouds
    for (n, m, d) in [(1024, 768, 128), (300, 200, 20)]:            # fused tcgen05 engine / streaming engine
        x, y = point_clouds(n, m, d, seed=8, device="cuda")
        a = torch.full((n,), 1.0 / n, device="cuda")
        b = torch.full((m,), 1.0 / m, device="cuda")
        res = K.sinkhorn_points(x, y, a, b, reg=0.05, max_iter=25, threshold=0.0)
        scale = 1.0 / float(K.cost_max(x, y, 0).item())
        chk = parallel.sharded_summary(x, y, a, b, res["u"], res["v"], scale, 0.05)
        s = res["summary"].cpu().tolist()
        assert abs(chk["cost"] - s[0]) < 1e-5 * s[0] and abs(chk["mass"] - s[1]) < 1e-6
        assert abs(chk["max_row_err"] - s[2]) < 1e-7 and abs(chk["max_col_err"] - s[3]) < 1e-6


# ------------------------------------------------------------------------------------------------- f3: stochastic operator, noise

@pytest.mark.parametrize("d,pg", [(24, 0.0), (128, 0.3), (256, 0.0)])
def test_stochastic_operator_native_vs_oracle(api, oracle, d, pg):
    """eq. 19 (reference w2_utils.py:774-793) through `otk_transport_operator_stochastic` (Newton-Schulz roots, inverse of
    the source in place of torch.linalg.pinv, fourteen products on the device) against the oracle's eigh / pinv pipeline."""
    lam_s, lam_t = torch.logspace(-1.3, 0, d, dtype=torch.double), torch.logspace(-1.0, 0, d, dtype=torch.double) * 1.8
    cs, ct = _spectrum_matrix(d, lam_s, 31), _spectrum_matrix(d, lam_t, 32)
    Top, Cw = api.compute_transport_operators(cs.cuda(), ct.cuda(), stochastic=True, diag=False, pg_star=pg, make_pd=True)
    want_T, want_Cw = oracle.transport_operator_full_stochastic(cs, ct, pg)
    assert Top.dtype == torch.double and rel(Top, want_T) < TOL_MATFUN
    # with a positive definite source the noise covariance cancels to zero: compare on the scale of the target covariance
    assert float((Cw.cpu() - want_Cw).abs().max()) < 1e-6 * float(ct.abs().max())


def test_stochastic_noise_is_sampled_with_the_requested_covariance(api):
    """`apply_transport` with a non-zero Cw (reference w2_utils.py:522-525): W = Cw^1/2 eps from the Newton-Schulz root and
    the streaming GEMM kernel; the empirical covariance of the noise must be Cw (statistical check, 40 000 draws)."""
    d, n = 32, 40000
    cw = _spectrum_matrix(d, torch.logspace(-1, 0, d, dtype=torch.double), 41).cuda()
    zero = torch.zeros(1, d, dtype=torch.double, device="cuda")
    eye = torch.eye(d, dtype=torch.double, device="cuda").unsqueeze(0)
    torch.manual_seed(5)
    y = api.apply_transport(torch.zeros(n, d, dtype=torch.double, device="cuda"), zero, zero, eye, cw.unsqueeze(0))
    emp = (y.T @ y) / n
    assert rel(emp, cw.cpu()) < 0.05 and float(y.mean(0).abs().max()) < 0.03
    vw = torch.rand(d, dtype=torch.double, device="cuda") + 0.5          # diagonal branch: Cw is the *scale* (reference quirk)
    yd = api.apply_transport(torch.zeros(n, d, dtype=torch.double, device="cuda"), zero[0], zero[0],
                             torch.ones(d, dtype=torch.double, device="cuda"), vw, diag=True)
    assert rel(yd.std(0), vw.cpu()) < 0.03


# ------------------------------------------------------------------------------------------------- f4: strided views into transport

def test_strided_token_latents_are_transported_in_place(api):
    """`GaussianTransport.transport` on the [T, B, D] token view of `permute_and_flatten(batch_first=False)` (strides
    (D, T D, 1)): `otk_apply_transport_prepared_strided` reads the view through TMA without the reference's `.contiguous()`
    copy (utils/__init__.py:260-261) and returns what the copied layout returns; odd strides take the FFMA engine."""
    from ot_vae_lightning_b200.utils import permute_and_flatten
    torch.manual_seed(4)
    for B, Tk, D in [(700, 5, 128), (300, 3, 512), (180, 4, 96), (64, 2, 40)]:
        lat = torch.randn(B, Tk, D, device="cuda") * 0.8 + torch.randn(1, Tk, D, device="cuda")
        tgt = torch.randn(B, Tk, D, device="cuda") * 1.3 - 0.4
        view, tview = permute_and_flatten(lat, (2,), batch_first=False), permute_and_flatten(tgt, (2,), batch_first=False)
        assert not view.is_contiguous()
        op = api.GaussianTransport(Tk, D, transport_cfg=dict(make_pd=True), source_cfg=dict(dtype=torch.double),
                                   target_cfg=dict(dtype=torch.double)).cuda()
        op.update(source_samples=view, target_samples=tview)
        op.compute()
        moved_view, moved_copy = op.transport(view), op.transport(view.contiguous())
        assert moved_view.shape == (Tk, B, D) and torch.equal(moved_view, moved_copy)
        want = (view.double() - op.source_model.mean[:, None]) @ op.transport_operator.transpose(-1, -2) + op.target_model.mean[:, None]
        assert rel(moved_view, want.cpu()) < TOL_MATFUN
        ragged = op.transport(view[:, 11:48])                                   # a slice of the view: offset + same strides
        # (a 37-row slice is in the latency regime: one FFMA launch instead of the tcgen05 kernels - same numbers to fp32)
        assert rel(ragged, moved_copy[:, 11:48].cpu()) < 1e-4
    # feature-sliced latents: row stride 130 (not a multiple of 4 elements) -> FFMA engine, same numbers
    wide = torch.randn(257, 130, device="cuda")
    sl = wide[:, :128]
    op = api.GaussianTransport(128, transport_cfg=dict(make_pd=True), source_cfg=dict(dtype=torch.double),
                               target_cfg=dict(dtype=torch.double)).cuda()
    op.update(source_samples=sl.contiguous(), target_samples=sl.contiguous() * 1.5 + 0.5)
    op.compute()
    assert rel(op.transport(sl), op.transport(sl.contiguous()).cpu()) < 1e-5


def test_paired_update_is_one_launch_and_equals_two_updates(api):
    """`GaussianTransport.update(source, target)` on a batch of the latency regime (the reference's batches of 250,
    tests/test_latent_transport.py:66-98) goes to `otk_stats_update_pair`: ONE kernel launch for both models, bit-identical
    to the two `GaussianModel.update` calls of the reference's `TransportOperator.update`; unequal shapes, large batches and
    stored samples fall back to the per-model path."""
    from ot_vae_lightning_b200 import _native as N_
    lib = N_.load()
    torch.manual_seed(3)
    d = 128
    cfg = dict(dtype=torch.double, reduce_on_update=False)
    pair = api.GaussianTransport(d, transport_cfg=dict(make_pd=True), source_cfg=dict(cfg), target_cfg=dict(cfg)).cuda()
    solo = api.GaussianTransport(d, transport_cfg=dict(make_pd=True), source_cfg=dict(cfg), target_cfg=dict(cfg)).cuda()
    for rows in (250, 7, 256):
        xs, xt = torch.randn(rows, d, device="cuda"), torch.randn(rows, d, device="cuda") * 1.5 + 0.5
        torch.cuda.synchronize()
        l0 = lib.otk_launch_count()
        pair.update(source_samples=xs, target_samples=xt)
        assert lib.otk_launch_count() - l0 == 1
        solo.source_model.update(xs)
        solo.target_model.update(xt)
    xs, xt = torch.randn(300, d, device="cuda"), torch.randn(300, d, device="cuda")          # above the latency regime
    pair.update(source_samples=xs, target_samples=xt); solo.source_model.update(xs); solo.target_model.update(xt)
    # the entry point itself also takes large batches (two back-to-back updates inside the library)
    from ot_vae_lightning_b200 import kernels as K_
    xs, xt = torch.randn(70_000, d, device="cuda"), torch.randn(70_000, d, device="cuda") * 0.7
    assert K_.StatsUpdatePairPlan(pair.source_model._fast_plan(), pair.target_model._fast_plan())(xs, xt)
    solo.source_model.update(xs); solo.target_model.update(xt)
    xs, xt = torch.randn(100, d, device="cuda"), torch.randn(90, d, device="cuda")           # unequal batches
    pair.update(source_samples=xs, target_samples=xt); solo.source_model.update(xs); solo.target_model.update(xt)
    for a, b in ((pair.source_model, solo.source_model), (pair.target_model, solo.target_model)):
        assert torch.equal(a._n_obs, b._n_obs) and torch.equal(a._running_sum, b._running_sum)
        assert torch.equal(a._running_sum_cov, b._running_sum_cov)
    assert float(pair.source_model._n_obs) == 250 + 7 + 256 + 300 + 70_000 + 100


# ------------------------------------------------------------------------------------------------- f2: codebook k-means kernel

@pytest.mark.parametrize("B,K,d,lead", [(1000, 1024, 64, ()), (512, 8192, 128, ()), (250, 48, 20, (3,)), (77, 6, 8, ()),
                                        (3112, 4096, 256, ()), (2100, 2048, 192, (2,))])
def test_kmeans_assign_kernel_vs_oracle(oracle, B, K, d, lead):
    """`otk_kmeans_assign` (nearest codeword + per-codeword counts / sums, no [B, K] matrices) against the oracle's
    energy -> softmax -> one-hot -> weights^T @ samples (reference base.py:206-253) at codebook sizes up to the reference's
    largest configuration (K = 8192, configs/dad/defaults.yaml:70).  Samples sit near codewords, so the arg-min does not
    hinge on round-off.  Cases with B * K >= 2^22 and d >= 192 take the tcgen05 contraction (row chunks of the score matrix; 3112
    rows at K = 4096 = one 3072-row chunk + a 40-row tail on the FFMA tiles), the others the FFMA tiles."""
    from ot_vae_lightning_b200 import kernels as K_
    g = torch.Generator().manual_seed(B + K)
    book = torch.randn(*lead, K, d, generator=g)
    pick = torch.randint(0, K, (*lead, B), generator=g)
    x = torch.gather(book, -2, pick.unsqueeze(-1).expand(*lead, B, d)) + 0.02 * torch.randn(*lead, B, d, generator=g)
    index, counts, sums = K_.kmeans_assign(x.cuda(), book.cuda(), sums_dtype=torch.float64)
    want_counts, want_sums, want_index = oracle.kmeans_step(x, book)
    assert index.dtype == torch.int64 and torch.equal(index.cpu(), want_index) and torch.equal(index.cpu(), pick)
    assert torch.equal(counts.cpu(), want_counts) and float(counts.sum()) == B * max(1, int(torch.Size(lead).numel()))
    assert rel(sums, want_sums) < 1e-6


def test_codebook_model_streaming_update_uses_the_kernel_and_matches_the_dense_path(api, monkeypatch):
    """`CodebookModel.update` in hard mode through the fused kernel == the same model forced onto the dense
    energy / softmax / one-hot / matmul path (the reference's formulation), batch after batch."""
    from ot_vae_lightning_b200.ot.distribution_models import base as dm_base
    g = torch.Generator().manual_seed(12)
    centres = torch.randn(16, 32, generator=g) * 4
    data = (centres[torch.randint(0, 16, (2000,), generator=g)] + 0.3 * torch.randn(2000, 32, generator=g)).cuda()
    models = []
    for fused in (True, False):
        torch.manual_seed(7)
        cb = api.CodebookModel(32, mixture_cfg=dict(n_components=16), dtype=torch.double, update_decay=0.9).cuda()
        if not fused:
            monkeypatch.setattr(dm_base.MixtureMixin, "_nearest_component_kernel_applies", lambda self, s: False)
        for lo in range(0, 2000, 250):
            torch.manual_seed(100 + lo)
            cb.update(data[lo:lo + 250])
        models.append(cb)
    assert rel(models[0].codebook, models[1].codebook.cpu()) < 1e-6 and rel(models[0]._n_obs, models[1]._n_obs.cpu()) < 1e-9
    assert rel(models[0]._running_sum, models[1]._running_sum.cpu()) < 1e-6


def test_native_fit_matches_the_stepwise_fit_and_skips_unseen_indices(api, monkeypatch):
    """`otk_gaussian_fit` (one launch: mean, raw covariance, symmetrised + shifted operand) == the step-by-step fit
    (mean_cov -> masked assignment, reference gaussian_model.py:159-183); a leading index without observations keeps its
    initial mean / covariance."""
    from ot_vae_lightning_b200.ot.distribution_models import gaussian_model as gm_mod
    torch.manual_seed(9)
    x = torch.randn(3, 500, 24, device="cuda") * 1.7 + 0.4
    models = []
    for native in (True, False):
        torch.manual_seed(1)
        gm = api.GaussianModel(3, 24, w2_cfg=dict(make_pd=True), dtype=torch.double, reduce_on_update=False).cuda()
        if not native:
            monkeypatch.setattr(gm_mod.GaussianModel, "_native_fit", lambda self, c, s, r=False: False)
        gm.update(x)
        gm._n_obs[1] = 0                                   # class 1 never observed
        operand = torch.empty(3, 24, 24, dtype=torch.double, device="cuda")
        gm.fit(cov_operand=operand, operand_shift=1e-8)
        models.append((gm, operand))
    (a, op_a), (b, op_b) = models
    assert rel(a.mean, b.mean.cpu()) < 1e-12 and rel(a.parametrizations.cov.original[[0, 2]], b.parametrizations.cov.original[[0, 2]].cpu()) < 1e-12
    assert torch.equal(a.mean[1], a.vec_init[1]) and torch.equal(a.parametrizations.cov.original[1], a.cov_init[1])
    assert rel(op_a[[0, 2]], op_b[[0, 2]].cpu()) < 1e-12
    raw = a.parametrizations.cov.original
    want = torch.triu(raw) + torch.triu(raw, 1).transpose(-1, -2) + 1e-8 * torch.eye(24, dtype=torch.double, device="cuda")
    assert rel(op_a, want.cpu()) < 1e-14


@pytest.mark.parametrize("n,m,d", [(2048, 4096, 128), (1024, 1280, 72), (1500, 1024, 512)])
def test_cost_matrix_tensor_core_contraction_vs_fp64(n, m, d):
    """`otk_cost_matrix` at sizes where the x.y^T contraction runs on tcgen05 (3xTF32) + one elementwise pass, both cost kinds
    (squared distance, w2_utils.py:121-125; inverse distance, codebook_model.py:155-160), against fp64."""
    from ot_vae_lightning_b200 import kernels as K
    g = torch.Generator().manual_seed(n + d)
    x, y = torch.randn(n, d, generator=g), torch.randn(m, d, generator=g) * 1.2 + 0.1
    x64, y64 = x.double(), y.double()
    sq = (x64 * x64).sum(-1, keepdim=True) + (y64 * y64).sum(-1)[None] - 2 * x64 @ y64.T
    got = K.cost_matrix(x.cuda(), y.cuda(), cost=0, scale=0.5)
    assert got.shape == (n, m) and rel(got, 0.5 * sq) < 2e-6
    assert float((got.double().cpu() - 0.5 * sq).abs().max()) < 1e-5 * float(sq.max())
    inv = K.cost_matrix(x.cuda(), y.cuda(), cost=1)
    assert rel(inv, 1.0 / (sq.clamp_min(0).sqrt() + 1e-8)) < 1e-5
