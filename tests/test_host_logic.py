"""Host-side logic of the Python mirror, run WITHOUT a GPU: the kernel wrappers are swapped for oracle-backed
stand-ins (tests/fake_kernels.py), so what is tested here is everything above the C ABI - argument validation and its
error conditions, broadcasting over leading dims, the GaussianModel / GaussianTransport state machine and state_dict,
layout helpers - against the golden outputs of the unmodified reference."""
import numpy as np
import pytest
import torch

from tests import fake_kernels


@pytest.fixture()
def api():
    import ot_vae_lightning_b200.ot as ot
    with fake_kernels.installed():
        yield ot


def T(a):
    return torch.from_numpy(np.asarray(a))


def close(got, want, rtol=1e-8, atol=1e-10):
    got, want = got.detach().double(), T(want).double()
    assert got.shape == want.shape, (got.shape, want.shape)
    assert torch.allclose(got, want, rtol=rtol, atol=atol), (got - want).abs().max().item()


def _run(api, g, lead, d, decay=None, pg_star=0.0):
    cfg = dict(dtype=torch.double, device="cpu")
    if decay is not None:
        cfg["update_decay"] = decay
    op = api.GaussianTransport(*lead, d, transport_cfg=dict(make_pd=True, pg_star=pg_star), source_cfg=dict(cfg),
                               target_cfg=dict(cfg))
    src, tgt, bs = T(g["src"]), T(g["tgt"]), int(g["batch"])
    for lo in range(0, src.shape[-2], bs):
        op.update(source_samples=src[..., lo:lo + bs, :])
    for lo in range(0, tgt.shape[-2], bs):
        op.update(target_samples=tgt[..., lo:lo + bs, :])
    w2 = op.compute()
    moved = op.transport(src)
    for name, key in (("_n_obs", "n_s"), ("_running_sum", "sum_s"), ("_running_sum_cov", "sumcov_s"), ("mean", "mean_s"),
                      ("cov", "cov_s")):
        close(getattr(op.source_model, name), g[key])
    close(op.target_model.cov, g["cov_t"]); close(w2, g["w2"]); close(op.transport_operator, g["T"])
    assert moved.dtype == torch.float32 and torch.allclose(moved, T(g["moved"]), rtol=1e-5, atol=1e-5)
    return op


def test_gaussian_transport_state_machine(api, golden):
    op = _run(api, golden("gaussian_d16"), (), 16)
    keys = {"mean", "vec_init", "mat_init", "cov_init", "_running_sum", "_running_sum_cov", "_n_obs",
            "parametrizations.cov.original"}
    assert set(op.source_model.state_dict()) == keys
    op.reset()
    assert op.transport_operator is None and float(op.source_model._n_obs) == 0
    close(op.source_model.mean, op.source_model.vec_init.numpy())
    with pytest.raises(ValueError):
        op.transport(torch.zeros(4, 15))


def test_paired_update_hook_falls_back_to_the_per_model_updates(api):
    """`GaussianTransport.update(source, target)` tries the one-launch pair entry only for CUDA-resident models with cached
    native plans (`GaussianModel._fast_plan`); on the host mirror alone (no device buffers) the plans do not exist and the
    call is exactly the reference's two `GaussianModel.update`s (ot/transport/base.py `update`), also for unequal batches and
    for a single side."""
    torch.manual_seed(0)
    op = api.GaussianTransport(6, transport_cfg=dict(make_pd=True), source_cfg=dict(dtype=torch.double, device="cpu"),
                               target_cfg=dict(dtype=torch.double, device="cpu"))
    assert op.source_model._fast_plan() is None and op.target_model._fast_plan() is None
    xs, xt = torch.randn(40, 6), torch.randn(40, 6) * 2 + 1
    op.update(source_samples=xs, target_samples=xt)
    op.update(source_samples=xs[:7], target_samples=xt[:9])
    op.update(target_samples=xt[:3])
    assert float(op.source_model._n_obs) == 47 and float(op.target_model._n_obs) == 52
    both = torch.cat([xs, xs[:7]]).double()
    close(op.source_model._running_sum, both.sum(0).numpy())
    close(op.source_model._running_sum_cov, (both.T @ both).numpy())


def test_leading_dims_pgstar_and_ema(api, golden):
    g = golden("gaussian_lead2_d8")
    _run(api, g, (2,), 8, pg_star=float(g["pg_star"]))
    g = golden("gaussian_ema_d8")
    _run(api, g, (), 8, decay=float(g["decay"]))


def test_functional_api_on_golden(api, golden):
    g = golden("w2_d3")
    m1, m2, c1, c2 = (T(g[k]) for k in ("m1", "m2", "c1", "c2"))
    close(api.w2_gaussian(m1, m2, c1, c2), g["w2"])
    Top, Cw = api.compute_transport_operators(c1, c2, stochastic=False, diag=False, make_pd=True)
    close(Top, g["T"], rtol=1e-7)
    y = api.apply_transport(T(g["x"]), m1.unsqueeze(-2), m2.unsqueeze(-2), Top.unsqueeze(-3), Cw.unsqueeze(-3))
    close(y, g["y"], rtol=1e-5, atol=1e-5)
    gm = golden("matrix")
    mean, cov = api.mean_cov(T(gm["sum"]), T(gm["sum_cov"]), T(gm["n"]))
    close(mean, gm["mean"]); close(cov, gm["cov"])
    fixed, shift = api.make_psd(T(gm["indef"]), strict=True, return_correction=True)
    close(fixed, gm["repaired"]); close(shift, gm["shift"])
    gs = golden("sinkhorn_points")
    close(api.sinkhorn_log(T(gs["a"]), T(gs["b"]), T(gs["C"]), reg=0.05, max_iter=25, threshold=0.0), gs["plan_fixed25"])


def test_validation_error_conditions(api):
    eye, v = torch.eye(4, dtype=torch.double), torch.zeros(4, dtype=torch.double)
    for bad in (lambda: api.w2_gaussian(v, v, eye, [[1.0]]),
                lambda: api.w2_gaussian(v, v, eye, -eye),
                lambda: api.w2_gaussian(v, v, eye, eye + torch.triu(torch.ones(4, 4, dtype=torch.double), 1)),
                lambda: api.w2_gaussian(v, v[:3], eye, eye),
                lambda: api.w2_gaussian(torch.zeros((), dtype=torch.double), v, eye, eye),
                lambda: api.apply_transport(v, v, v, eye, None),
                lambda: api.batch_ot_gmm(torch.zeros(3, 4), torch.zeros(2, 4), -torch.ones(3, 4), torch.ones(2, 4), diag=True),
                lambda: api.batch_ot_gmm(torch.zeros(3, 4), torch.zeros(2, 4), torch.ones(3, 4), torch.ones(2, 4), diag=True,
                                         weight_source=torch.tensor([0.5, 0.2, 0.2]))):
        with pytest.raises(ValueError):
            bad()
    with pytest.warns(UserWarning):
        assert float(api.w2_gaussian(v, v, eye, -eye, make_pd=True, verbose=True)) > 0


def test_diag_variants_and_gmm(api):
    g = torch.Generator().manual_seed(0)
    ms, mt = torch.randn(2, 5, 3, generator=g), torch.randn(2, 4, 3, generator=g)
    vs, vt = torch.rand(2, 5, 3, generator=g) + 0.1, torch.rand(2, 4, 3, generator=g) + 0.1
    D = api.batch_w2_dissimilarity_gaussian_diag(ms, mt, vs, vt)
    want = ((ms[:, :, None] - mt[:, None]) ** 2).sum(-1) + ((vs.sqrt()[:, :, None] - vt.sqrt()[:, None]) ** 2).sum(-1)
    close(D, want.double().numpy(), rtol=1e-5, atol=1e-6)
    cost, plan = api.batch_ot_gmm(ms, mt, vs, vt, diag=True, reg=0.05, max_iter=200, threshold=1e-9)
    assert plan.shape == (2, 5, 4) and torch.allclose(plan.sum(-1), torch.full((2, 5), 0.2, dtype=plan.dtype), atol=1e-6)
    Tm, Cw = api.compute_transport_operators(vs, vs * 4, stochastic=False, diag=True)
    close(Tm, torch.full_like(vs, 2.0).double().numpy(), rtol=1e-6)
    mb, vb = api.gaussian_barycenter(ms, vs, torch.full((2, 5), 0.2), diag=True)
    assert mb.shape == (2, 3) and vb.shape == (2, 3)


def test_layout_helpers_round_trip():
    from ot_vae_lightning_b200.utils import ema, permute_and_flatten, unflatten_and_unpermute, unsqueeze_like
    x = torch.randn(10, 1, 2, 3, 4, 5)
    assert permute_and_flatten(x, (1, 3)).shape == (10, 40, 3)
    assert permute_and_flatten(x, (1, 3), batch_first=False).shape == (40, 10, 3)
    assert permute_and_flatten(x, (1, 3), flatten_batch=True).shape == (400, 3)
    for kw in (dict(), dict(batch_first=False), dict(flatten_batch=True)):
        assert torch.equal(unflatten_and_unpermute(permute_and_flatten(x, (1, 3), **kw), x.shape, (1, 3), **kw), x)
    lat = torch.randn(6, 128, 1, 1)  # MNIST32 CNN-VAE latent, transport_dims=(1,2,3) -> [B,128]
    assert permute_and_flatten(lat, (1, 2, 3)).shape == (6, 128)
    assert permute_and_flatten(lat, (1, 2, 3), flatten_batch=True).shape == (768,)  # reference quirk (utils:258)
    assert unsqueeze_like(torch.ones(3), torch.ones(3, 4, 5)).shape == (3, 1, 1)
    with pytest.raises(ValueError):
        unsqueeze_like(torch.ones(3, 4), torch.ones(3))
    assert float(ema(torch.tensor(2.0), torch.tensor(4.0), None)) == 6.0
    assert float(ema(torch.tensor(2.0), torch.tensor(4.0), 0.5)) == 3.0


def test_install_as_reference_aliases():
    import sys
    import ot_vae_lightning_b200 as pkg
    saved = {k: v for k, v in sys.modules.items() if k == "ot_vae_lightning" or k.startswith("ot_vae_lightning.")}
    for k in saved:
        del sys.modules[k]
    try:
        pkg.install_as_reference()
        from ot_vae_lightning.ot.w2_utils import mean_cov, w2_gaussian  # noqa: F401
        from ot_vae_lightning.ot.transport.gaussian_transport import GaussianTransport
        import ot_vae_lightning_b200.ot as ot
        assert GaussianTransport is ot.GaussianTransport
    finally:
        for k in [k for k in sys.modules if k == "ot_vae_lightning" or k.startswith("ot_vae_lightning.")]:
            del sys.modules[k]
        sys.modules.update(saved)


# ------------------------------------------------------------------------------------------------- GMM

def run_gmm_case(api, g, diag, device="cpu", rtol=1e-7, source_mode="argmax"):
    d, bs = int(g["src"].shape[1]), int(g["batch"])
    cfg = dict(dtype=torch.double, device=device)
    op = api.GMMTransport(d, transport_type="argmax",
                          transport_cfg=dict(diag=diag, stochastic=False, make_pd=True, dtype=torch.double),
                          source_cfg=dict(mixture_cfg=dict(n_components=int(g["n_s"].shape[0]), training_mode=source_mode,
                                                           inference_mode=source_mode), **cfg),
                          target_cfg=dict(mixture_cfg=dict(n_components=int(g["n_t"].shape[0])), **cfg))
    src, tgt, probe = (T(g[k]).to(device) for k in ("src", "tgt", "probe"))
    torch.manual_seed(11)                       # same seeds as tests/golden/make_golden.py::run_gmm
    for lo in range(0, src.shape[0], bs):
        op.update(source_samples=src[lo:lo + bs])
    torch.manual_seed(12)
    for lo in range(0, tgt.shape[0], bs):
        op.update(target_samples=tgt[lo:lo + bs])
    cost = op.compute()
    torch.manual_seed(13)
    moved = op.transport(probe)
    for tag, m in (("s", op.source_model), ("t", op.target_model)):
        close(m._n_obs.cpu(), g[f"n_{tag}"], rtol=rtol)
        close(m._running_sum.cpu(), g[f"sum_{tag}"], rtol=rtol, atol=1e-6)
        close(m._running_sum_cov.cpu(), g[f"sumcov_{tag}"], rtol=max(rtol, 1e-6), atol=1e-3)
        close(m.mean.cpu(), g[f"mean_{tag}"], rtol=max(rtol, 1e-6), atol=1e-6)
        close(m.variances.cpu(), g[f"var_{tag}"], rtol=1e-4, atol=1e-5)
        close(m.weights.cpu(), g[f"w_{tag}"], rtol=rtol)
    return op, cost, moved


@pytest.mark.parametrize("name,diag", [("gmm_full_argmax", False), ("gmm_diag_argmax", True)])
def test_gmm_transport_against_reference_golden(api, golden, name, diag):
    g = golden(name)
    op, cost, moved = run_gmm_case(api, g, diag)
    close(cost, g["cost"], rtol=1e-6)
    close(op.transport_matrix, g["coupling"], rtol=1e-5, atol=1e-9)
    close(op.source_model.energy(T(g["probe"]).double()), g["energy_s"], rtol=1e-6, atol=1e-6)
    assert moved.dtype == torch.float32 and torch.allclose(moved, T(g["moved"]), rtol=1e-4, atol=1e-4)
    keys = set(op.source_model.state_dict())
    assert {"mean", "_running_sum", "_running_sum_cov", "_n_obs", "weight_init", "parametrizations.cov.original",
            "parametrizations._weights.original"} <= keys
    op.reset()
    assert op.transport_matrix is None and float(op.source_model._n_obs.sum()) == 0


# ------------------------------------------------------------------------------- operator variants

def check_operator_variants(api, g, dev, rtol, cw_atol):
    cs, ct, vs, vt = (T(g[k]).to(dev) for k in ("cs", "ct", "vs", "vt"))
    for tag, pg in (("p0", 0.0), ("p3", 0.3)):
        Tm, Cw = api.compute_transport_operators(cs.clone(), ct.clone(), stochastic=True, diag=False, pg_star=pg, make_pd=True)
        want = T(g[f"full_st_T_{tag}"])
        assert ((Tm.cpu() - want).norm() / want.norm()).item() < rtol
        assert (Cw.cpu() - T(g[f"full_st_Cw_{tag}"])).abs().max().item() < cw_atol     # analytically zero for a full-rank source
        for st, key in ((False, "diag"), (True, "diag_st")):
            Tm, Cw = api.compute_transport_operators(vs.clone(), vt.clone(), stochastic=st, diag=True, pg_star=pg)
            close(Tm.cpu(), g[f"{key}_T_{tag}"], rtol=1e-9)
            close(Cw.cpu(), g[f"{key}_Cw_{tag}"], rtol=1e-6, atol=1e-12)
    Td, Cwd = api.compute_transport_operators(vs.clone(), vt.clone(), stochastic=False, diag=True)
    y = api.apply_transport(T(g["x"]).to(dev), T(g["ms"]).to(dev), T(g["mt"]).to(dev), Td.unsqueeze(-2), Cwd.unsqueeze(-2),
                            diag=True)
    close(y.cpu(), g["y_diag"], rtol=1e-9)


def test_operator_variants_against_reference_golden(api, golden):
    check_operator_variants(api, golden("operator_variants"), "cpu", rtol=1e-6, cw_atol=1e-6)


def test_gmm_soft_assignments_against_reference_golden(api, golden):
    """'mean' (soft) source assignments: fractional weighted statistics and one Gaussian map per input"""
    g = golden("gmm_full_soft")
    op, cost, moved = run_gmm_case(api, g, False, source_mode="mean", rtol=1e-6)
    close(cost, g["cost"], rtol=1e-6)
    close(op.transport_matrix, g["coupling"], rtol=1e-5, atol=1e-9)
    assert torch.allclose(moved, T(g["moved"]), rtol=1e-4, atol=1e-4)


def test_public_surface_matches_the_reference_exports():
    """every name the reference star-exports from `ot_vae_lightning.ot` (matrix_utils.py:20-31,
    w2_utils.py:26-36, the model / transport classes) resolves here, also under the reference's module paths."""
    import importlib

    import ot_vae_lightning_b200 as pkg
    import ot_vae_lightning_b200.ot as ot
    names = ["eye_like", "sqrtm", "invsqrtm", "is_spd", "is_pd", "is_symmetric", "min_eig", "make_psd", "mean_cov",
             "STABILITY_CONST", "w2_gaussian", "batch_w2_dissimilarity_gaussian_diag", "batch_w2_dissimilarity_gaussian",
             "batch_ot_gmm", "sinkhorn_log", "gaussian_barycenter", "compute_transport_operators", "apply_transport",
             "W2Mixin", "GaussianModel", "CodebookModel", "CategoricalEmbeddings", "GaussianMixtureModel",
             "TransportOperator", "GaussianTransport", "DiscreteTransport", "GMMTransport"]
    assert [n for n in names if not hasattr(ot, n)] == []
    from ot_vae_lightning_b200.ot import w2_utils
    assert hasattr(w2_utils, "mean_cov")        # w2_utils re-exports matrix_utils (reference w2_utils.py:24, fid.py:24)
    pkg.install_as_reference("ot_vae_lightning_surface_check")
    for mod, cls in (("ot.transport.gaussian_transport", "GaussianTransport"), ("ot.transport.gmm_transport", "GMMTransport"),
                     ("ot.transport.discrete_transport", "DiscreteTransport"),
                     ("ot.distribution_models.gassian_mixture_model", "GaussianMixtureModel"),
                     ("ot.distribution_models.gaussian_model", "GaussianModel"), ("metrics.fid", "FrechetInceptionDistance")):
        assert hasattr(importlib.import_module(f"ot_vae_lightning_surface_check.{mod}"), cls)


# ------------------------------------------------------------------------------------ DiscreteTransport

def run_discrete_case(api, g, kind, device="cpu"):
    d = int(g["src"].shape[1])
    cfg = dict(dtype=torch.double, device=device)
    op = api.DiscreteTransport(d, transport_type=kind, sinkhorn_reg=0.05, sinkhorn_max_iter=300, sinkhorn_threshold=1e-9,
                               source_cfg=dict(mixture_cfg=dict(n_components=int(g["n_s"].shape[0])), **cfg),
                               target_cfg=dict(mixture_cfg=dict(n_components=int(g["n_t"].shape[0])), **cfg)).to(device)
    src, tgt, bs = T(g["src"]).to(device), T(g["tgt"]).to(device), int(g["batch"])
    torch.manual_seed(21)                        # same seeds as tests/golden/make_golden.py::case_discrete
    for lo in range(0, src.shape[0], bs):
        op.update(source_samples=src[lo:lo + bs])
    torch.manual_seed(22)
    for lo in range(0, tgt.shape[0], bs):
        op.update(target_samples=tgt[lo:lo + bs])
    cost = op.compute()
    torch.manual_seed(23)
    moved = op.transport(T(g["probe"]).to(device))
    return op, cost, moved


@pytest.mark.parametrize("kind", ["argmax", "mean"])
def test_discrete_transport_against_reference_golden(api, golden, kind):
    g = golden("discrete")
    op, cost, moved = run_discrete_case(api, g, kind)
    close(op.source_model.codebook, g["codebook_s"], rtol=1e-9)
    close(op.target_model.codebook, g["codebook_t"], rtol=1e-9)
    close(op.source_model._n_obs, g["n_s"]); close(op.target_model._n_obs, g["n_t"])
    close(op.source_model._running_sum, g["sum_s"], rtol=1e-9)
    close(op.source_model.weights, g["w_s"], rtol=1e-9); close(op.target_model.weights, g["w_t"], rtol=1e-9)
    close(cost, g[f"cost_{kind}"], rtol=1e-7)
    close(op.transport_matrix, g[f"plan_{kind}"], rtol=1e-6, atol=1e-10)
    assert moved.dtype == torch.float32 and torch.allclose(moved, T(g[f"moved_{kind}"]), rtol=1e-5, atol=1e-6)


def check_barycenter(api, g, dev, tol):
    mean, cov, var, w = (T(g[k]).to(dev) for k in ("mean", "cov", "var", "w"))
    torch.manual_seed(31)
    mb, cb = api.gaussian_barycenter(mean, cov, w, diag=False, n_iter=100)
    assert ((mb.cpu() - T(g["mean_b"])).norm() / T(g["mean_b"]).norm()).item() < 1e-12
    assert ((cb.cpu() - T(g["cov_b"])).norm() / T(g["cov_b"]).norm()).item() < tol
    mb, vb = api.gaussian_barycenter(mean, var, w, diag=True)
    close(mb.cpu(), g["mean_b_diag"]); close(vb.cpu(), g["var_b_diag"])


def test_gaussian_barycenter_against_reference_golden(api, golden):
    check_barycenter(api, golden("barycenter"), "cpu", 1e-8)
