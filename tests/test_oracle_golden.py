"""Pins oracle/ot_oracle.py against outputs of the UNMODIFIED reference (tests/golden/*.npz, made by
tests/golden/make_golden.py).  fp64 on both sides, so the tolerance is round-off only."""
import numpy as np
import torch

from oracle import ot_oracle as O

RTOL, ATOL = 1e-10, 1e-12


def T(a):
    return torch.from_numpy(np.asarray(a))


def close(got, want, rtol=RTOL, atol=ATOL):
    got, want = torch.as_tensor(got).double(), T(want).double()
    assert got.shape == want.shape, (got.shape, want.shape)
    assert torch.allclose(got, want, rtol=rtol, atol=atol), (got - want).abs().max().item()


def test_matrix_primitives(golden):
    g = golden("matrix")
    A = T(g["A"])
    close(O.sqrtm(A), g["sqrtm"])
    close(O.invsqrtm(A), g["invsqrtm"])
    close(O.min_eig(A), g["min_eig"])
    close(O.min_eig(T(g["indef"])), g["indef_min_eig"])
    repaired, shift = O.make_psd(T(g["indef"]), strict=True)
    close(repaired, g["repaired"]); close(shift, g["shift"])
    mean, cov = O.mean_cov(T(g["sum"]), T(g["sum_cov"]), T(g["n"]))
    close(mean, g["mean"]); close(cov, g["cov"])
    assert O.is_symmetric(T(g["asym"])).tolist() == g["asym_is_symmetric"].tolist()
    assert (O.is_symmetric(A) & O.is_pd(A)).tolist() == g["A_is_spd"].tolist()
    assert O.is_pd(T(g["indef"])).tolist() == g["indef_is_pd"].tolist()


def _pipeline(g, lead_decay=None, pg_star=0.0):
    src, tgt = T(g["src"]), T(g["tgt"])
    out = O.gaussian_transport_pipeline(src, tgt, int(g["batch"]), decay=lead_decay, pg_star=pg_star)
    for k in ("mean_s", "cov_s", "mean_t", "cov_t", "w2", "T"):
        close(out[k], g[k], rtol=1e-9, atol=1e-10)
    got, want = out["moved"], T(g["moved"])
    assert got.dtype == want.dtype == torch.float32
    assert torch.allclose(got, want, rtol=1e-6, atol=1e-6)


def test_gaussian_pipeline_d16(golden):
    g = golden("gaussian_d16")
    _pipeline(g)
    st = O.GaussianStats(16)
    src = T(g["src"])
    for lo in range(0, src.shape[0], 100):
        st.update(src[lo:lo + 100])
    close(st.n_obs, g["n_s"]); close(st.sum, g["sum_s"]); close(st.sum_cov, g["sumcov_s"])
    assert np.allclose(g["Cw"], 0)


def test_gaussian_pipeline_leading_dims_and_pgstar(golden):
    g = golden("gaussian_lead2_d8")
    _pipeline(g, pg_star=float(g["pg_star"]))


def test_gaussian_pipeline_ema(golden):
    g = golden("gaussian_ema_d8")
    _pipeline(g, lead_decay=float(g["decay"]))


def test_w2_functions(golden):
    for name in ("w2_d3", "w2_d24"):
        g = golden(name)
        m1, m2, c1, c2 = (T(g[k]) for k in ("m1", "m2", "c1", "c2"))
        close(O.w2_gaussian(m1, m2, c1, c2), g["w2"], rtol=1e-9, atol=1e-10)
        Top, _ = O.transport_operator_full(c1, c2)
        close(Top, g["T"], rtol=1e-8, atol=1e-9)
        if "x" in g:
            close(O.apply_transport(T(g["x"]), m1, m2, Top), g["y"], rtol=1e-8, atol=1e-9)


def test_sinkhorn(golden):
    g = golden("sinkhorn_points")
    a, b, C = T(g["a"]), T(g["b"]), T(g["C"])
    close(O.sqeuclidean_cost(T(g["x"]), T(g["y"])) /
          O.sqeuclidean_cost(T(g["x"]), T(g["y"])).amax(dim=(-1, -2), keepdim=True), g["C"])
    close(O.sinkhorn_log(a, b, C, reg=0.05, max_iter=25, threshold=0.0), g["plan_fixed25"])
    close(O.sinkhorn_log(a, b, C, reg=0.05, max_iter=1000, threshold=1e-6), g["plan_thr1e6"])
    g3 = golden("sinkhorn_3x3")
    close(O.sinkhorn_log(T(g3["a"]), T(g3["b"]), T(g3["C"]), reg=1e-5, max_iter=1000, threshold=1e-8), g3["plan"])


def test_energy_cost(golden):
    g = golden("energy")
    close(O.inverse_distance_energy(T(g["pts"]), T(g["codebook"])), g["energy"])


def test_operator_variants(golden):
    """stochastic / diagonal / pg_star-blended operators (ot/w2_utils.py:714-793)"""
    g = golden("operator_variants")
    cs, ct, vs, vt = T(g["cs"]), T(g["ct"]), T(g["vs"]), T(g["vt"])
    for tag, pg in (("p0", 0.0), ("p3", 0.3)):
        Tm, Cw = O.transport_operator_full_stochastic(cs, ct, pg)
        close(Tm, g[f"full_st_T_{tag}"], rtol=1e-6, atol=1e-8)
        assert (Cw - T(g[f"full_st_Cw_{tag}"])).abs().max().item() < 1e-6         # analytically zero: round-off on both sides
        Tm, Cw = O.transport_operator_diag(vs, vt, pg)
        close(Tm, g[f"diag_T_{tag}"]); close(Cw, g[f"diag_Cw_{tag}"])
        Tm, Cw = O.transport_operator_diag_stochastic(vs, vt, pg)
        close(Tm, g[f"diag_st_T_{tag}"]); close(Cw, g[f"diag_st_Cw_{tag}"], rtol=1e-6, atol=1e-12)


def test_gmm_layer(golden):
    """energies, component OT and hard-assignment transport of the GMM layer on the reference's fitted parameters"""
    for name, diag in (("gmm_full_argmax", False), ("gmm_diag_argmax", True)):
        g = golden(name)
        probe = T(g["probe"]).double()
        ms, mt, vs, vt, ws, wt = (T(g[k]) for k in ("mean_s", "mean_t", "var_s", "var_t", "w_s", "w_t"))
        energy = O.gmm_energy(probe, ms, vs, ws, diag)
        close(energy, g["energy_s"], rtol=1e-9, atol=1e-9)
        cost, plan = O.ot_gmm(ms, mt, vs, vt, ws, wt, diag, max_iter=100)
        close(cost, g["cost"], rtol=1e-9); close(plan, g["coupling"], rtol=1e-8, atol=1e-12)
        src_idx = energy.argmax(-1)
        tgt_idx = (torch.nn.functional.one_hot(src_idx, ws.numel()).double() @ plan).argmax(-1)
        moved = O.gmm_transport_hard(probe, src_idx, tgt_idx, ms, mt, vs, vt, diag)
        assert torch.allclose(moved.float(), T(g["moved"]), rtol=1e-5, atol=1e-5)
    # weighted statistics: one-hot weights reduce to per-group batch statistics
    x = torch.randn(50, 4, dtype=torch.double, generator=torch.Generator().manual_seed(1))
    w = torch.nn.functional.one_hot(torch.arange(50) % 3, 3).double()
    n, s, ss = O.gmm_weighted_stats(x, w, diag=False)
    for k in range(3):
        nk, sk, ssk = O.batch_stats(x[k::3])
        close(n[k], nk); close(s[k], sk); close(ss[k], ssk)


def test_discrete_transport_on_fitted_codebooks(golden):
    """cost = source.energy(target codebook) (inverse distance, codebook_model.py:155-160), Sinkhorn plan and total cost
    of DiscreteTransport.compute (discrete_transport.py:55-68) on the reference's fitted codebooks"""
    g = golden("discrete")
    cb_s, cb_t, w_s, w_t = (T(g[k]) for k in ("codebook_s", "codebook_t", "w_s", "w_t"))
    cost = O.inverse_distance_energy(cb_t, cb_s)
    plan = O.sinkhorn_log(w_s, w_t, cost, reg=0.05, max_iter=300, threshold=1e-9)
    close(plan, g["plan_argmax"], rtol=1e-8, atol=1e-12)
    close((cost * plan).sum(), g["cost_argmax"], rtol=1e-9)


def test_gaussian_barycenter(golden):
    g = golden("barycenter")
    mean, cov, var, w = (T(g[k]) for k in ("mean", "cov", "var", "w"))
    for start in (0, 3):                                  # the fixed point does not depend on the start
        mb, cb = O.gaussian_barycenter(mean, cov, w, diag=False, n_iter=100, start=start)
        close(mb, g["mean_b"]); close(cb, g["cov_b"], rtol=1e-8, atol=1e-10)
    mb, vb = O.gaussian_barycenter(mean, var, w, diag=True)
    close(mb, g["mean_b_diag"]); close(vb, g["var_b_diag"])


def _fid_oracle_run(name):
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from fid_stub import CASES, FeatureNet, images
    fsize, n_gen, n_smp, batch = CASES[name]
    net, st = FeatureNet(fsize), O.FidStats(fsize)
    gen, smp = images(11, n_gen), images(12, n_smp, gain=0.85, offset=0.05)
    for lo in range(0, n_gen, batch):
        st.update(generated_features=net(gen[lo:lo + batch]))
    for lo in range(0, n_smp, batch):
        st.update(sample_features=net(smp[lo:lo + batch]))
    return st


def test_fid_statistics_and_score(golden):
    """metrics/fid.py:99-130 behind the torchmetrics stub (tests/golden/_load_reference.py): running sums, correlations,
    counts and the score, at feature_size 64 and 2048 (fewer and more observations than features)."""
    g = golden("fid")
    for name, tol in (("f64", 1e-10), ("f2048_deficient", 2e-6), ("f2048_full", 1e-9)):
        st = _fid_oracle_run(name)
        close(st.sum["real"], g[f"{name}_real_sum"]); close(st.sum["fake"], g[f"{name}_fake_sum"])
        assert st.n["real"].tolist() == g[f"{name}_num_real"].tolist() and st.n["fake"].tolist() == g[f"{name}_num_fake"].tolist()
        step = max(1, st.corr["real"].shape[0] // 32)
        close(st.corr["real"][::step, ::step], g[f"{name}_real_corr_sample"], rtol=1e-9, atol=1e-9)
        close(st.corr["fake"].trace(), g[f"{name}_fake_trace"], rtol=1e-9)
        # the singular case: sqrt(eigvals) of a non-symmetric singular product (reference side) vs eigvalsh of the
        # symmetric form (oracle) agree to ~1e-7 relative
        want = float(g[f"{name}_score"])
        assert abs(float(st.compute()) - want) <= tol * abs(want), (name, float(st.compute()), want)
    few = O.FidStats(64)
    few.update(torch.zeros(999, 64), torch.zeros(1200, 64))
    assert torch.isinf(few.compute()).all() and np.isinf(g["few_score"]).all()
