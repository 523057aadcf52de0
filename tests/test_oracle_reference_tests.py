"""The reference's own runnable checks (those that need neither POT nor a GPU), restated on oracle/ot_oracle.py:
property tests (`W2(x, x) = 0`, tests/test_w2_utils.py:34-83), the scipy cross-check of the Gelbrich distance
(tests/test_w2_utils.py:106-122) and streaming-vs-one-shot covariance (tests/test_empirical_cov.py:40-72).
Same shapes, seeds-by-generator instead of the reference's `@retry`."""
import numpy as np
import scipy.linalg as spl
import torch

from oracle import ot_oracle as O

DIM = 3                       # tests/test_w2_utils.py:23
EPS = O.STABILITY_CONST


def rand_mean_cov(gen, lead, dim, dtype=torch.double):
    """tests/test_w2_utils.py:25-33"""
    mean = torch.randn(*lead, dim, generator=gen, dtype=dtype)
    root = torch.randn(*lead, dim, dim, generator=gen, dtype=dtype)
    return mean, root @ root.transpose(-1, -2) + torch.eye(dim, dtype=dtype) * 1e-5


def test_w2_gaussian_same_yields_0():
    mean, cov = rand_mean_cov(torch.Generator().manual_seed(1), (2, 3), DIM)
    res = O.w2_gaussian(mean, mean, cov, cov)
    assert res.shape == torch.Size([2, 3])
    assert torch.allclose(res, torch.zeros_like(res), atol=1e-7)       # reference: 1e-8 * DIM on its fp64 eigh path


def test_batch_w2_same_yields_0_on_the_diagonal():
    mean, cov = rand_mean_cov(torch.Generator().manual_seed(2), (2, 3), DIM)
    var = torch.diagonal(cov, dim1=-1, dim2=-2)
    res = O.w2_dissimilarity(mean, mean, var, var, diag=True)
    assert res.shape == torch.Size([2, 3, 3])
    assert torch.allclose(torch.diagonal(res, dim1=-1, dim2=-2), torch.zeros(2, 3, dtype=torch.double), atol=1e-12)
    full = O.w2_dissimilarity(mean, mean, cov, cov, diag=False)
    assert torch.allclose(torch.diagonal(full, dim1=-1, dim2=-2), torch.zeros(2, 3, dtype=torch.double), atol=1e-7)


def test_ot_gmm_same_yields_0():
    """tests/test_w2_utils.py:62-83 (default reg = 1e-5: the plan is the identity coupling up to exp(-1/reg))"""
    mean, cov = rand_mean_cov(torch.Generator().manual_seed(3), (2, 3), DIM)
    var = torch.diagonal(cov, dim1=-1, dim2=-2)
    w = torch.full((2, 3), 1.0 / 3, dtype=torch.double)
    res, plan = O.ot_gmm(mean, mean, var, var, w, w, diag=True)
    assert res.shape == torch.Size([2]) and torch.allclose(res, torch.zeros_like(res), atol=1e-6)
    res, _ = O.ot_gmm(mean, mean, cov, cov, w, w, diag=False)
    assert torch.allclose(res, torch.zeros_like(res), atol=1e-6)


def test_w2_gaussian_vs_scipy():
    """tests/test_w2_utils.py:106-122: Julie Delon's scipy formula as the independent implementation"""
    g = torch.Generator().manual_seed(4)
    for _ in range(5):
        m0, c0 = rand_mean_cov(g, (), 6)
        m1, c1 = rand_mean_cov(g, (), 6)
        root0 = spl.sqrtm(c0.numpy())
        want = np.linalg.norm(m0.numpy() - m1.numpy()) ** 2 + np.trace(
            c0.numpy() + c1.numpy() - 2 * spl.sqrtm(root0 @ c1.numpy() @ root0))
        got = O.w2_gaussian(m0, m1, c0, c1).item()
        assert abs(got - np.real(want)) < 1e-6 * max(1.0, abs(want))


def test_streaming_statistics_equal_one_shot():
    """tests/test_empirical_cov.py:40-72: batched accumulation of (n, sum x, sum x x^T) + mean_cov reproduces the
    one-shot empirical mean / covariance to 1e-8 (relative) and W2 between the two to sqrt(1e-8)."""
    g = torch.Generator().manual_seed(5)
    d, n, batch = 16, 10000, 250
    mix = torch.randn(d, d, generator=g, dtype=torch.double) / d ** 0.5
    z = torch.randn(n, d, generator=g, dtype=torch.double) @ mix.T + torch.randn(d, generator=g, dtype=torch.double)
    mean_all = z.mean(0)
    cov_all = (z - mean_all).T @ (z - mean_all) / n
    st = O.GaussianStats(d)
    for lo in range(0, n, batch):
        st.update(z[lo:lo + batch])
    mean, cov = O.mean_cov(st.sum, st.sum_cov, st.n_obs)
    assert ((mean - mean_all).norm() / mean_all.norm()).item() < EPS
    assert ((cov - cov_all).norm() / cov_all.norm()).item() < EPS
    ridge = 1e-8 * torch.eye(d, dtype=torch.double)
    assert O.w2_gaussian(mean_all, mean, cov_all + ridge, cov + ridge).abs().item() < EPS ** 0.5


def test_gaussian_barycenter_of_identical_components_is_the_component():
    """tests/test_w2_utils.py:85-104"""
    g = torch.Generator().manual_seed(6)
    mean, cov = rand_mean_cov(g, (2, 1), DIM)
    mean, cov = mean.repeat(1, 3, 1), cov.repeat(1, 3, 1, 1)
    var = torch.diagonal(cov, dim1=-1, dim2=-2)
    w = torch.randn(2, 3, generator=g, dtype=torch.double).abs()
    w = w / w.sum(-1, keepdim=True)
    mb, vb = O.gaussian_barycenter(mean, var, w, diag=True)
    assert torch.allclose(mb, mean[:, 0]) and torch.allclose(vb, var[:, 0])
    mb, cb = O.gaussian_barycenter(mean, cov, w, diag=False)
    assert torch.allclose(mb, mean[:, 0]) and torch.allclose(cb, cov[:, 0], atol=1e-7)
