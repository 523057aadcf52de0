"""Unit tests of the hand-written sm_100a kernels against plain torch fp64 products (GPU only)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def relerr(got, want):
    return ((got.double() - want).norm() / want.norm()).item()


@pytest.mark.parametrize("nn", [False, True])
@pytest.mark.parametrize("M,N,K,batch", [(128, 128, 128, 1), (256, 128, 64, 1), (512, 512, 512, 2), (200, 136, 72, 3),
                                          (4096, 128, 128, 1), (1024, 1024, 1024, 1), (96, 64, 40, 1)])
def test_tcgen05_gemm_3xtf32_matches_fp64(nn, M, N, K, batch):
    from ot_vae_lightning_b200 import kernels as Kn
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    A = torch.randn(batch, M, K, device="cuda", generator=g)
    B = torch.randn(batch, K, N, device="cuda", generator=g) if nn else torch.randn(batch, N, K, device="cuda", generator=g)
    want = A.double() @ (B.double() if nn else B.double().transpose(-1, -2))
    simt = Kn.gemm(A, B, engine=1, nn=nn)
    assert relerr(simt, want) < 2e-6
    got3 = Kn.gemm(A, B, alpha=0.5, engine=2, nn=nn)
    # fp32-accurate; the TMEM accumulator truncates, so the error grows ~6e-8 per K=8 accumulation step
    assert relerr(got3, 0.5 * want) < 2e-6 + 1e-8 * K, relerr(got3, 0.5 * want)
    got1 = Kn.gemm(A, B, engine=3, nn=nn)
    e1 = relerr(got1, want)
    assert 1e-5 < e1 < 2e-3, e1                                                 # plain TF32: visibly coarser


def test_gemm_engine_rejects_ineligible_shapes():
    from ot_vae_lightning_b200 import kernels as Kn
    A = torch.randn(1, 30, 30, device="cuda")
    with pytest.raises(ValueError):
        Kn.gemm(A, A, engine=2)
    assert relerr(Kn.gemm(A, A, engine=0), A.double() @ A.double().transpose(-1, -2)) < 2e-6


@pytest.mark.parametrize("L,rows,d", [(1, 4096, 128), (1, 1000, 64), (2, 777, 192), (1, 70000, 512), (1, 33, 128),
                                      (3, 5000, 132), (1, 300000, 128)])
def test_tcgen05_stats_kernel_matches_fp64(L, rows, d):
    """sum x, sum x x^T, n through the tcgen05 SYRK (pivot-shifted, 3xTF32) vs fp64 torch; a large common offset
    (|mean| >> sigma) is added on purpose: the pivot shift must keep the covariance accurate."""
    from ot_vae_lightning_b200 import kernels as Kn
    g = torch.Generator(device="cuda").manual_seed(rows + d)
    x = torch.randn(L, rows, d, device="cuda", generator=g) * 0.3 + 5.0 + torch.randn(L, 1, d, device="cuda", generator=g)
    n = torch.zeros(L, dtype=torch.float64, device="cuda")
    s = torch.zeros(L, d, dtype=torch.float64, device="cuda")
    ss = torch.zeros(L, d, d, dtype=torch.float64, device="cuda")
    Kn.stats_update(x, n, s, ss, None)
    x64 = x.double()
    assert torch.equal(n, torch.full_like(n, rows))
    assert relerr(s, x64.sum(1)) < 1e-7
    assert relerr(ss, x64.transpose(1, 2) @ x64) < 1e-7
    assert float((ss - ss.transpose(1, 2)).abs().max()) == 0.0
    mean = s / rows
    cov = ss / rows - mean.unsqueeze(-1) * mean.unsqueeze(-2)
    xc = x64 - x64.mean(1, keepdim=True)
    assert relerr(cov, xc.transpose(1, 2) @ xc / rows) < 3e-5     # no cancellation blow-up despite |mean| ~ 17 sigma
    # second call accumulates (decay=None) and EMA works
    Kn.stats_update(x, n, s, ss, None)
    assert relerr(ss, 2 * (x64.transpose(1, 2) @ x64)) < 1e-7
    Kn.stats_update(x, n, s, ss, 0.75)
    assert relerr(ss, (0.75 * 2 + 0.25) * (x64.transpose(1, 2) @ x64)) < 1e-7 and relerr(n, torch.full_like(n, 1.75 * rows)) < 1e-12


@pytest.mark.parametrize("L,rows,d", [(1, 20000, 128), (5, 3000, 96), (150, 700, 64), (1, 131, 72), (1, 40000, 512),
                                      (3, 2500, 260), (1, 170000, 192), (10, 1500, 1024)])
def test_fp16_split_stats_kernel_scales_and_overflow_fallback(L, rows, d):
    """The FP16 hi/lo split kernels (stats_h.cu; one block for dim <= 128, the upper block triangle beyond, several
    launches for long inputs): per-feature power-of-two scales must absorb feature
    scales six decades apart, and a value outside the FP16 window (planted AFTER the head rows the scales are taken
    from) must raise the device flag and be recomputed by the gated TF32 kernel - same accuracy either way."""
    from ot_vae_lightning_b200 import kernels as Kn
    g = torch.Generator(device="cuda").manual_seed(7 * rows + d)
    feat = torch.logspace(-3, 3, d, device="cuda")
    x = torch.randn(L, rows, d, device="cuda", generator=g) * feat + 2.0 * feat
    for outlier in (False, True):
        if outlier:
            x[:, rows // 2 + 3, d // 3] = 3.0e8
        n = torch.zeros(L, dtype=torch.float64, device="cuda")
        s = torch.zeros(L, d, dtype=torch.float64, device="cuda")
        ss = torch.zeros(L, d, d, dtype=torch.float64, device="cuda")
        Kn.stats_update(x, n, s, ss, None)
        x64 = x.double()
        want = x64.transpose(1, 2) @ x64
        assert torch.equal(n, torch.full_like(n, rows))
        assert relerr(s, x64.sum(1)) < 1e-7
        # every entry is accurate relative to its own scale sqrt(S_ii S_jj), not only relative to the largest one
        scale = torch.sqrt(torch.diagonal(want, dim1=1, dim2=2).unsqueeze(-1) * torch.diagonal(want, dim1=1, dim2=2).unsqueeze(-2))
        assert float(((ss - want).abs() / scale).max()) < 5e-6
        assert float((ss - ss.transpose(1, 2)).abs().max()) == 0.0


@pytest.mark.parametrize("L,rows,d", [(1, 4096, 128), (1, 1000, 64), (2, 777, 192), (1, 65536, 512), (1, 5, 128), (2, 3000, 260)])
def test_tcgen05_apply_kernel_matches_fp64(L, rows, d):
    from ot_vae_lightning_b200 import kernels as Kn
    g = torch.Generator(device="cuda").manual_seed(rows * 3 + d)
    x = torch.randn(L, rows, d, device="cuda", generator=g) + 2.0
    T = torch.randn(L, d, d, device="cuda", generator=g, dtype=torch.float64) / d ** 0.5
    ms = torch.randn(L, d, device="cuda", generator=g, dtype=torch.float64) + 2.0
    mt = torch.randn(L, d, device="cuda", generator=g, dtype=torch.float64)
    y = Kn.apply_transport(x, ms, mt, T)
    want = (x.double() - ms.unsqueeze(1)) @ T.transpose(1, 2) + mt.unsqueeze(1)
    assert y.dtype == torch.float32 and relerr(y, want) < 2e-6 + 1e-8 * d
