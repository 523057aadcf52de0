"""CPU oracle for the latent optimal-transport path of theoad/ot-vae-lightning.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import it, and only as the checker or as
the timed CPU baseline.  The product (`ot-vae-lightning_b200/`) never imports this module and has no
CPU fallback.

It restates, in plain torch-on-CPU fp64 (the same LAPACK / MKL calls the reference itself bottoms out
in: `torch.linalg.eigh`, `einsum`, `logsumexp`), what the reference computes on the hot path.  Every
function cites the reference file:line (relative to /root/reference/ot_vae_lightning) it follows.

Parity is PINNED: `tests/golden/make_golden.py` imports the unmodified reference in the authoring
container and stores its outputs on seeded inputs under `tests/golden/*.npz`;
`tests/test_oracle_golden.py` checks every function below against those fixtures (fp64, rtol 1e-12).
The reference's own cross-checks that are runnable without POT (scipy Gelbrich distance,
`tests/test_w2_utils.py:113-122,179-195`; streaming-vs-one-shot covariance,
`tests/test_empirical_cov.py:47-72`) are re-stated in `tests/test_oracle_reference_tests.py`.
Sinkhorn-vs-POT (`tests/test_w2_utils.py:236-256`) needs the un-vendored `pot` package (unpinned in
`tests/requirements.txt:11`), which is absent here: Sinkhorn parity is anchored on the reference's own
`sinkhorn_log` outputs (fixtures) instead.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch import Tensor

STABILITY_CONST = 1e-8  # ot/matrix_utils.py:33


# ---------------------------------------------------------------------------------------------------
# L0: matrix primitives (ot/matrix_utils.py)
# ---------------------------------------------------------------------------------------------------

def spectral_apply(mats: Tensor, fn) -> Tensor:
    """V f(lambda) V^T with the lower triangle read by eigh.  ot/matrix_utils.py:37-46."""
    lam, vec = torch.linalg.eigh(mats, UPLO="L")
    return (vec * fn(lam).unsqueeze(-2)) @ vec.transpose(-1, -2)


def sqrtm(mats: Tensor) -> Tensor:
    """ot/matrix_utils.py:59-65."""
    return spectral_apply(mats, torch.sqrt)


def invsqrtm(mats: Tensor) -> Tensor:
    """ot/matrix_utils.py:68-76."""
    return spectral_apply(mats, lambda lam: 1.0 / torch.sqrt(lam))


def identity_like(mats: Tensor) -> Tensor:
    """ot/matrix_utils.py:49-56."""
    d = mats.shape[-1]
    return torch.eye(d, dtype=mats.dtype).expand(mats.shape)


def is_symmetric(mats: Tensor) -> Tensor:
    """sum((A - A^T)^2) < 1e-8.  ot/matrix_utils.py:79-88."""
    if mats.shape[-1] != mats.shape[-2]:
        return torch.zeros(mats.shape[:-2], dtype=torch.bool)
    skew = mats - mats.transpose(-1, -2)
    return (skew * skew).sum(dim=(-1, -2)) < STABILITY_CONST


def min_eig(mats: Tensor) -> Tensor:
    """ot/matrix_utils.py:91-98."""
    return torch.linalg.eigvalsh(mats).amin(dim=-1)


def is_pd(mats: Tensor, strict: bool = True) -> Tensor:
    """ot/matrix_utils.py:101-109."""
    lam = min_eig(mats)
    return lam > 0 if strict else lam >= 0


def make_psd(mats: Tensor, strict: bool = False) -> Tuple[Tensor, Tensor]:
    """A + (max(0, -lambda_min) [+ 1e-8 if strict]) I; returns (repaired, shift).  ot/matrix_utils.py:123-142."""
    shift = min_eig(mats).clamp(max=0).abs()
    if strict:
        shift = shift + STABILITY_CONST
    return mats + identity_like(mats) * shift[..., None, None], shift


def mean_cov(sum_x: Tensor, sum_xx: Tensor, n_obs: Tensor) -> Tuple[Tensor, Tensor]:
    """mean = Sx/n, cov = Sxx/n - mean mean^T (biased).  ot/matrix_utils.py:145-158 (full-matrix branch)."""
    n_vec = n_obs.reshape(n_obs.shape + (1,) * (sum_x.dim() - n_obs.dim()))
    n_mat = n_obs.reshape(n_obs.shape + (1,) * (sum_xx.dim() - n_obs.dim()))
    mean = sum_x / n_vec
    cov = sum_xx / n_mat - mean.unsqueeze(-1) * mean.unsqueeze(-2)
    return mean, cov


# ---------------------------------------------------------------------------------------------------
# streaming sufficient statistics (ot/distribution_models/gaussian_model.py, metrics/fid.py)
# ---------------------------------------------------------------------------------------------------

def batch_stats(samples: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
    """n = B, Sx = sum_b x_b, Sxx = sum_b x_b x_b^T.  gaussian_model.py:144-148 (and fid.py:102-104)."""
    n = torch.as_tensor(float(samples.shape[-2]), dtype=samples.dtype)
    return n, samples.sum(-2), torch.einsum("...bi,...bj->...ij", samples, samples)


def accumulate(running: Tensor, new: Tensor, decay: Optional[float]) -> Tensor:
    """utils/__init__.py:204-206: plain sum, or old*decay + new*(1-decay)."""
    if decay is None:
        return running + new
    return running * decay + new * (1.0 - decay)


class GaussianStats:
    """The running buffers of `GaussianModel` (gaussian_model.py:57-63, 99-108) for leading shape `lead`."""

    def __init__(self, *size: int, decay: Optional[float] = None, dtype=torch.double):
        lead, d = tuple(size[:-1]), size[-1]
        self.decay = decay
        self.n_obs = torch.zeros(lead, dtype=torch.double)
        self.sum = torch.zeros(*lead, d, dtype=dtype)
        self.sum_cov = torch.zeros(*lead, d, d, dtype=dtype)

    def update(self, samples: Tensor) -> None:
        x = samples.to(self.sum.dtype)  # gaussian_model.py:103
        n, s, ss = batch_stats(x)
        self.n_obs = accumulate(self.n_obs, n.to(self.n_obs.dtype), self.decay)
        self.sum = accumulate(self.sum, s, self.decay)
        self.sum_cov = accumulate(self.sum_cov, ss, self.decay)

    def fit(self) -> Tuple[Tensor, Tensor]:
        """mean and the covariance *as read through the parametrizations* (gaussian_model.py:110-126, 204-229)."""
        mean, cov = mean_cov(self.sum, self.sum_cov, self.n_obs.to(self.sum.dtype))
        return mean, parametrized_cov(cov)


def parametrized_cov(raw: Tensor) -> Tensor:
    """What `.cov` returns: triu mirrored (Symmetric, gaussian_model.py:220-229) then
    make_psd(strict=True) (MakePositiveDefinite, gaussian_model.py:204-217)."""
    sym = raw.triu() + raw.triu(1).transpose(-1, -2)
    return make_psd(sym, strict=True)[0]


# ---------------------------------------------------------------------------------------------------
# L1: Gaussian W2 machinery (ot/w2_utils.py)
# ---------------------------------------------------------------------------------------------------

def w2_gaussian(mean_s: Tensor, mean_t: Tensor, cov_s: Tensor, cov_t: Tensor) -> Tensor:
    """|mu_s - mu_t|^2 + tr(Cs + Ct - 2 (Ct^1/2 Cs Ct^1/2)^1/2).  ot/w2_utils.py:70-80 (inputs already valid)."""
    mean_s, mean_t, cov_s, cov_t = (t.double() for t in (mean_s, mean_t, cov_s, cov_t))
    root_t = sqrtm(cov_t)
    mix = root_t @ cov_s @ root_t
    shift = ((mean_s - mean_t) ** 2).sum(-1)
    tr = torch.diagonal(cov_s + cov_t - 2.0 * sqrtm(mix), dim1=-2, dim2=-1).sum(-1)
    return shift + tr


def transport_operator_full(cov_s: Tensor, cov_t: Tensor, pg_star: float = 0.0) -> Tuple[Tensor, Tensor]:
    """T = (1-p) Cs^-1/2 (Cs^1/2 Ct Cs^1/2)^1/2 Cs^-1/2 + p I ; Cw = 0.  ot/w2_utils.py:756-768."""
    cov_s, cov_t = cov_s.double(), cov_t.double()
    eye = identity_like(cov_s)
    root = sqrtm(cov_s)
    iroot = invsqrtm(cov_s + STABILITY_CONST * eye)
    T = (1.0 - pg_star) * (iroot @ sqrtm(root @ cov_t @ root) @ iroot) + pg_star * eye
    return T, torch.zeros_like(T)


def apply_transport(x: Tensor, mean_s: Tensor, mean_t: Tensor, T: Tensor) -> Tensor:
    """y = T (x - mu_s) + mu_t per latent, fp64.  ot/w2_utils.py:517-520 (Cw ignored: :507).
    Shapes as `W2Mixin.apply_transport` builds them (:581-597): x [*L,B,d], mean [*L,d], T [*L,d,d]."""
    x64 = x.double()
    centered = x64 - mean_s.double().unsqueeze(-2)
    moved = (T.double().unsqueeze(-3) @ centered.unsqueeze(-1)).squeeze(-1)
    return moved + mean_t.double().unsqueeze(-2)


def gaussian_transport_pipeline(src: Tensor, tgt: Tensor, batch: int, decay: Optional[float] = None,
                                pg_star: float = 0.0):
    """What `GaussianTransport.update* -> compute -> transport` does end to end
    (transport/gaussian_transport.py:64-95, transport/base.py:107-149) with fp64 buffers."""
    d = src.shape[-1]
    lead = src.shape[:-2]
    s_stats, t_stats = GaussianStats(*lead, d, decay=decay), GaussianStats(*lead, d, decay=decay)
    for lo in range(0, src.shape[-2], batch):
        s_stats.update(src[..., lo:lo + batch, :])
    for lo in range(0, tgt.shape[-2], batch):
        t_stats.update(tgt[..., lo:lo + batch, :])
    mean_s, cov_s = s_stats.fit()
    mean_t, cov_t = t_stats.fit()
    w2 = w2_gaussian(mean_s, mean_t, cov_s, cov_t)
    T, _ = transport_operator_full(cov_s, cov_t, pg_star)
    moved = apply_transport(src, mean_s, mean_t, T).to(src.dtype)  # gaussian_transport.py:95
    return dict(mean_s=mean_s, cov_s=cov_s, mean_t=mean_t, cov_t=cov_t, w2=w2, T=T, moved=moved)


# ---------------------------------------------------------------------------------------------------
# operator variants (ot/w2_utils.py:714-793) and the Gaussian-mixture layer built on the same pieces
# (ot/w2_utils.py:86-270, distribution_models/gassian_mixture_model.py, transport/gmm_transport.py)
# ---------------------------------------------------------------------------------------------------

def transport_operator_diag(var_s: Tensor, var_t: Tensor, pg_star: float = 0.0) -> Tuple[Tensor, Tensor]:
    """T = (1-p) sqrt(vt / vs + 1e-8) + p ; Cw = 0.  ot/w2_utils.py:714-721."""
    var_s, var_t = var_s.double(), var_t.double()
    T = (1.0 - pg_star) * torch.sqrt(var_t / var_s + STABILITY_CONST) + pg_star
    return T, torch.zeros_like(T)


def transport_operator_diag_stochastic(var_s: Tensor, var_t: Tensor, pg_star: float = 0.0) -> Tuple[Tensor, Tensor]:
    """Diagonal form of eq. 19: T = (1-p) sqrt(vt vs) pinv(vs) + p, Cw = sqrt(1-p) vt (1 - vt pinv(vs) T*^2) with
    T* = sqrt(vs / vt + 1e-8) and the thresholded pseudo-inverse of the source variances.  ot/w2_utils.py:729-750."""
    var_s, var_t = var_s.double(), var_t.double()
    T_star = torch.sqrt(var_s / var_t + STABILITY_CONST)
    pinv_s = torch.where(var_s > STABILITY_CONST, 1.0 / var_s, torch.zeros_like(var_s))
    T = (1.0 - pg_star) * torch.sqrt(var_t * var_s) * pinv_s + pg_star
    Cw = (1.0 - pg_star) ** 0.5 * var_t * (1.0 - var_t * pinv_s * T_star ** 2)
    return T, Cw


def transport_operator_full_stochastic(cov_s: Tensor, cov_t: Tensor, pg_star: float = 0.0) -> Tuple[Tensor, Tensor]:
    """Eq. 19 of Freirich et al. with the pseudo-inverse of the source covariance.  ot/w2_utils.py:774-793."""
    cov_s, cov_t = cov_s.double(), cov_t.double()
    eye = identity_like(cov_s)
    pinv_s = torch.linalg.pinv(cov_s)
    root_t = sqrtm(cov_t)
    iroot_t = invsqrtm(cov_t + STABILITY_CONST * eye)
    T_star, _ = transport_operator_full(cov_t, cov_s, 0.0)
    T = (1.0 - pg_star) * (root_t @ sqrtm(root_t @ cov_s @ root_t) @ iroot_t @ pinv_s) + pg_star * eye
    Cw = (1.0 - pg_star) ** 0.5 * root_t @ (eye - root_t @ T_star @ pinv_s @ T_star @ root_t) @ root_t
    return T, Cw


def w2_dissimilarity(mean_s: Tensor, mean_t: Tensor, var_s: Tensor, var_t: Tensor, diag: bool) -> Tensor:
    """All-pairs Gaussian W2^2 between N source and M target components, [*, N, M].
    ot/w2_utils.py:86-134 (diagonal: |ms-mt|^2 + |sqrt(vs)-sqrt(vt)|^2) and :140-191 (full)."""
    mean_s, mean_t, var_s, var_t = (t.double() for t in (mean_s, mean_t, var_s, var_t))
    if diag:
        dm = ((mean_s.unsqueeze(-2) - mean_t.unsqueeze(-3)) ** 2).sum(-1)
        dv = ((var_s.sqrt().unsqueeze(-2) - var_t.sqrt().unsqueeze(-3)) ** 2).sum(-1)
        return dm + dv
    n, m = mean_s.shape[-2], mean_t.shape[-2]
    lead = mean_s.shape[:-2]
    return w2_gaussian(mean_s.unsqueeze(-2).expand(*lead, n, m, -1), mean_t.unsqueeze(-3).expand(*lead, n, m, -1),
                       var_s.unsqueeze(-3).expand(*lead, n, m, -1, -1), var_t.unsqueeze(-4).expand(*lead, n, m, -1, -1))


def ot_gmm(mean_s: Tensor, mean_t: Tensor, var_s: Tensor, var_t: Tensor, w_s: Tensor, w_t: Tensor, diag: bool,
           **sinkhorn_kwargs) -> Tuple[Tensor, Tensor]:
    """Entropic OT between the components: cost normalised by its max for the solve, total reported with the
    un-normalised cost.  ot/w2_utils.py:197-270."""
    cost = w2_dissimilarity(mean_s, mean_t, var_s, var_t, diag)
    plan = sinkhorn_log(w_s.double(), w_t.double(), cost / cost.amax(dim=(-2, -1), keepdim=True), **sinkhorn_kwargs)
    return (cost * plan).sum(dim=(-2, -1)), plan


def gaussian_barycenter(mean: Tensor, cov: Tensor, weights: Tensor, diag: bool, n_iter: int = 100, start: int = 0
                        ) -> Tuple[Tensor, Tensor]:
    """W2 barycenter of N(mean_i, cov_i) with weights w_i: mean = sum w_i mean_i; variances (sum w_i sqrt(v_i))^2 when
    `diag`, else the fixed point S <- sum_i w_i (S^1/2 C_i S^1/2)^1/2 started from C_start (the reference draws the start
    at random; the fixed point does not depend on it).  ot/w2_utils.py:325-385."""
    mean, cov, weights = mean.double(), cov.double(), weights.double()
    mean_b = (weights.unsqueeze(-2) @ mean).squeeze(-2)
    if diag:
        return mean_b, ((weights.unsqueeze(-2) @ cov.sqrt()) ** 2).squeeze(-2)
    w = weights[..., None, None]
    cov_b = cov.select(-3, start).unsqueeze(-3)
    for _ in range(n_iter):
        root = sqrtm(cov_b)
        cov_b = (w * sqrtm(root @ cov @ root)).sum(-3, keepdim=True)
    return mean_b, cov_b.squeeze(-3)


def gmm_energy(x: Tensor, mean: Tensor, var: Tensor, weights: Tensor, diag: bool) -> Tensor:
    """log N(x_b | mean_k, var_k) + log w_k, [*, B, K].  gassian_mixture_model.py:86-94."""
    x, mean, var, weights = (t.double() for t in (x, mean, var, weights))
    d = x.shape[-1]
    diff = x.unsqueeze(-2) - mean.unsqueeze(-3)                                  # [*, B, K, d]
    if diag:
        maha = (diff * diff / var.unsqueeze(-3)).sum(-1)
        logdet = var.log().sum(-1).unsqueeze(-2)
    else:
        chol = torch.linalg.cholesky(var)                                        # [*, K, d, d]
        sol = torch.linalg.solve_triangular(chol.unsqueeze(-4), diff.unsqueeze(-1), upper=False).squeeze(-1)
        maha = (sol * sol).sum(-1)
        logdet = 2.0 * torch.diagonal(chol, dim1=-2, dim2=-1).log().sum(-1).unsqueeze(-2)
    log_prob = -0.5 * (maha + logdet + d * torch.log(torch.tensor(2.0 * torch.pi, dtype=torch.double)))
    return log_prob + torch.log_softmax(weights.log(), dim=-1).unsqueeze(-2)


def gmm_weighted_stats(x: Tensor, weights: Tensor, diag: bool) -> Tuple[Tensor, Tensor, Tensor]:
    """(sum_b w_bk, sum_b w_bk x_b, sum_b w_bk x_b x_b^T) per component, the dense expression of
    gassian_mixture_model.py:104-117 (it materialises the B x d^2 outer products)."""
    x, weights = x.double(), weights.double()
    wt = weights.transpose(-1, -2)
    if diag:
        return weights.sum(-2), wt @ x, wt @ (x * x)
    outer = (x.unsqueeze(-1) @ x.unsqueeze(-2)).flatten(-2)
    return weights.sum(-2), wt @ x, (wt @ outer).unflatten(-1, (x.shape[-1], x.shape[-1]))


def gmm_transport_hard(x: Tensor, src_idx: Tensor, tgt_idx: Tensor, mean_s: Tensor, mean_t: Tensor, cov_s: Tensor,
                       cov_t: Tensor, diag: bool) -> Tensor:
    """Every input goes through the Gaussian map between the source component it is assigned to and its target
    component (hard assignments).  gmm_transport.py:82-121 with one-hot assignments."""
    x64 = x.double()
    ms, mt = mean_s.double()[src_idx], mean_t.double()[tgt_idx]
    if diag:
        T, _ = transport_operator_diag(cov_s[src_idx], cov_t[tgt_idx])
        return T * (x64 - ms) + mt
    T, _ = transport_operator_full(cov_s[src_idx], cov_t[tgt_idx])
    return (T @ (x64 - ms).unsqueeze(-1)).squeeze(-1) + mt


# ---------------------------------------------------------------------------------------------------
# log-domain Sinkhorn (ot/w2_utils.py:276-319) and its cost producers
# ---------------------------------------------------------------------------------------------------

def sinkhorn_log(a: Tensor, b: Tensor, C: Tensor, reg: float = 1e-5, max_iter: int = 1000,
                 threshold: float = STABILITY_CONST, return_potentials: bool = False):
    """v first, then u; stop when the MIN over batch elements of sum|du|+sum|dv| < threshold
    (checked after each full iteration); plan = exp(u_i + v_j - C_ij/reg).  ot/w2_utils.py:301-319."""
    log_a = torch.log(a + STABILITY_CONST)
    log_b = torch.log(b + STABILITY_CONST)
    neg_c = -C / reg
    u = torch.zeros_like(a)
    v = torch.zeros_like(b)
    n_done = 0
    for _ in range(max_iter):
        u_prev, v_prev = u, v
        v = log_b - torch.logsumexp(neg_c + u.unsqueeze(-1), dim=-2)
        u = log_a - torch.logsumexp(neg_c + v.unsqueeze(-2), dim=-1)
        n_done += 1
        moved = (u - u_prev).abs().sum(-1) + (v - v_prev).abs().sum(-1)
        if moved.min().item() < threshold:
            break
    plan = torch.exp(u.unsqueeze(-1) + v.unsqueeze(-2) + neg_c)
    if return_potentials:
        return plan, u, v, n_done
    return plan


def sqeuclidean_cost(x: Tensor, y: Tensor) -> Tensor:
    """|x_i|^2 + |y_j|^2 - 2 x_i.y_j, the mean part of ot/w2_utils.py:121-125."""
    return (x * x).sum(-1, keepdim=True) + (y * y).sum(-1).unsqueeze(-2) - 2.0 * (x @ y.transpose(-1, -2))


def inverse_distance_energy(x: Tensor, codebook: Tensor, p: float = 2.0) -> Tensor:
    """1 / (cdist_p(x, c) + 1e-8): `CodebookModel.energy`, codebook_model.py:155-160 (euclidean metric)."""
    return 1.0 / (torch.cdist(x, codebook, p) + 1e-8)


def sinkhorn_summary(a: Tensor, b: Tensor, C: Tensor, plan: Tensor):
    """row marginals, column marginals, <C, plan> (discrete_transport.py:67, w2_utils.py:269)."""
    return plan.sum(-1), plan.sum(-2), (C * plan).sum(dim=(-1, -2))


# ---------------------------------------------------------------------------------------------------
# synthetic inputs shared by the parity tests and the CPU baseline
# ---------------------------------------------------------------------------------------------------

def synthetic_gaussian(n: int, d: int, seed: int, kappa: float = 1e2, dtype=torch.float32):
    """mu ~ N(0,1)^d, Sigma = Q diag(lambda) Q^T with lambda log-spaced in [1/kappa, 1]; x = mu + L z."""
    g = torch.Generator().manual_seed(seed)
    mu = torch.randn(d, generator=g, dtype=torch.double)
    q, _ = torch.linalg.qr(torch.randn(d, d, generator=g, dtype=torch.double))
    lam = torch.logspace(-torch.log10(torch.tensor(kappa)).item(), 0.0, d, dtype=torch.double)
    half = q * lam.sqrt()
    z = torch.randn(n, d, generator=g, dtype=torch.double)
    x = mu + z @ half.T
    return x.to(dtype), mu, (half @ half.T)


# ---------------------------------------------------------------------------------------------------
# FID accumulation + score (metrics/fid.py:99-130)
# ---------------------------------------------------------------------------------------------------

def frechet_distance(mean_1: Tensor, cov_1: Tensor, mean_2: Tensor, cov_2: Tensor) -> Tensor:
    """`_compute_fid` of torchmetrics (third-party, `torchmetrics>=0.9.2` UNPINNED in the reference's requirements.txt:4,
    absent from /root/reference and from this image; call site metrics/fid.py:130).  Published algorithm
    (torchmetrics/image/fid.py, v1.x): |m1-m2|^2 + tr(C1) + tr(C2) - 2 * sum(sqrt(eigvals(C1 C2))).real.
    C1 C2 is similar to the symmetric PSD matrix C1^1/2 C2 C1^1/2, whose eigenvalues are evaluated here with `eigh`
    (negative round-off eigenvalues contribute nothing, as `.sqrt().real` drops them there)."""
    root = spectral_apply(cov_1, lambda lam: torch.sqrt(lam.clamp_min(0.0)))
    mid = root @ cov_2 @ root
    lam = torch.linalg.eigvalsh((mid + mid.transpose(-1, -2)) / 2).clamp_min(0.0)
    diff = mean_1 - mean_2
    return (diff * diff).sum(-1) + torch.diagonal(cov_1, dim1=-2, dim2=-1).sum(-1) \
        + torch.diagonal(cov_2, dim1=-2, dim2=-1).sum(-1) - 2.0 * lam.sqrt().sum(-1)


class FidStats:
    """States and update rule of the reference's `FrechetInceptionDistance` (metrics/fid.py:88-122): fp64 feature sums
    and `features.T @ features`, int64 counts of shape [1]; NB `generated` feeds the `real_*` states (:113-117)."""

    def __init__(self, feature_size: int):
        self.sum = {k: torch.zeros(feature_size, dtype=torch.double) for k in ("real", "fake")}
        self.corr = {k: torch.zeros(feature_size, feature_size, dtype=torch.double) for k in ("real", "fake")}
        self.n = {k: torch.zeros(1, dtype=torch.long) for k in ("real", "fake")}

    def update(self, generated_features: Optional[Tensor] = None, sample_features: Optional[Tensor] = None) -> None:
        for kind, feats in (("real", generated_features), ("fake", sample_features)):
            if feats is None:
                continue
            f = feats.double().reshape(feats.shape[0], -1)          # fid.py:102
            self.sum[kind] += f.sum(dim=0)                          # fid.py:103
            self.corr[kind] += f.T @ f                              # fid.py:104
            self.n[kind] += feats.shape[0]

    def compute(self) -> Tensor:
        """fid.py:124-130"""
        if self.n["fake"] < 1e3 or self.n["real"] < 1000:
            return torch.ones(1) * float("inf")
        m_r, c_r = mean_cov(self.sum["real"], self.corr["real"], self.n["real"].double())
        m_f, c_f = mean_cov(self.sum["fake"], self.corr["fake"], self.n["fake"].double())
        return frechet_distance(m_r, c_r, m_f, c_f)


# ---------------------------------------------------------------------------------------------------
# codebook k-means step (distribution_models/base.py:206-253, codebook_model.py:155-160)
# ---------------------------------------------------------------------------------------------------

def kmeans_step(samples: Tensor, codebook: Tensor, temperature: float = 1.0, mode: str = "argmax"
                ) -> Tuple[Tensor, Tensor, Tensor]:
    """`MixtureMixin.assign` + `kmean_iteration` with the Euclidean energy: energy = 1/(cdist + 1e-8)
    (codebook_model.py:159-160), weights = softmax(energy / T) (base.py:220), hardened to one-hot rows of their argmax in
    'argmax' mode (:229-230); returns (weights.sum(-2), weights^T @ samples, argmax index) (:249-253)."""
    samples, codebook = samples.double(), codebook.double()
    energy = inverse_distance_energy(samples, codebook)
    weights = torch.softmax(energy / temperature, dim=-1)
    index = weights.argmax(-1)
    if mode == "argmax":
        weights = torch.nn.functional.one_hot(index, energy.size(-1)).to(weights.dtype)
    return weights.sum(-2), weights.transpose(-1, -2) @ samples, index
