"""Gaussian W2 machinery and log-domain Sinkhorn on B200, behind the reference's function names and
signatures (`ot_vae_lightning/ot/w2_utils.py:26-36`; like the reference, this module re-exports
`matrix_utils` so that `from ...w2_utils import mean_cov` keeps working, reference :24).

What runs where
  w2_gaussian / compute_transport_operators (full, deterministic) / apply_transport (full) / sinkhorn_log
      -> libotk kernels (kernels.py).  Math results are returned in `dtype` (fp64 by default), on the
         device of the inputs; internally the matrix functions are fp32-accurate (3xTF32 + fp32 accumulate).
  diagonal variants, GMM costs, barycenter, stochastic operator (SURVEY 8f "next" rows)
      -> short torch expressions on the same device around the same kernels.
Argument validation follows the reference's rules and raises ValueError under the same conditions
(reference :605-708).
"""
from __future__ import annotations

import warnings
from functools import partial
from math import sqrt
from typing import Optional, Tuple, Union

import torch
import torch.distributions as D
from torch import Tensor
from torch.types import _dtype

from .. import kernels as K
from .matrix_utils import *  # noqa: F401,F403  (re-export, as the reference does)
from .matrix_utils import STABILITY_CONST, eye_like, invsqrtm, is_pd, is_spd, is_symmetric, make_psd, mean_cov, sqrtm

__all__ = [
    "w2_gaussian", "batch_w2_dissimilarity_gaussian_diag", "batch_w2_dissimilarity_gaussian", "batch_ot_gmm",
    "sinkhorn_log", "gaussian_barycenter", "compute_transport_operators", "apply_transport", "W2Mixin",
]

_VECTOR_KINDS = ("vec", "var")
_MATRIX_KINDS = ("mat", "spd", "spsd", "pd", "psd")


def _validate_args(*spec, make_pd: bool = False, verbose: bool = False, dtype: Optional[_dtype] = None,
                   tol: float = 1e-5) -> Tuple[Tensor, ...]:
    """`spec` is a flat sequence of (tensor, name, kind) triples, kind in
    vec | var | prob | mat | spd | spsd | pd | psd.  Same acceptance rules as reference w2_utils.py:605-708:
    rank, sign, normalisation, symmetry and definiteness checks, optional repair (`make_pd`), cast to `dtype`,
    and matching feature / component / broadcastable batch dimensions."""
    with_components = "prob" in spec
    extra = int(with_components)
    feature_dims, component_dims, batch_shapes, out = [], [], [], []

    def consistent(values) -> bool:
        return len(set(values)) <= 1

    for arg, name, kind in zip(spec[0::3], spec[1::3], spec[2::3]):
        if not isinstance(arg, Tensor):
            raise ValueError(f"`{name}` is expected to be a torch.Tensor, got `{type(arg)}` instead.")
        if kind in _VECTOR_KINDS:
            if arg.dim() < 1 + extra:
                raise ValueError(f"`{name}` should be 1-dim vectors"
                                 f"{' with a leading component dimension' if with_components else ''} "
                                 f"(+ optional leading batch dimensions), got `{name}.dim()={arg.dim()}`")
            if kind == "var" and bool((arg < 0).any()):
                raise ValueError(f"`{name}` is expected to be a valid variance vector with positive entries.")
            feature_dims.append(arg.size(-1))
            if with_components:
                component_dims.append(arg.size(-2))
            batch_shapes.append(tuple(arg.shape[:-1 - extra]))
        elif kind == "prob":
            if arg.dim() < 1:
                raise ValueError(f"`{name}` should be a 1-dim vectors (+ optional leading batch dimensions), "
                                 f"got `{name}.dim()={arg.dim()}`")
            total = arg.sum(-1)
            if bool((arg < -tol).any()) or bool((total < 1 - tol).any()) or bool((total > 1 + tol).any()):
                raise ValueError(f"`{name}` is expected to be a valid probability vector with positive entries "
                                 f"that sum up to 1.")
            component_dims.append(arg.size(-1))
            batch_shapes.append(tuple(arg.shape[:-1]))
        elif kind in _MATRIX_KINDS:
            if arg.dim() < 2 + extra:
                raise ValueError(f"`{name}` should be a 2-dim matrix"
                                 f"{' with a leading component dimension' if with_components else ''} "
                                 f"(+ optional leading batch dimensions), got `{name}.dim()`={arg.dim()}.")
            wants_symmetry = kind.startswith("s")
            core = kind[1:] if wants_symmetry else kind          # 'pd' | 'psd' | 'at'(from 'mat')
            if wants_symmetry:
                sym = is_symmetric(arg)
                if not bool(sym.all()):
                    raise ValueError(f"`{name}` should be symmetric. Found {int((~sym).sum())} non-symmetric matrices.")
            if "pd" in kind:
                strict = "s" not in core                        # 'pd' strict, 'psd' semi-definite
                semi = "" if strict else "semi "
                if not bool(is_pd(arg, strict=strict).all()):
                    if not make_pd:
                        raise ValueError(f"`{name}` should be symmetric and positive {semi}definite. "
                                         f"Use `make_pd=True` to automatically add a small value to the matrix diagonals.")
                    arg, added = make_psd(arg, strict=strict, return_correction=True)
                    if verbose:
                        warnings.warn(f"`{name}` is not positive {semi}definite. Adding a small value to the diagonal "
                                      f"(<{added.max().item():.2e}) to ensure the matrices are positive {semi}definite")
            feature_dims.extend([arg.size(-2), arg.size(-1)])
            if with_components:
                component_dims.append(arg.size(-3))
            batch_shapes.append(tuple(arg.shape[:-2 - extra]))
        else:
            raise AssertionError(f"unknown argument kind {kind!r}")
        out.append(arg.to(dtype) if dtype is not None else arg)

        if not consistent(feature_dims):
            raise ValueError(f"All the inputs dimensionalities should match, got {feature_dims}")
        if not consistent(batch_shapes) and tuple(torch.broadcast_shapes(*batch_shapes)) not in batch_shapes:
            raise ValueError(f"All the inputs leading batch dimensions should be broadcastable, got {batch_shapes}")
        if not consistent(component_dims):
            raise ValueError(f"All the inputs component dimension should match, got {component_dims}")
    return tuple(out)


# ---------------------------------------------------------------------------------------------------------------------

def w2_gaussian(mean_source: Tensor, mean_target: Tensor, cov_source: Tensor, cov_target: Tensor,
                make_pd: bool = False, verbose: bool = False, dtype: Optional[_dtype] = torch.double) -> Tensor:
    """Squared W2 (Gelbrich) distance between N(mean_source, cov_source) and N(mean_target, cov_target):
    |ms-mt|^2 + tr(Cs + Ct - 2 (Ct^1/2 Cs Ct^1/2)^1/2).  Shapes [*, D], [*, D], [*, D, D], [*, D, D] -> [*]
    (reference w2_utils.py:40-80; the first validation is always done in fp64, reference :67)."""
    mean_source, mean_target, cov_source, cov_target = _validate_args(
        mean_source, "mean_source", "vec", mean_target, "mean_target", "vec",
        cov_source, "cov_source", "spd", cov_target, "cov_target", "spd",
        make_pd=make_pd, verbose=verbose, dtype=torch.double)
    # the reference re-validates Ct^1/2 Cs Ct^1/2 as 'spsd' (:73-76); it is a congruence of an SPD matrix, so the
    # check can only trip on round-off.  The kernel symmetrises that product before taking its root.
    out = K.w2_gaussian(mean_source, mean_target, cov_source, cov_target)
    return out.to(device=cov_source.device)  # fp64, like the reference (its math runs on the fp64-validated args)


def batch_w2_dissimilarity_gaussian_diag(mean_source: Tensor, mean_target: Tensor, var_source: Tensor,
                                         var_target: Tensor, dtype: Optional[_dtype] = torch.double) -> Tensor:
    """D[b,i,j] = W2^2(N(ms_bi, diag vs_bi), N(mt_bj, diag vt_bj)), [*, N, D] x [*, M, D] -> [*, N, M]
    (reference w2_utils.py:86-134)."""
    mean_source, var_source = _validate_args(mean_source, "mean_source", "vec", var_source, "var_source", "var",
                                             dtype=dtype)
    mean_target, var_target = _validate_args(mean_target, "mean_target", "vec", var_target, "var_target", "var",
                                             dtype=dtype)
    # |ms-mt|^2 + |sqrt(vs)-sqrt(vt)|^2 = squared distance between the stacked features [m, sqrt(v)]
    fs = torch.cat([mean_source, var_source.sqrt()], dim=-1)
    ft = torch.cat([mean_target, var_target.sqrt()], dim=-1)
    return (fs * fs).sum(-1, keepdim=True) + (ft * ft).sum(-1).unsqueeze(-2) - 2 * (fs @ ft.transpose(-2, -1))


def batch_w2_dissimilarity_gaussian(mean_source: Tensor, mean_target: Tensor, cov_source: Tensor, cov_target: Tensor,
                                    make_pd: bool = False, verbose: bool = False,
                                    dtype: Optional[_dtype] = torch.double) -> Tensor:
    """Full-covariance dissimilarity matrix [*, N, M] (reference w2_utils.py:140-191): all N*M pairs in one batched
    `w2_gaussian` call."""
    mean_source, cov_source = _validate_args(mean_source, "mean_source", "vec", cov_source, "cov_source", "spd",
                                             dtype=dtype)
    mean_target, cov_target = _validate_args(mean_target, "mean_target", "vec", cov_target, "cov_target", "spd",
                                             dtype=dtype)
    n, m = mean_source.size(-2), mean_target.size(-2)
    lead = mean_source.shape[:-2]
    ms = mean_source.unsqueeze(-2).expand(*lead, n, m, -1)
    mt = mean_target.unsqueeze(-3).expand(*lead, n, m, -1)
    cs = cov_source.unsqueeze(-3).expand(*lead, n, m, -1, -1)
    ct = cov_target.unsqueeze(-4).expand(*lead, n, m, -1, -1)
    return w2_gaussian(ms, mt, cs, ct, make_pd=make_pd, verbose=verbose, dtype=dtype)


def batch_ot_gmm(mean_source: Tensor, mean_target: Tensor, cov_source: Tensor, cov_target: Tensor, diag: bool,
                 weight_source: Optional[Tensor] = None, weight_target: Optional[Tensor] = None, verbose: bool = False,
                 dtype: Optional[_dtype] = torch.double, **sinkhorn_kwargs) -> Tuple[Tensor, Tensor]:
    """Entropic OT between two Gaussian mixtures (reference w2_utils.py:197-270): component-wise W2^2 cost,
    normalised by its max for the Sinkhorn solve, total cost reported with the un-normalised cost."""
    if weight_source is None:
        weight_source = torch.ones_like(mean_source.select(dim=-1, index=0)) / mean_source.size(-2)
    if weight_target is None:
        weight_target = torch.ones_like(mean_target.select(dim=-1, index=0)) / mean_target.size(-2)
    kind = "var" if diag else "spd"
    mean_source, cov_source, weight_source = _validate_args(
        mean_source, "mean_source", "vec", cov_source, "cov_source", kind, weight_source, "weight_source", "prob",
        dtype=dtype)
    mean_target, cov_target, weight_target = _validate_args(
        mean_target, "mean_target", "vec", cov_target, "cov_target", kind, weight_target, "weight_target", "prob",
        dtype=dtype)
    if diag:
        routed = _ot_gmm_diag_on_points(mean_source, mean_target, cov_source, cov_target, weight_source, weight_target,
                                        **sinkhorn_kwargs)
        if routed is not None:
            return routed
        cost = batch_w2_dissimilarity_gaussian_diag(mean_source, mean_target, cov_source, cov_target, dtype=dtype)
    else:
        cost = batch_w2_dissimilarity_gaussian(mean_source, mean_target, cov_source, cov_target, make_pd=True,
                                               verbose=verbose, dtype=dtype)
    peak = cost.amax(dim=(-2, -1), keepdim=True)
    coupling = sinkhorn_log(weight_source, weight_target, cost / peak, **sinkhorn_kwargs)
    return (cost * coupling).sum(dim=(-2, -1)), coupling


def _ot_gmm_diag_on_points(mean_s: Tensor, mean_t: Tensor, var_s: Tensor, var_t: Tensor, w_s: Tensor, w_t: Tensor,
                           reg: float = 1e-5, max_iter: int = 1000, threshold: float = STABILITY_CONST
                           ) -> Optional[Tuple[Tensor, Tensor]]:
    """The diagonal mixture OT without a host-visible cost matrix: W2^2 between diagonal Gaussians is the squared distance
    between the stacked features [m, sqrt(v)] (reference w2_utils.py:121-125), so the component problem is a point-cloud
    Sinkhorn (`otk_sinkhorn_points`, squared-Euclidean tiles, cost normalised by its max as at :265-266).  Used for
    un-batched CUDA inputs whose 1 / reg stays within fp32 (the normalised cost is <= 1); None otherwise."""
    if mean_s.dim() != 2 or not mean_s.is_cuda or 1.0 / reg > 1.0e3:
        return None
    fs = torch.cat([mean_s, var_s.sqrt()], dim=-1).float().contiguous()
    ft = torch.cat([mean_t, var_t.sqrt()], dim=-1).float().contiguous()
    peak = float(K.cost_max(fs, ft, 0))
    if not peak > 0:
        return None
    res = K.sinkhorn_points(fs, ft, w_s, w_t, reg=reg, max_iter=max_iter, threshold=threshold, scale=1.0 / peak, precision=1,
                            want_iters=False)
    coupling = K.points_plan(fs, ft, res["u"], res["v"], 1.0 / peak, reg).to(mean_s.dtype)
    return (res["summary"][0] * peak).to(mean_s.dtype), coupling


def sinkhorn_log(a: Tensor, b: Tensor, C: Tensor, reg: float = 1e-5, max_iter: int = 1000,
                 threshold: float = STABILITY_CONST) -> Tensor:
    """Log-domain Sinkhorn plan [*, N, M] for marginals a [*, N], b [*, M] and cost C [*, N, M]
    (reference w2_utils.py:276-319, same recurrences and stop rule; see include/otk.h otk_sinkhorn_dense).
    fp64 inputs are solved in fp64, anything else in fp32."""
    plan, _, _, _ = K.sinkhorn_dense(a, b, C, reg, max_iter, threshold, want_plan=True)
    return plan.to(device=C.device, dtype=C.dtype if C.is_floating_point() else plan.dtype)


def gaussian_barycenter(mean: Tensor, cov: Tensor, weights: Tensor, diag: bool, n_iter: int = 100,
                        dtype: Optional[_dtype] = torch.double) -> Tuple[Tensor, Tensor]:
    """W2 barycenter of N(mean_i, cov_i) with weights w_i (reference w2_utils.py:325-385): closed form when `diag`,
    the Alvarez-Esteban fixed point S <- sum_i w_i (S^1/2 C_i S^1/2)^1/2 otherwise."""
    mean, cov, weights = _validate_args(mean, "mean", "vec", cov, "cov", "var" if diag else "spd",
                                        weights, "weights", "prob", dtype=dtype)
    mean_b = (weights.unsqueeze(-2) @ mean).squeeze(-2)
    if diag:
        return mean_b, ((weights.unsqueeze(-2) @ cov.sqrt()) ** 2).squeeze(-2)
    w = weights[..., None, None]
    start = int(torch.randint(size=(1,), high=cov.size(-3)).item())
    cov_b = cov.select(dim=-3, index=start).unsqueeze(-3)
    for _ in range(n_iter):
        root = sqrtm(cov_b)
        cov_b = (w * sqrtm(root @ cov @ root)).sum(-3, keepdim=True)
    return mean_b, cov_b.squeeze(-3)


# ---------------------------------------------------------------------------------------------------------------------

def compute_transport_operators(cov_source: Tensor, cov_target: Tensor, stochastic: bool, diag: bool,
                                pg_star: float = 0, make_pd: bool = False, verbose: bool = False,
                                dtype: Optional[_dtype] = torch.double) -> Tuple[Tensor, Tensor]:
    """Transport operator T (and noise covariance Cw) of Freirich et al. eq. 17 / 19 (reference w2_utils.py:391-458).
    [*, D, D] (or [*, D] when `diag`) -> (T, Cw).  The full-matrix deterministic branch is the libotk kernel."""
    if stochastic and diag:
        cov_source[cov_source < STABILITY_CONST] = 0
    cov_source, cov_target = _validate_args(
        cov_source, "cov_source", "var" if diag else ("spsd" if stochastic else "spd"),
        cov_target, "cov_target", "var" if diag else ("spd" if stochastic else "spsd"),
        make_pd=make_pd, dtype=dtype, verbose=verbose)
    fallback_msg = ("The noise covariance matrix is not positive definite. "
                    "Falling back to the non-stochastic implementation")
    T = Cw = None
    if diag and stochastic:
        T, Cw = _compute_transport_diag_stochastic(cov_source, cov_target, pg_star)
        if bool((Cw <= 0).any()) and verbose:
            warnings.warn(fallback_msg)
            stochastic = False
    if diag and not stochastic:
        T, Cw = _compute_transport_diag(cov_source, cov_target, pg_star)
    if not diag and stochastic:
        T, Cw = _compute_transport_full_mat_stochastic(cov_source, cov_target, pg_star)
        if not bool(is_spd(Cw, strict=True).all()) and verbose:
            warnings.warn(fallback_msg)
            stochastic = False
    if not diag and not stochastic:
        T, Cw = _compute_transport_full_mat(cov_source, cov_target, pg_star)
    return T, Cw


def apply_transport(input: Tensor, mean_source: Tensor, mean_target: Tensor, T: Tensor, Cw: Optional[Tensor] = None,
                    diag: bool = False, make_pd: bool = False, verbose: bool = False,
                    dtype: Optional[_dtype] = torch.double) -> Tensor:
    """T (input - mean_source) + mean_target (+ W, W ~ N(0, Cw) when Cw is not all-zero)
    (reference w2_utils.py:464-527).  As in the reference, `Cw=None` is rejected by the validation (:619-622)."""
    ignore_cw = Cw is None or bool(torch.allclose(Cw, torch.zeros_like(Cw)))
    input, mean_source, mean_target, T, Cw = _validate_args(
        input, "input", "vec", mean_source, "mean_source", "vec", mean_target, "mean_target", "vec",
        T, "T", "vec" if diag else "mat",
        Cw, "Cw", ("vec" if diag else "mat") if ignore_cw else ("var" if diag else "spd"),
        make_pd=make_pd, verbose=verbose, dtype=dtype)
    if diag:
        moved = T * (input - mean_source) + mean_target
    else:
        moved = _apply_full(input, mean_source, mean_target, T)
    if not ignore_cw:
        # W ~ N(0, Cw) (reference :522-525) without leaving the device kernels: standard normal draws (torch's generator is
        # plumbing) shaped by Cw^1/2 - the Newton-Schulz root, applied with the same streaming GEMM as the map itself -
        # instead of `MultivariateNormal`'s Cholesky factor; the two factors give the same distribution, not the same draws.
        # The diagonal branch keeps the reference's quirk of passing Cw as the *scale* of `D.Normal`.
        eps = torch.randn(moved.shape, dtype=moved.dtype, device=moved.device)
        if diag:
            moved = moved + eps * Cw
        else:
            zero = torch.zeros_like(mean_target)
            moved = moved + _apply_full(eps, zero, zero, sqrtm(Cw))
    return moved


def _apply_full(x: Tensor, mean_s: Tensor, mean_t: Tensor, T: Tensor) -> Tensor:
    """Route the broadcast mat-vec `T @ (x - ms) + mt` (x [*, D], T [*, D, D]) through the GEMM kernel.
    The caller's layout (W2Mixin.apply_transport) is x [*L, B, D], means [*L, 1, D], T [*L, 1, D, D]."""
    out_shape = torch.broadcast_shapes(x.shape, mean_s.shape, mean_t.shape, T.shape[:-1])
    if T.dim() >= 3 and T.size(-3) == 1 and mean_s.dim() >= 2 and mean_s.size(-2) == 1 and mean_t.size(-2) == 1 \
            and x.dim() >= 2:
        y = K.apply_transport(x, mean_s.squeeze(-2), mean_t.squeeze(-2), T.squeeze(-3))
    else:
        # one latent per operator: treat each as a batch of one row
        xb = x.expand(out_shape).unsqueeze(-2)
        ms = mean_s.expand(out_shape)
        mt = mean_t.expand(out_shape)
        Tb = T.expand(*out_shape[:-1], *T.shape[-2:])
        y = K.apply_transport(xb, ms, mt, Tb).squeeze(-2)
    return y.to(device=x.device, dtype=x.dtype).expand(out_shape)


# ---------------------------------------------------------------------------------------------------------------------

class W2Mixin(object):
    """Carrier of the W2 configuration (`stochastic, diag, pg_star, make_pd, verbose, dtype`; unknown keys are
    ignored) and the bound helpers, as in reference w2_utils.py:533-600."""

    def __init__(self, **kwargs):
        self._orig_kwargs = kwargs
        self.stochastic = kwargs.pop("stochastic", False)
        self.diag = kwargs.pop("diag", False)
        self.pg_star = kwargs.pop("pg_star", 0.)
        self.make_pd = kwargs.pop("make_pd", False)
        self.verbose = kwargs.pop("verbose", False)
        self.dtype = kwargs.pop("dtype", torch.double)

        self.mean_cov = partial(mean_cov, diag=self.diag)
        self.batch_w2_dissimilarity_gaussian_diag = partial(batch_w2_dissimilarity_gaussian_diag, dtype=self.dtype)
        self.batch_w2_dissimilarity_gaussian = partial(batch_w2_dissimilarity_gaussian, make_pd=self.make_pd,
                                                       verbose=self.verbose, dtype=self.dtype)
        self.batch_ot_gmm = partial(batch_ot_gmm, diag=self.diag, verbose=self.verbose, dtype=self.dtype)
        self.gaussian_barycenter = partial(gaussian_barycenter, diag=self.diag, dtype=self.dtype)
        self.compute_transport_operators = partial(
            compute_transport_operators, diag=self.diag, stochastic=self.stochastic, pg_star=self.pg_star,
            make_pd=self.make_pd, verbose=self.verbose, dtype=self.dtype)

    def get_var_normal(self, distribution: Union[D.Normal, D.MultivariateNormal]):
        return distribution.variance if self.diag else distribution.covariance_matrix

    def instantiate_normal(self, *args, **kwargs):
        if self.diag:
            for k in ("covariance_matrix", "precision_matrix", "scale_tril"):
                kwargs.pop(k, None)
            return D.Independent(D.Normal(*args, **kwargs), 1)
        kwargs.pop("scale", None)
        return D.MultivariateNormal(*args, **kwargs)

    def w2_gaussian(self, mean_source: Tensor, mean_target: Tensor, cov_source: Tensor, cov_target: Tensor) -> Tensor:
        return w2_gaussian(mean_source, mean_target,
                           torch.diag_embed(cov_source) if self.diag else cov_source,
                           torch.diag_embed(cov_target) if self.diag else cov_target,
                           make_pd=self.make_pd, verbose=self.verbose, dtype=self.dtype)

    def apply_transport(self, inputs: Tensor, mean_source: Tensor, mean_target: Tensor, T: Tensor, Cw: Tensor,
                        batch_dim: Optional[int] = None) -> Tensor:
        if batch_dim is not None:
            mat_dim = batch_dim - int(not self.diag)
            mean_source, mean_target = mean_source.unsqueeze(batch_dim), mean_target.unsqueeze(batch_dim)
            T, Cw = T.unsqueeze(mat_dim), Cw.unsqueeze(mat_dim)
        return apply_transport(inputs, mean_source, mean_target, T, Cw, diag=self.diag, make_pd=self.make_pd,
                               verbose=self.verbose, dtype=self.dtype)

    def __repr__(self):
        return ", ".join(f"{k}={v}" for k, v in self._orig_kwargs.items())


# ---------------------------------------------------------------------------------------------------------------------
# operator helpers (no argument checking; reference w2_utils.py:714-793)

def _compute_transport_diag(cov_source: Tensor, cov_target: Tensor, p_gstar) -> Tuple[Tensor, Tensor]:
    T = (1 - p_gstar) * torch.sqrt(cov_target / cov_source + STABILITY_CONST) + p_gstar
    return T, torch.zeros_like(T)


def _compute_transport_diag_stochastic(cov_source: Tensor, cov_target: Tensor, p_gstar: float) -> Tuple[Tensor, Tensor]:
    T_star = torch.sqrt(cov_source / cov_target + STABILITY_CONST)
    live = cov_source > STABILITY_CONST
    pinv_source = torch.where(live, 1 / torch.where(live, cov_source, torch.ones_like(cov_source)),
                              torch.zeros_like(cov_source))
    T = (1 - p_gstar) * torch.sqrt(cov_target * cov_source) * pinv_source + p_gstar
    var_w = sqrt(1 - p_gstar) * cov_target * (1 - cov_target * pinv_source * T_star ** 2)
    return T, var_w


def _compute_transport_full_mat(cov_source: Tensor, cov_target: Tensor, p_gstar: float) -> Tuple[Tensor, Tensor]:
    """T = (1-p) Cs^-1/2 (Cs^1/2 Ct Cs^1/2)^1/2 Cs^-1/2 + p I, Cw = 0 (reference :756-768) in one libotk call."""
    T, _ = K.transport_operator(cov_source, cov_target, pg_star=float(p_gstar))
    T = T.to(device=cov_source.device, dtype=cov_source.dtype)
    return T, torch.zeros_like(T)


def _compute_transport_full_mat_stochastic(cov_source: Tensor, cov_target: Tensor, pg_star: float
                                           ) -> Tuple[Tensor, Tensor]:
    """eq. 19 (reference :774-793) in one libotk call: the roots, the pseudo-inverse of the source (its inverse, taken from
    a Newton-Schulz inverse root: the reference's own pipeline is non-finite for a rank-deficient source) and the fourteen
    d x d products all run on the device; nothing goes through `torch.linalg`."""
    T, Cw = K.transport_operator_stochastic(cov_source, cov_target, pg_star=float(pg_star))
    return (T.to(device=cov_source.device, dtype=cov_source.dtype),
            Cw.to(device=cov_source.device, dtype=cov_source.dtype))
