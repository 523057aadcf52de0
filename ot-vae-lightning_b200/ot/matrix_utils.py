"""SPD matrix primitives of the latent-OT path on B200 (same names/signatures as the reference's
`ot_vae_lightning/ot/matrix_utils.py:20-31`).

The reference evaluates every matrix function through an fp64 `torch.linalg.eigh`; here
`sqrtm`/`invsqrtm` are coupled Newton-Schulz iterations on the tensor cores (libotk `otk_sqrtm`) and
`min_eig` is a Lanczos + Sturm-bisection kernel (`otk_min_eig`).  Results come back in the input dtype.
"""
from __future__ import annotations

from typing import Tuple, Union

import torch
from torch import BoolTensor, Tensor

from .. import kernels as K
from ..utils import unsqueeze_like

__all__ = ["eye_like", "sqrtm", "invsqrtm", "is_spd", "is_pd", "is_symmetric", "min_eig", "make_psd", "mean_cov",
           "STABILITY_CONST"]

STABILITY_CONST = 1e-8


def _back(result: Tensor, like: Tensor) -> Tensor:
    """Return `result` on the device (and float dtype) of the caller's tensor."""
    dt = like.dtype if like.is_floating_point() else result.dtype
    return result.to(device=like.device, dtype=dt)


def eye_like(matrices: Tensor) -> Tensor:
    """Identity broadcast to the shape of `matrices` [*, C, D] (reference matrix_utils.py:49-56)."""
    rows, cols = matrices.shape[-2:]
    return torch.eye(rows, cols, dtype=matrices.dtype, device=matrices.device).expand_as(matrices)


def sqrtm(matrices: Tensor) -> Tensor:
    """Principal square root of a batch of SPSD matrices (reference matrix_utils.py:59-65)."""
    root, _ = K.sqrtm_pair(matrices, want_root=True, want_iroot=False)
    return _back(root, matrices)


def invsqrtm(matrices: Tensor) -> Tensor:
    """Inverse square root of a batch of SPD matrices (reference matrix_utils.py:68-76)."""
    _, iroot = K.sqrtm_pair(matrices, want_root=False, want_iroot=True)
    return _back(iroot, matrices)


def is_symmetric(matrices: Tensor) -> BoolTensor:
    """sum((A - A^T)^2) < 1e-8 per matrix; non-square -> all False (reference matrix_utils.py:79-88)."""
    if matrices.size(-1) != matrices.size(-2):
        return torch.zeros(matrices.shape[:-2], dtype=torch.bool, device=matrices.device)
    return (K.asymmetry(matrices) < STABILITY_CONST).to(matrices.device)


def min_eig(matrices: Tensor) -> Tensor:
    """Smallest (signed) eigenvalue per matrix, lower triangle read (reference matrix_utils.py:91-98)."""
    return _back(K.min_eig(matrices), matrices)


def is_pd(matrices: Tensor, strict=True) -> BoolTensor:
    """reference matrix_utils.py:101-109"""
    lam = min_eig(matrices)
    return lam > 0 if strict else lam >= 0


def is_spd(matrices: Tensor, strict=True) -> BoolTensor:
    """reference matrix_utils.py:112-120"""
    return torch.logical_and(is_symmetric(matrices), is_pd(matrices, strict=strict)).bool()


def make_psd(matrices: Tensor, strict: bool = False, return_correction: bool = False, diag: bool = False
             ) -> Union[Tensor, Tuple[Tensor, Tensor]]:
    """Add max(0, -lambda_min) (+1e-8 if strict) to the diagonal (reference matrix_utils.py:123-142)."""
    lowest = matrices.min(-1)[0] if diag else min_eig(matrices)
    shift = lowest.clamp(max=0).abs()
    if strict:
        shift = shift + STABILITY_CONST
    if diag:
        fixed = matrices + shift[..., None]
    else:
        fixed = matrices + eye_like(matrices) * shift[..., None, None]
    return (fixed, shift) if return_correction else fixed


def mean_cov(sum: Tensor, sum_corr: Tensor, num_obs: Union[Tensor, int], diag: bool = False) -> Tuple[Tensor, Tensor]:
    """mean = sum/n ; cov = sum_corr/n - mean mean^T (biased; reference matrix_utils.py:145-158).
    `num_obs` must be a tensor, as in the reference (`unsqueeze_like` has no int path)."""
    if diag:
        mean = sum / unsqueeze_like(num_obs, sum)
        return mean, sum_corr / unsqueeze_like(num_obs, sum_corr) - mean ** 2
    unsqueeze_like(num_obs, sum)  # same failure mode as the reference for python numbers / too many dims
    mean, cov = K.mean_cov(sum, sum_corr, num_obs)
    out_dt = torch.promote_types(sum.dtype, num_obs.dtype) if num_obs.is_floating_point() else sum.dtype
    return mean.to(device=sum.device, dtype=out_dt), cov.to(device=sum_corr.device, dtype=out_dt)
