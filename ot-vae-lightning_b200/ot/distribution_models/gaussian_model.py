"""Streaming multivariate-Gaussian estimator on B200 (mirror of reference
ot/distribution_models/gaussian_model.py:29-229: same constructor, buffers, parameters and state_dict keys).

`update` streams the latents through the libotk statistics kernel (fused sum x / sum x x^T / count / EMA), `fit`
combines ranks with ONE packed all-reduce and finalises mean / covariance on the device.  Reading `.cov` goes through
the same two parametrizations as the reference (triu mirror, then + (max(0,-lambda_min)+1e-8) I); wrap repeated reads
in `torch.nn.utils.parametrize.cached()` to evaluate them once.
"""
from __future__ import annotations

from typing import Optional, Tuple, Union

import torch
import torch.distributions as D
import torch.nn as nn
import torch.nn.utils.parametrize as P
from torch import Tensor

from ... import kernels as K
from ...utils import _dist_on, ddp_reduce_func_default
from ..matrix_utils import eye_like, make_psd
from ..w2_utils import W2Mixin
from .base import DistributionModel

__all__ = ["GaussianModel"]


class GaussianModel(DistributionModel, W2Mixin):
    Distribution = Union[D.Independent, D.MultivariateNormal]

    def __init__(self, *size: int, w2_cfg={}, **kwargs):
        DistributionModel.__init__(self, *size, **kwargs)
        W2Mixin.__init__(self, **dict(w2_cfg))
        self.batch_dim = -2
        self.register_buffer("cov_init", torch.ones_like(self.vec_init) if self.diag else eye_like(self.mat_init).clone())
        self.mean = nn.Parameter(self.vec_init.clone(), requires_grad=self.update_with_autograd)
        self.cov = nn.Parameter(self.cov_init.clone(), requires_grad=self.update_with_autograd)
        if self.update_with_autograd:
            P.register_parametrization(self, "cov", ExpScaleTril(diag=self.diag), unsafe=True)
        else:
            self.register_buffer("_running_sum", torch.zeros_like(self.mean.data))
            self.register_buffer("_running_sum_cov", torch.zeros_like(self.cov.data))
            self.register_buffer("_n_obs", torch.zeros(self.vec_shape[:-1], dtype=self.dtype,
                                                       device=self.vec_init.device))
            P.register_parametrization(self, "cov", Symmetric(diag=self.diag), unsafe=True)
            P.register_parametrization(self, "cov", MakePositiveDefinite(diag=self.diag, strict=True), unsafe=True)

    # ------------------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def reset(self) -> None:
        self._update_mean(self.vec_init)
        self._update_cov(self.cov_init)
        if not self.update_with_autograd:
            self._running_sum.zero_()
            self._running_sum_cov.zero_()
            self._n_obs.zero_()

    @property
    def distribution(self) -> Distribution:
        cov = self.cov
        return self.instantiate_normal(
            self.mean, scale=cov ** 0.5,
            scale_tril=cov if self.update_with_autograd else None,
            covariance_matrix=cov if not self.update_with_autograd else None)

    @property
    def batched_distribution(self) -> Distribution:
        cov = self.cov
        full = not self.diag
        return self.instantiate_normal(
            self.mean.unsqueeze(self.batch_dim), scale=cov.unsqueeze(self.batch_dim) ** 0.5,
            scale_tril=cov.unsqueeze(self.batch_dim - 1) if self.update_with_autograd and full else None,
            covariance_matrix=cov.unsqueeze(self.batch_dim - 1) if not self.update_with_autograd and full else None)

    @property
    def variances(self) -> Tensor:
        return self.get_var_normal(self.distribution)

    # ------------------------------------------------------------------------------------------------------------
    def _reduce_is_active(self) -> bool:
        """True if `self.reduce` can change values (a process group exists, or the user injected a function)."""
        return _dist_on() if self.reduce is ddp_reduce_func_default else True

    @torch.no_grad()
    def update(self, samples: Tensor) -> None:
        """reference gaussian_model.py:99-108"""
        if samples.dim() == 2 and self._update_fast(samples):
            return
        self._update_warn()
        self._validate_samples(samples)
        if self.diag:
            x = samples.detach().to(self._running_sum)
            n, s, ss = self._stats(x, reduce=self.reduce_on_update)
            self._n_obs = self.ema_update(self._n_obs, n.to(self._n_obs))
            self._running_sum = self.ema_update(self._running_sum, s)
            self._running_sum_cov = self.ema_update(self._running_sum_cov, ss)
            return
        if self.reduce_on_update and self._reduce_is_active():
            # exact reference semantics: the BATCH statistics are summed over ranks before the EMA
            n, s, ss = self._batch_stats(samples)
            n, s, ss = self._packed_reduce(n, s, ss)
            self._n_obs = self.ema_update(self._n_obs, n.to(self._n_obs))
            self._running_sum = self.ema_update(self._running_sum, s)
            self._running_sum_cov = self.ema_update(self._running_sum_cov, ss)
        else:
            self._require_cuda_buffers()
            K.stats_update(samples, self._n_obs, self._running_sum, self._running_sum_cov, self.decay)

    def _update_fast(self, samples: Tensor) -> bool:
        """The per-batch call of the latency mode (cfg1: 40 batches of 250 per model and epoch) with the constant part of
        the native call cached: a contiguous fp32 [B, d] batch on the model's device, one model (no leading shape), no
        per-update reduction.  Returns False when the batch does not qualify (the general `update` takes it)."""
        plan = self._fast_plan()
        return plan is not None and plan(samples)

    def _fast_plan(self):
        """the cached `StatsUpdatePlan` of `_update_fast` (built on demand), or None when the model does not qualify"""
        plan = self.__dict__.get("_update_plan")
        rs = self._buffers.get("_running_sum")
        if rs is None:
            return None
        if plan is None or plan[0] is not rs or plan[1] is not self._buffers["_running_sum_cov"] or plan[2] is not self._buffers["_n_obs"]:
            if self.diag or self.update_with_autograd or len(self.leading_shape) or not rs.is_cuda \
                    or (self.reduce_on_update and self._reduce_is_active()):
                return None
            rc, no = self._buffers["_running_sum_cov"], self._buffers["_n_obs"]
            if not (rs.is_contiguous() and rc.is_contiguous() and no.is_contiguous()):
                return None
            plan = (rs, rc, no, K.StatsUpdatePlan(no, rs, rc, self.decay))
            self.__dict__["_update_plan"] = plan
        return plan[3]

    @torch.no_grad()
    def fit(self, samples: Optional[Tensor] = None, cov_operand: Optional[Tensor] = None, operand_shift: float = 0.0,
            already_reduced: bool = False) -> None:
        """reference gaussian_model.py:110-126.  `cov_operand` (an extension used by `GaussianTransport.compute`): a buffer
        [*L, d, d] of the parameter dtype that additionally receives triu-mirror(raw covariance) + `operand_shift` I;
        `already_reduced`: the caller has summed the running statistics over the ranks (one joint all-reduce for both
        models of a `GaussianTransport`)."""
        self._fit_warn()
        if self.update_with_autograd:
            if samples is None:
                return
            mean, cov, seen = self._compute_mean_cov(*self._stats(samples, reduce=True))
            self._update_mean(mean, seen)
            self._update_cov(cov, seen)
        if samples is not None:
            self.update(samples)
        if self._native_fit(cov_operand, operand_shift, already_reduced):
            return
        self._n_obs, self._running_sum, self._running_sum_cov = self._stats(None, reduce=not already_reduced)
        mean, cov, seen = self._compute_mean_cov(self._n_obs, self._running_sum, self._running_sum_cov)
        self._update_mean(mean, seen)
        self._update_cov(cov, seen)
        self._fit_generation = getattr(self, "_fit_generation", 0) + 1
        if cov_operand is not None:
            shift = torch.full(self.vec_shape[:-1], operand_shift, dtype=cov_operand.dtype, device=cov_operand.device)
            K.symmetrize_shift(self.parametrizations.cov.original.to(cov_operand.dtype), shift, out=cov_operand)

    def _native_fit(self, cov_operand: Optional[Tensor], operand_shift: float, already_reduced: bool = False) -> bool:
        """The whole fit as ONE kernel (`otk_gaussian_fit`): mean = sum / n and the raw covariance written in place for every
        leading index that has observations (the others keep their state), no host read-back.  Full-covariance models on
        a CUDA device; under a process group the packed all-reduce of the statistics runs first."""
        if self.diag or self.update_with_autograd or not self._running_sum.is_cuda:
            return False
        mean, raw = self.mean, self.parametrizations.cov.original
        ok = (mean.dtype == raw.dtype and mean.dtype in (torch.float32, torch.float64) and mean.is_contiguous()
              and raw.is_contiguous() and mean.device == self._running_sum.device
              and (cov_operand is None or (cov_operand.dtype == raw.dtype and cov_operand.shape == raw.shape
                                           and cov_operand.is_contiguous() and cov_operand.device == raw.device)))
        if not ok:
            return False
        if self._reduce_is_active() and not already_reduced:
            self._n_obs, self._running_sum, self._running_sum_cov = self._stats(None, reduce=True)
        for buf in (self._n_obs, self._running_sum, self._running_sum_cov):
            if not buf.is_contiguous():
                return False
        K.gaussian_fit(self._n_obs, self._running_sum, self._running_sum_cov, mean.data, raw.data, cov_operand, operand_shift)
        self._fit_generation = getattr(self, "_fit_generation", 0) + 1     # caches keyed on the fitted state look at this
        return True

    def predict(self, samples: Tensor) -> Tensor:
        self._validate_samples(samples)
        return self.batched_distribution.log_prob(samples.to(self.mean))

    def w2(self, other: Distribution) -> Tensor:
        return self.w2_gaussian(self.mean, other.mean, self.variances, self.get_var_normal(other))

    def extra_repr(self) -> str:
        return super().extra_repr() + W2Mixin.__repr__(self)

    # ------------------------------------------------------------------------------------------------------------
    def _require_cuda_buffers(self):
        if not self._running_sum.is_cuda:
            raise RuntimeError("GaussianModel buffers are on the CPU: move the model to a CUDA device "
                               "(`model.to('cuda')`); the statistics kernels have no CPU path.")

    def _batch_stats(self, samples: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
        """(n, sum x, sum x x^T) of one batch in fresh buffers of the running dtype (kernel call)."""
        self._require_cuda_buffers()
        n = torch.zeros_like(self._n_obs)
        s = torch.zeros_like(self._running_sum)
        ss = torch.zeros_like(self._running_sum_cov)
        K.stats_update(samples, n, s, ss, None)
        return n, s, ss

    def _packed_reduce(self, n: Tensor, s: Tensor, ss: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
        """One reduction of [n | sum | sum_cov] instead of the reference's three (gaussian_model.py:153-156)."""
        flat = torch.cat([n.reshape(-1).to(ss.dtype), s.reshape(-1), ss.reshape(-1)])
        flat = self.reduce(flat)
        a, b = n.numel(), n.numel() + s.numel()
        return flat[:a].reshape(n.shape).to(n.dtype), flat[a:b].reshape(s.shape), flat[b:].reshape(ss.shape)

    def _stats(self, samples: Optional[Tensor], reduce=True):
        """reference gaussian_model.py:144-157"""
        if samples is not None:
            if self.diag:
                n = torch.as_tensor(samples.size(-2), dtype=samples.dtype, device=samples.device)
                s, ss = samples.sum(-2), (samples ** 2).sum(-2)
            else:
                n, s, ss = self._batch_stats(samples)
        else:
            n, s, ss = self._n_obs, self._running_sum, self._running_sum_cov
        if reduce and self._reduce_is_active():
            if self.diag or n.dim() != s.dim() - 1:
                n, s, ss = self.reduce(n), self.reduce(s), self.reduce(ss)
            else:
                n, s, ss = self._packed_reduce(n, s, ss)
        return n, s, ss

    def _compute_mean_cov(self, n_obs: Tensor, sum: Tensor, sum_cov: Tensor
                          ) -> Tuple[Optional[Tensor], Optional[Tensor], Optional[Tensor]]:
        """reference gaussian_model.py:159-165; computed for every leading index, `seen` masks the assignment."""
        if bool((n_obs == 0).all()):
            return None, None, None
        seen = n_obs > 1e-8
        safe_n = torch.where(seen, n_obs, torch.ones_like(n_obs))
        mean, cov = self.mean_cov(sum, sum_cov, safe_n)
        return mean, cov, seen

    def _update_mean(self, val: Optional[Tensor], seen: Optional[Tensor] = None):
        if val is None:
            return
        val = val.to(self.mean)
        if seen is None:
            self.mean.copy_(val)
        else:
            # written through the Parameter itself (fit runs under no_grad), so that `mean._version` moves and caches keyed
            # on it (GaussianTransport's prepared operator) notice a refit of a single model
            self.mean.copy_(torch.where(seen.unsqueeze(-1), val, self.mean.detach()))

    def _update_cov(self, val: Optional[Tensor], seen: Optional[Tensor] = None):
        if val is None:
            return
        if seen is None or bool(seen.all()):
            # every leading index is overwritten: no need to read (and re-parametrize) the current value
            self.cov = val.to(self.parametrizations.cov.original)
        else:
            current = self.cov  # parametrized read, as in the reference (:181)
            mask = seen[..., None] if self.diag else seen[..., None, None]
            self.cov = torch.where(mask, val.to(current), current)


class ExpScaleTril(nn.Module):
    """exp/tril parametrization for the autograd mode (reference gaussian_model.py:186-201); stock torch."""

    def __init__(self, diag):
        super().__init__()
        self.diag = diag

    def forward(self, x: Tensor) -> Tensor:
        if self.diag:
            return x.exp()
        return x.tril(-1) + torch.diag_embed(x.diagonal(dim1=-1, dim2=-2).exp())

    def right_inverse(self, x: Tensor) -> Tensor:
        return x if self.diag else x.tril()


class MakePositiveDefinite(nn.Module):
    """x + (max(0,-lambda_min) + 1e-8 if strict) I on every read (reference gaussian_model.py:204-217)."""

    def __init__(self, diag, strict):
        super().__init__()
        self.diag = diag
        self.strict = strict

    def forward(self, x):
        return make_psd(x, strict=self.strict, return_correction=False, diag=self.diag)

    def right_inverse(self, x):
        return x


class Symmetric(nn.Module):
    """Mirror the upper triangle (reference gaussian_model.py:220-229); libotk `otk_symmetrize_shift`."""

    def __init__(self, diag):
        super().__init__()
        self.diag = diag

    def forward(self, X):
        if self.diag:
            return X
        return K.symmetrize_shift(X, None).to(device=X.device, dtype=X.dtype)

    def right_inverse(self, X):
        return X
