"""Streaming Gaussian-mixture estimator on B200 (mirror of reference
ot/distribution_models/gassian_mixture_model.py:28-189 - the reference's file name, typo included, so that
`ot_vae_lightning.ot.distribution_models.gassian_mixture_model` resolves after `install_as_reference()`).

Same class layout as the reference: `GaussianMixtureModel(GaussianModel, CodebookModel)` with the component axis as one
more leading axis of the Gaussian buffers (`mean [*L, K, d]`, `cov [*L, K, d, d]`, `_n_obs [*L, K]`), the k-means style
`update` / `fit` drivers of `CodebookModel` and the `_weights` parameter behind the `NormSum` parametrization.

Hot step: `kmean_iteration` - the per-component weighted statistics
    n_k = sum_b w_bk ,   s_k = sum_b w_bk x_b ,   P_k = sum_b w_bk x_b x_b^T .
The reference materialises the B x d^2 outer products and multiplies them by the assignment matrix
(gassian_mixture_model.py:109-115).  Here P_k is a SYRK of the rows the component owns, scaled by sqrt(w_bk): the
libotk statistics kernel (otk_stats_update: FP16-split tcgen05 SYRK / fused fp64 path for small groups) runs once per
component on the gathered rows; nothing of size B x d^2 exists.  Log-likelihood energies go through
torch.distributions exactly as in the reference (:86-94).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributions as D
import torch.nn as nn
import torch.nn.utils.parametrize as P
from torch import Tensor

from ... import kernels as K
from ..w2_utils import W2Mixin
from .base import MixtureMixin
from .codebook_model import CodebookModel
from .gaussian_model import GaussianModel

__all__ = ["GaussianMixtureModel", "NormSum"]


class GaussianMixtureModel(GaussianModel, CodebookModel):
    Distribution = D.MixtureSameFamily

    def __init__(self, *size: int, mixture_cfg={}, **kwargs):
        # as in the reference (:48-51): the grand-parent mixin and GaussianModel are initialised, CodebookModel is not -
        # its codebook / buffers are replaced by the Gaussian ones
        MixtureMixin.__init__(self, *size[:-1], **dict(mixture_cfg))
        GaussianModel.__init__(self, *size, **kwargs)
        self.batch_dim = -3
        self.register_buffer("weight_init", self._weight_init.to(self.vec_init))
        self._weights = nn.Parameter(self.weight_init.clone(), requires_grad=self.update_with_autograd)
        if self.update_with_autograd:
            P.register_parametrization(self, "_weights", nn.Softmax(-1))
        else:
            P.register_parametrization(self, "_weights", NormSum(1.))

    @property
    def weights(self):
        return self._weights

    @property
    def vec_shape(self):
        return *self.leading_shape, self.n_components, self.dim

    def reset(self) -> None:
        GaussianModel.reset(self)
        with torch.no_grad():
            self._weights = self.weight_init

    @property
    def distribution(self) -> Distribution:
        return D.MixtureSameFamily(D.Categorical(self.weights), GaussianModel.distribution.fget(self))

    @property
    def batched_distribution(self) -> Distribution:
        return D.MixtureSameFamily(D.Categorical(self.weights.unsqueeze(-2)),
                                   GaussianModel.batched_distribution.fget(self))

    @property
    def variances(self) -> Tensor:
        return self.get_var_normal(self.distribution.component_distribution)

    @property
    def batched_variances(self) -> Tensor:
        return self.get_var_normal(self.batched_distribution.component_distribution)

    def update(self, samples: Tensor) -> None:
        return CodebookModel.update(self, samples)

    @torch.no_grad()
    def fit(self, samples: Optional[Tensor] = None) -> None:
        """reference :83-84 -> codebook_model.py:136-149.  Without samples every k-means iteration maps the same running
        buffers to the same parameters, so one pass replaces the reference's `kmeans_iter` identical ones (it differs by
        the 1e-8 ridge the PSD parametrization re-adds per read of a never-seen component: < 1e-6 absolute)."""
        if samples is not None:
            return CodebookModel.fit(self, samples)
        self._fit_warn()
        if self.kmeans_iter > 0:
            res = self.kmean_iteration(None)
            self._update_parameters(*[self.reduce(r) for r in res])
            self._update_buffers(*res, decay=False)

    def energy(self, samples: Tensor) -> Tensor:
        """log N(x | mean_k, cov_k) + log w_k, [*L, B, K] (reference :86-94)."""
        self._validate_samples(samples)
        dist = self.batched_distribution
        samples = samples.to(dist.mean)
        padded = dist._pad(samples)                                        # [*L, B, 1, d]
        log_prob_x = dist.component_distribution.log_prob(padded)           # [*L, B, K]
        log_mix = torch.log_softmax(dist.mixture_distribution.logits, dim=-1)
        return (log_prob_x + log_mix).type_as(samples)

    def predict_mean_var(self, assignments: Tensor) -> Tuple[Tensor, Tensor]:
        """[*L, B, K] assignments -> per-sample (mean [*L, B, d], var [*L, B, d(, d)]) (reference :96-102)."""
        mean = assignments.type_as(self.mean) @ self.mean
        variances = self.variances
        var = assignments.type_as(variances) @ (variances if self.diag else variances.flatten(-2))
        if not self.diag:
            var = var.unflatten(-1, (self.dim, self.dim))
        return mean.type_as(assignments), var.type_as(assignments)

    def kmean_iteration(self, samples: Optional[Tensor]) -> Tuple[Tensor, Tensor, Tensor]:
        """(sum_b w_bk, sum_b w_bk x_b, sum_b w_bk x_b x_b^T) per component (reference :104-117)."""
        if samples is None:
            return self._n_obs, self._running_sum, self._running_sum_cov
        weights, _, _ = self.assign(samples)                                # [*L, B, K]
        weights = weights.to(samples)
        weights_sum = weights.sum(-2)
        weighted_sum = weights.transpose(-1, -2) @ samples                  # [*L, K, d]
        if self.diag:
            return weights_sum, weighted_sum, weights.transpose(-1, -2) @ (samples ** 2)
        return weights_sum, weighted_sum, self._weighted_syrk(samples, weights)

    def _weighted_syrk(self, samples: Tensor, weights: Tensor) -> Tensor:
        """P[*L, k] = sum_b w_bk x_b x_b^T through the statistics kernel (a SYRK of the rows scaled by sqrt(w))."""
        self._require_cuda_buffers()
        lead, k, d = weights.shape[:-2], weights.size(-1), samples.size(-1)
        b = samples.size(-2)
        if k * b * d * max(1, int(torch.Size(lead).numel())) * 4 <= (1 << 30):
            # ONE batched launch: component k of leading index l is "leading index (l, k)" of the statistics kernel, fed with
            # the batch scaled by sqrt(w_bk) (rows a component does not own are zero) - no per-component Python loop, no
            # host read-back of the assignment pattern
            scaled = (weights.transpose(-1, -2).sqrt().unsqueeze(-1) * samples.expand(*lead, b, d).unsqueeze(-3)).float()
            out = torch.zeros(*lead, k, d, d, dtype=self._running_sum_cov.dtype, device=self._running_sum_cov.device)
            n = torch.zeros(*lead, k, dtype=torch.float64, device=out.device)
            s = torch.zeros(*lead, k, d, dtype=out.dtype, device=out.device)
            K.stats_update(scaled.contiguous(), n, s, out, None)
            return out
        # very large batches: gather the rows of each component instead of materialising the K scaled copies
        x2 = samples.expand(*lead, *samples.shape[-2:]).reshape(-1, samples.size(-2), d)
        w2 = weights.reshape(-1, weights.size(-2), k)
        flat = out.view(-1, k, d, d)
        hard = bool(((w2 == 0) | (w2 == 1)).all())
        for li in range(x2.size(0)):
            if hard:
                owner = w2[li].argmax(-1)
                order = torch.argsort(owner, stable=True)
                counts = torch.bincount(owner, minlength=k).tolist()
                rows = x2[li].index_select(0, order).float()
                lo = 0
                for ki, c in enumerate(counts):
                    if c:
                        self._syrk_into(rows[lo:lo + c], flat[li, ki])
                    lo += c
            else:
                for ki in range(k):
                    sel = torch.nonzero(w2[li, :, ki] > 0).squeeze(-1)
                    if sel.numel():
                        rows = (x2[li].index_select(0, sel) * w2[li, sel, ki].sqrt().unsqueeze(-1)).float()
                        self._syrk_into(rows, flat[li, ki])
        return out

    @staticmethod
    def _syrk_into(rows: Tensor, target: Tensor) -> None:
        n = torch.zeros((), dtype=torch.float64, device=target.device)
        s = torch.zeros(rows.size(-1), dtype=target.dtype, device=target.device)
        K.stats_update(rows.contiguous(), n, s, target, None)

    def w2(self, other: Distribution) -> Tensor:
        total_cost, _ = self.batch_ot_gmm(
            self.mean, other.component_distribution.mean, self.variances,
            self.get_var_normal(other.component_distribution), weight_source=self.weights,
            weight_target=other.mixture_distribution.probs, max_iter=100)
        return total_cost

    def extra_repr(self) -> str:
        return GaussianModel.extra_repr(self) + W2Mixin.__repr__(self) + MixtureMixin.extra_repr(self)

    # -------------------------------------------------------------------------------------------------------------
    def _update_weights(self, val: Optional[Tensor], seen: Optional[Tensor] = None):
        """reference :141-150: unseen components keep their (normalised) current weight, seen ones take the raw count;
        `NormSum` renormalises on the next read."""
        if val is None:
            return
        with torch.no_grad():
            if seen is None:
                self._weights = val.type_as(self._weights)
            else:
                current = self._weights
                self._weights = torch.where(seen, val.type_as(current), current)

    def _update_parameters(self, *kmeans_iter_res):
        weights_sum, samples_sum, samples_cov_sum = kmeans_iter_res
        mean, cov, seen = self._compute_mean_cov(self.laplace_smoothing(weights_sum), samples_sum, samples_cov_sum)
        self._update_mean(mean, seen)
        self._update_cov(cov, seen)
        if seen is not None:
            self._update_weights(weights_sum, seen)

    def _update_buffers(self, *kmeans_iter_res, decay=False):
        weights_sum, samples_sum, samples_cov_sum = kmeans_iter_res
        hit = weights_sum > 1e-8
        hv = hit.unsqueeze(-1)
        hm = hv if self.diag else hit[..., None, None]
        if decay:
            self._n_obs = torch.where(hit, self.ema_update(self._n_obs, weights_sum.to(self._n_obs)), self._n_obs)
            self._running_sum = torch.where(hv, self.ema_update(self._running_sum, samples_sum), self._running_sum)
            self._running_sum_cov = torch.where(hm, self.ema_update(self._running_sum_cov, samples_cov_sum),
                                                self._running_sum_cov)
        else:
            self._n_obs = torch.where(hit, weights_sum.to(self._n_obs), self._n_obs)
            self._running_sum = torch.where(hv, samples_sum.to(self._running_sum), self._running_sum)
            self._running_sum_cov = torch.where(hm, samples_cov_sum.to(self._running_sum_cov), self._running_sum_cov)
        return self._n_obs, self._running_sum, self._running_sum_cov

    def _init_parameters(self, samples: Tensor) -> None:
        """reference :173-177: the means start from randomly picked samples (host-side `randperm`, as the reference
        draws it, so a seeded run picks the same rows)."""
        if bool(torch.allclose(self.mean, self.vec_init)):
            pick = torch.randperm(samples.size(-2))[:self.n_components].to(samples.device)
            self._update_mean(samples[..., pick, :])
            self._n_obs += 1


class NormSum(nn.Module):
    """weights / sum(weights) on every read (reference :180-189)."""

    def __init__(self, val=1.):
        super().__init__()
        self.val = val

    def forward(self, X):
        return self.val * X / X.sum(-1, keepdim=True)

    def right_inverse(self, X):
        return X
