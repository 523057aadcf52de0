"""Gaussian-mixture transport on B200 (mirror of reference ot/transport/gmm_transport.py:28-124, after Chen, Georgiou &
Tannenbaum, "Optimal transport for Gaussian mixture models"): entropic OT between the components (their pairwise
Gaussian W2^2 as cost, the libotk Sinkhorn kernel as solver), then every input is moved by the closed-form Gaussian
map between the source component it is assigned to and its target component.

`transport` differs from the reference in HOW, not in what: the reference builds one d x d operator per INPUT
(`compute_transport_operators` on [B, d, d] covariances: three `eigh` per sample, :115-118).  With hard assignments
('argmax' / 'sample') there are at most K_s x K_t distinct (source component, target component) pairs, so the
operators are computed once per distinct pair (batched Newton-Schulz kernel) and the inputs of a pair go through the
streaming transport kernel together.  'barycenter' targets differ per input and keep the per-input form.
"""
from __future__ import annotations

from typing import Literal

import torch
import torch.nn.functional as F
from torch import Tensor
from torch.distributions import Categorical

from ..distribution_models.gassian_mixture_model import GaussianMixtureModel
from ..w2_utils import W2Mixin
from .base import TransportOperator

__all__ = ["GMMTransport"]


class GMMTransport(TransportOperator, W2Mixin):
    def __init__(self, *size, transport_type: Literal["sample", "argmax", "barycenter"], source_cfg={}, target_cfg={},
                 transport_cfg={}, **kwargs):
        W2Mixin.__init__(self, **dict(transport_cfg))
        TransportOperator.__init__(
            self, *size,
            source_model=GaussianMixtureModel(*size, w2_cfg=dict(transport_cfg), **source_cfg),
            target_model=GaussianMixtureModel(*size, w2_cfg=dict(transport_cfg), **target_cfg),
            **kwargs)
        self.transport_type = transport_type
        self.transport_matrix = None

    def reset(self) -> None:
        super().reset()
        self.transport_matrix = None

    def compute(self) -> Tensor:
        """reference :65-80"""
        self.fit_models()
        total_cost, coupling = self.batch_ot_gmm(
            self.source_model.mean, self.target_model.mean,
            self.source_model.variances.squeeze(), self.target_model.variances.squeeze(),
            weight_source=self.source_model.weights, weight_target=self.target_model.weights, max_iter=100)
        self.transport_matrix = coupling.type_as(self.source_model.mean)
        return total_cost

    @torch.no_grad()
    def transport(self, inputs: Tensor) -> Tensor:
        """reference :82-121"""
        assignments, _, _ = self.source_model.assign(inputs.to(self.dtype))
        target_assignments = assignments @ self.transport_matrix.to(assignments)
        if self.transport_type in ("sample", "argmax"):
            if self.transport_type == "argmax":
                idx = target_assignments.argmax(-1)
            else:
                idx = Categorical(target_assignments / target_assignments.sum(-1, keepdim=True)).sample()
            hard_source = bool(((assignments == 0) | (assignments == 1)).all())
            if hard_source and not self.diag and not self.stochastic and inputs.dim() == 2 and not self.leading_shape:
                return self._transport_by_pair(inputs, assignments.argmax(-1), idx)
            target_assignments = F.one_hot(idx, target_assignments.size(-1)).type_as(target_assignments)
            target_means, target_vars = self.target_model.predict_mean_var(target_assignments)
        elif self.transport_type == "barycenter":
            target_means, target_vars = self.gaussian_barycenter(
                self.target_model.batched_distribution.component_distribution.mean,
                self.target_model.batched_variances, target_assignments, n_iter=100)
        else:
            raise NotImplementedError()
        source_means, source_vars = self.source_model.predict_mean_var(assignments)
        return self.apply_transport(inputs, source_means, target_means,
                                    *self.compute_transport_operators(source_vars, target_vars)).type_as(inputs)

    def _transport_by_pair(self, inputs: Tensor, src_idx: Tensor, tgt_idx: Tensor) -> Tensor:
        """Hard assignments, one operator: the (source component i, target component j) pairs that actually occur get
        their operator T_ij from one batched `compute_transport_operators` call; the inputs of a pair are moved
        together by `apply_transport` (streaming kernel)."""
        ks, kt = self.source_model.n_components, self.target_model.n_components
        pair = src_idx * kt + tgt_idx                                       # [B]
        used = torch.unique(pair)
        cov_s, cov_t = self.source_model.variances, self.target_model.variances
        T, Cw = self.compute_transport_operators(cov_s[used // kt], cov_t[used % kt])     # [P, d, d]
        mean_s, mean_t = self.source_model.mean, self.target_model.mean
        order = torch.argsort(pair, stable=True)
        counts = torch.bincount(torch.searchsorted(used, pair), minlength=used.numel()).tolist()
        out = torch.empty_like(inputs)
        lo = 0
        for p, c in enumerate(counts):
            rows = order[lo:lo + c]
            i, j = int(used[p]) // kt, int(used[p]) % kt
            moved = self.apply_transport(inputs.index_select(0, rows), mean_s[i], mean_t[j], T[p], Cw[p],
                                         batch_dim=-2)
            out.index_copy_(0, rows, moved.type_as(inputs))
            lo += c
        return out

    def extra_repr(self) -> str:
        return super().extra_repr() + W2Mixin.__repr__(self) + f", transport_type={self.transport_type}"
