"""Discrete (codebook-to-codebook) transport with the entropic plan from the Sinkhorn kernels (mirror of reference
ot/transport/discrete_transport.py:27-98).  The cost is the reference's inverse distance `energy`.

`compute()` goes to the point-cloud solver (`otk_sinkhorn_points`, inverse-Euclidean cost tiles built from the two
codebooks on the device): it returns <C, pi> from the solver's own summary pass and keeps only the potentials; the
`transport_matrix` the reference stores eagerly (:60-66) is materialised from them the first time it is read
(`otk_sinkhorn_points_plan`).  Problems whose cost / reg ratio is beyond fp32 (the reference default reg = 1e-5 on an
un-normalised inverse distance) and batched / non-Euclidean codebooks take the dense fp64 solver instead."""
from __future__ import annotations

from typing import Literal

import torch
import torch.nn.functional as F
from torch import Tensor
from torch.distributions import Categorical

from ... import _native as N
from ... import kernels as K
from ..distribution_models.codebook_model import CodebookModel
from ..w2_utils import sinkhorn_log
from .base import TransportOperator

# fp32 potentials resolve the log-domain recurrences while max(C) / reg stays below this (error ~ ratio * 2^-24)
FP32_SAFE_COST_OVER_REG = 1.0e3

__all__ = ["DiscreteTransport"]


class DiscreteTransport(TransportOperator):
    def __init__(self, *size: int, source_cfg={}, target_cfg={}, transport_type: Literal["sample", "argmax", "mean"],
                 sinkhorn_reg: float = 1e-5, sinkhorn_max_iter: int = 1000, sinkhorn_threshold: float = 1e-6, **kwargs):
        super().__init__(*size, source_model=CodebookModel(*size, **source_cfg),
                         target_model=CodebookModel(*size, **target_cfg), **kwargs)
        self.transport_type = transport_type
        self.sinkhorn_reg = sinkhorn_reg
        self.sinkhorn_max_iter = sinkhorn_max_iter
        self.sinkhorn_threshold = sinkhorn_threshold
        self._plan = None            # the materialised plan, or None while only the potentials exist
        self._potentials = None      # (rows, cols, u, v) of the last point-cloud solve

    @property
    def transport_matrix(self):
        """[*L, n, k] entropic plan; built from the potentials on first access after a point-cloud `compute()`."""
        if self._plan is None and self._potentials is not None:
            rows, cols, u, v = self._potentials
            self._plan = K.points_plan(rows, cols, u, v, 1.0, self.sinkhorn_reg, N.COST_INV_EUCLIDEAN).to(
                self.source_model.codebook.dtype)
        return self._plan

    @transport_matrix.setter
    def transport_matrix(self, value):
        self._plan, self._potentials = value, None

    def reset(self) -> None:
        super().reset()
        self.transport_matrix = None

    def _compute_on_points(self):
        """<C, pi> from `otk_sinkhorn_points` (no plan yet), or None if the problem does not qualify.  As in the reference
        (:58-66) the cost matrix is `source.energy(target codebook)`: rows = target codewords, columns = source codewords,
        solved against a = source probabilities, b = target probabilities (hence the equal-size requirement)."""
        sm, tm = self.source_model, self.target_model
        if not (sm.metric == "euclidean" and sm.p == 2 and len(sm.leading_shape) == 0 and sm.codebook.is_cuda
                and sm.n_components == tm.n_components):
            return None
        rows, cols = tm.codebook.detach().float().contiguous(), sm.codebook.detach().float().contiguous()
        if float(K.cost_max(rows, cols, N.COST_INV_EUCLIDEAN)) > FP32_SAFE_COST_OVER_REG * self.sinkhorn_reg:
            return None
        res = K.sinkhorn_points(rows, cols, self.source_distribution.probs, self.target_distribution.probs,
                                reg=self.sinkhorn_reg, max_iter=self.sinkhorn_max_iter, threshold=self.sinkhorn_threshold,
                                cost=N.COST_INV_EUCLIDEAN, scale=1.0, precision=1, want_iters=False)
        self._plan, self._potentials = None, (rows, cols, res["u"], res["v"])
        return res["summary"][0].to(sm.codebook.dtype)

    def compute(self) -> Tensor:
        self.fit_models()
        total = self._compute_on_points()
        if total is not None:
            return total
        cost = self.source_model.energy(self.target_model.codebook)
        self.transport_matrix = sinkhorn_log(self.source_distribution.probs, self.target_distribution.probs, cost,
                                             reg=self.sinkhorn_reg, max_iter=self.sinkhorn_max_iter,
                                             threshold=self.sinkhorn_threshold)
        return (cost * self.transport_matrix).sum(dim=(-2, -1))

    def transport(self, inputs: Tensor) -> Tensor:
        was_training = self.training
        self.eval()
        assignments, _, _ = self.source_model.assign(inputs)
        routed = assignments @ self.transport_matrix  # un-normalised plan rows, as in the reference (:77)
        if self.transport_type == "argmax":
            routed = F.one_hot(routed.argmax(-1), routed.size(-1)).type_as(routed)
        elif self.transport_type == "sample":
            routed = F.one_hot(Categorical(routed).sample(), routed.size(-1)).type_as(routed)
        elif self.transport_type != "mean":
            raise NotImplementedError()
        moved = (routed @ self.target_model.codebook).type_as(inputs)
        self.train(was_training)
        return moved

    def extra_repr(self) -> str:
        return super().extra_repr() + (f"sinkhorn_reg={self.sinkhorn_reg}, sinkhorn_max_iter={self.sinkhorn_max_iter}, "
                                       f"sinkhorn_threshold={self.sinkhorn_threshold}")
