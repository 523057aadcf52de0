"""Discrete (codebook-to-codebook) transport with the entropic plan from the Sinkhorn kernels (mirror of reference
ot/transport/discrete_transport.py:27-98).  The cost is the reference's inverse distance `energy`."""
from __future__ import annotations

from typing import Literal

import torch
import torch.nn.functional as F
from torch import Tensor
from torch.distributions import Categorical

from ..distribution_models.codebook_model import CodebookModel
from ..w2_utils import sinkhorn_log
from .base import TransportOperator

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
        self.transport_matrix = None

    def reset(self) -> None:
        super().reset()
        self.transport_matrix = None

    def compute(self) -> Tensor:
        self.fit_models()
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
