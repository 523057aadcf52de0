"""Discrete codebook model with streaming k-means, kernel-backed (public surface of reference
ot/distribution_models/codebook_model.py:27-214: `CategoricalEmbeddings`, `CodebookModel(*size, mixture_cfg={}, ...)`, the
`codebook` parameter and the `_running_sum` / `_n_obs` / `weight_init` buffers, `update` / `fit` / `predict` / `energy` /
`w2`).

Hot steps
  * `energy`   1 / (|x - c|_2 + 1e-8): the libotk cost-tile kernel (`otk_cost_matrix`, inverse-Euclidean kind);
  * `update` / `fit`  one fused nearest-codeword + per-codeword-sum kernel per k-means step (`otk_kmeans_assign`, through
    `MixtureMixin.kmean_iteration`), then the exponential-moving-average book-keeping on [n_comp] / [n_comp, dim] tensors;
  * `w2`       entropic OT between two codebooks with the Sinkhorn kernels.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributions as D
import torch.nn as nn
import torch.nn.functional as F
from torch import Tensor

from ... import _native as N
from ... import kernels as K
from ..w2_utils import sinkhorn_log
from .base import DistributionModel, MixtureMixin

__all__ = ["CategoricalEmbeddings", "CodebookModel"]

OCCUPIED = 1e-8      # a codeword takes part in an update only if some weight was assigned to it (reference :186,196)


class CategoricalEmbeddings(D.Categorical):
    """A categorical distribution whose outcomes are rows of an embedding table [*batch, n, dim]: `mean` is the probability
    weighted row, `mode` the most likely row, `sample()` a drawn row (reference codebook_model.py:27-66)."""

    def __init__(self, embeddings: Tensor, probs: Optional[Tensor] = None, logits: Optional[Tensor] = None) -> None:
        super().__init__(probs, logits)
        if tuple(self.probs.shape) != tuple(embeddings.shape[:-1]):
            raise ValueError("`probs` and `embeddings` should have the same leading dimensions")
        self.embeddings = embeddings

    def _rows(self, weights: Tensor) -> Tensor:
        """weights [*, n] -> weights . embeddings [*, dim]"""
        return torch.einsum("...n,...nd->...d", weights.type_as(self.embeddings), self.embeddings)

    def _row_at(self, index: Tensor) -> Tensor:
        return self._rows(F.one_hot(index, self._num_events).type_as(index))

    def expand(self, batch_shape, _instance=None):
        out = super().expand(batch_shape, _instance)
        out.embeddings = out.embeddings.expand(torch.Size(batch_shape) + torch.Size((self._num_events,)))
        return out

    @property
    def mean(self):
        return self._rows(self.probs)

    @property
    def mode(self):
        return self._row_at(self.probs.argmax(-1))

    def sample(self, sample_shape=torch.Size()) -> Tensor:
        return self._row_at(super().sample(sample_shape))


class CodebookModel(DistributionModel, MixtureMixin):
    Distribution = CategoricalEmbeddings

    def __init__(self, *size: int, mixture_cfg={}, **kwargs) -> None:
        MixtureMixin.__init__(self, *size[:-1], **mixture_cfg)
        DistributionModel.__init__(self, *size, **kwargs)         # `vec_init` is [*L, n_comp, dim] (vec_shape below)
        self.register_buffer("weight_init", self._weight_init.to(self.vec_init))
        self.codebook = nn.Parameter(self.vec_init.clone(), requires_grad=self.update_with_autograd)
        if not self.update_with_autograd:
            self.register_buffer("_running_sum", torch.zeros_like(self.vec_init))
            self.register_buffer("_n_obs", torch.zeros_like(self.vec_init[..., 0]))

    @property
    def vec_shape(self):
        return *self.leading_shape, self.n_components, self.dim

    @property
    def weights(self) -> Tensor:
        """mixture weights = normalised occupation counts; uniform until something was observed"""
        seen = getattr(self, "_n_obs", None)
        if seen is None or bool(torch.allclose(seen, torch.zeros_like(seen))):
            return self.weight_init.type_as(self.codebook)
        seen = seen.type_as(self.codebook)
        return seen / seen.sum(-1, keepdim=True)

    @property
    def distribution(self) -> Distribution:
        return CategoricalEmbeddings(self.codebook, probs=self.weights)

    @property
    def batched_distribution(self) -> Distribution:
        return CategoricalEmbeddings(self.codebook.unsqueeze(-3), probs=self.weights.unsqueeze(-2))

    @torch.no_grad()
    def reset(self) -> None:
        self.codebook.copy_(self.vec_init)
        if not self.update_with_autograd:
            for buf in (self._running_sum, self._n_obs):
                buf.zero_()

    # ------------------------------------------------------------------------------------------------ streaming k-means
    def _prepared(self, samples: Tensor) -> Tensor:
        self._validate_samples(samples)
        samples = samples.detach().to(self._running_sum)
        self._init_parameters(samples)
        return samples

    @torch.no_grad()
    def update(self, samples: Tensor) -> None:
        """one k-means step on the batch, folded into the running sums with the EMA rule (reference :122-131)"""
        self._update_warn()
        step = self.kmean_iteration(self._prepared(samples))
        if self.reduce_on_update:
            step = tuple(self.reduce(t) for t in step)
        self._update_parameters(*self._update_buffers(*step, decay=True))

    @torch.no_grad()
    def fit(self, samples: Optional[Tensor] = None) -> None:
        """`kmeans_iter` Lloyd iterations on `samples` (or a refresh from the running sums when None), reference :133-146"""
        self._fit_warn()
        if samples is not None:
            samples = self._prepared(samples)
        step = None
        for _ in range(self.kmeans_iter):
            step = self.kmean_iteration(samples)
            self._update_parameters(*(self.reduce(t) for t in step))
        if step is not None:
            self._update_buffers(*step, decay=False)

    def kmean_iteration(self, samples: Optional[Tensor]) -> Tuple[Tensor, ...]:
        if samples is None:                       # nothing new: the accumulated (counts, sums) are the statistics
            return self._n_obs, self._running_sum
        return MixtureMixin.kmean_iteration(self, samples)

    def _update_parameters(self, *kmeans_iter_res: Tensor) -> None:
        """codeword = sum / Laplace-smoothed count, for the occupied codewords only"""
        counts, sums = kmeans_iter_res
        live = counts > OCCUPIED
        smoothed = self.laplace_smoothing(counts[live])
        self.codebook.data[live] = sums[live] / smoothed.unsqueeze(-1)

    def _update_buffers(self, *kmeans_iter_res: Tensor, decay: bool = False):
        counts, sums = kmeans_iter_res
        live = counts > OCCUPIED
        for buf, new in ((self._n_obs, counts), (self._running_sum, sums)):
            buf[live] = self.ema_update(buf[live], new[live]) if decay else new[live]
        return self._n_obs, self._running_sum

    def _init_parameters(self, samples: Tensor) -> None:
        """first batch: the codebook starts from `n_components` distinct rows of it, picked by a HOST-side permutation as
        in the reference (:211), so that a seeded run selects the same rows"""
        if bool(torch.allclose(self.codebook, self.vec_init)):
            rows = torch.randperm(samples.size(-2))[:self.n_components].to(samples.device)
            self.codebook.copy_(samples[..., rows, :])
            self._n_obs += 1

    # ------------------------------------------------------------------------------------------------ energies / distances
    def predict(self, features: Tensor) -> Tuple[Tensor, Tensor, D.Categorical]:
        weights, indices, distribution = self.assign(features)
        return weights @ self.codebook, indices, distribution

    def energy(self, samples: Tensor) -> Tensor:
        """[*L, b, d] -> [*L, b, n_comp]: inverse p-distance ('euclidean') or |cosine| similarity (reference :155-168)"""
        self._validate_samples(samples)
        samples = samples.to(self.codebook)
        book = self.codebook
        if self.metric == "euclidean":
            if self.p == 2 and samples.is_cuda and samples.dim() == book.dim():
                return self._inverse_distance_tiles(samples, book)
            return 1 / (torch.cdist(samples, book, self.p) + 1e-8)
        if self.metric == "cosine":
            p_norm = lambda t: t.abs().pow(self.p).sum(-1)
            overlap = (samples @ book.transpose(-2, -1)).abs()
            return overlap / (p_norm(samples).unsqueeze(-1) * p_norm(book).unsqueeze(-2) + 1e-8) ** (1 / self.p)
        raise NotImplementedError(f"Supported `metric`: 'cosine', 'euclidean'. Got `metric`={self.metric}")

    @staticmethod
    def _inverse_distance_tiles(samples: Tensor, book: Tensor) -> Tensor:
        """one `otk_cost_matrix` launch group per leading index (the kernel takes 2-D clouds)"""
        lead = torch.broadcast_shapes(samples.shape[:-2], book.shape[:-2])
        b, k = samples.shape[-2], book.shape[-2]
        xs = samples.expand(*lead, *samples.shape[-2:]).reshape(-1, b, samples.shape[-1])
        cs = book.expand(*lead, *book.shape[-2:]).reshape(-1, k, book.shape[-1])
        out = torch.stack([K.cost_matrix(x, c, N.COST_INV_EUCLIDEAN) for x, c in zip(xs, cs)])
        return out.reshape(*lead, b, k).to(samples.dtype)

    def w2(self, other: Distribution) -> Tensor:
        """entropic OT cost to another codebook distribution; the ground cost inverts the energy back to a distance
        (reference :170-183: reg 1e-5, 100 iterations, threshold 1e-3)"""
        ground = 1 / (self.energy(other.embeddings) + 1e-8)
        plan = sinkhorn_log(self.distribution.probs, other.probs, ground, reg=1e-5, max_iter=100, threshold=1e-3)
        return (ground * plan).sum(dim=(-2, -1))

    def extra_repr(self) -> str:
        return DistributionModel.extra_repr(self) + ", " + MixtureMixin.extra_repr(self)


# `MixtureMixin` only swaps in the nearest-codeword kernel while `energy` is the Euclidean one defined above
CodebookModel._euclidean_energy_owner = CodebookModel.energy
