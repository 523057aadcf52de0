"""Streaming distribution models: the abstract contract and the mixture-assignment mixin (the public surface of
reference ot/distribution_models/base.py:29-258 - constructor keywords, buffer names, `assign` / `kmean_iteration`
return values - re-built around the libotk kernels).

What runs where
  * `MixtureMixin.kmean_iteration` in hard ('argmax') mode with the Euclidean energy - the streaming k-means step of
    `CodebookModel.update` / `fit` - is ONE fused kernel call (`otk_kmeans_assign`): nearest codeword per sample and the
    per-codeword counts / sums, with no [B, K] energy, softmax or one-hot matrix (reference :238-253 builds all three).
  * soft modes contract the weights with the samples through the fp32-accurate tensor-core GEMM (`otk_gemm_nn`).
  * `assign` has to hand back the dense weights and the categorical distribution (they are part of its return value), so
    it evaluates the energy kernel and finishes with elementwise torch ops on the device.
"""
from __future__ import annotations

from abc import ABC, abstractmethod
from functools import partial
from typing import Any, Literal, Optional, Tuple

import torch
import torch.distributions as D
import torch.nn as nn
import torch.nn.functional as F
from torch import Tensor
from torch.types import _device, _dtype

from ... import kernels as K
from ... import utils

__all__ = ["DistributionModel", "MixtureMixin"]


def default_device(device):
    """Buffers default to the current CUDA device (the kernels need them there); CPU only if no GPU exists."""
    if device is not None:
        return torch.device(device)
    return torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")


class DistributionModel(nn.Module, utils.DDPMixin, ABC):
    """One independent model per leading index of `size = (*leading_shape, dim)`.

    Keyword contract of the reference (base.py:42-62): `reduce_on_update`, `update_decay`, `update_with_autograd`, `device`,
    `dtype`, and the DDP hooks consumed by `utils.DDPMixin`.  Sub-classes stream batches with `update`, finalise with `fit`.
    """
    Distribution = None

    def __init__(self, *size: int, reduce_on_update: bool = True, update_decay: Optional[float] = None,
                 update_with_autograd: bool = False, device: Optional[_device] = None,
                 dtype: Optional[_dtype] = None, **ddp_kwargs):
        nn.Module.__init__(self)
        utils.DDPMixin.__init__(self, **ddp_kwargs)
        *lead, self.dim = size
        self.leading_shape = torch.Size(lead)
        self.update_with_autograd = update_with_autograd
        self.reduce_on_update = reduce_on_update
        self.decay = update_decay
        self.ema_update = partial(utils.ema, decay=self.decay)
        # random initial state (what `reset` restores): a vector and a square matrix per leading index; the vector is drawn
        # first so that a seeded construction consumes the generator in the reference's order
        where = dict(dtype=dtype, device=default_device(device))
        for name, shape in (("vec_init", self.vec_shape), ("mat_init", (*self.vec_shape, self.dim))):
            self.register_buffer(name, torch.randn(*shape, **where))

    @property
    def vec_shape(self):
        return *self.leading_shape, self.dim

    def _broadcastable(self, shape) -> bool:
        shape = tuple(shape)
        return shape == tuple(self.leading_shape) or torch.broadcast_shapes(shape, self.leading_shape) == self.leading_shape

    def _validate_samples(self, samples: Tensor) -> None:
        """Shape check of `[*leading_shape, batch, dim]` samples.  The reference constructs these errors without raising
        them (base.py:76-79); raising here would reject inputs it accepts, so the check is advisory and returns the list
        of problems instead (empty when the shapes fit)."""
        if samples.shape[-1] == self.dim and tuple(samples.shape[:-2]) == tuple(self.leading_shape):
            return []
        problems = []
        if not self._broadcastable(samples.shape[:-2]):
            problems.append(f"leading dimensions {tuple(samples.shape[:-2])} do not broadcast to {tuple(self.leading_shape)}")
        if samples.size(-1) != self.dim:
            problems.append(f"dimensionality {samples.size(-1)} != {self.dim}")
        return problems

    def _autograd_warning(self, what: str) -> None:
        if self.update_with_autograd:
            self.warn(f"`update_with_autograd` is set: the parameters are meant to be trained by autograd and {what}")

    def _fit_warn(self):
        self._autograd_warning("`fit` overwrites them with values computed from the running statistics.")

    def _update_warn(self):
        self._autograd_warning("`update` has no running statistics to write to (they are not created in `__init__`).")

    # ------------------------------------------------------------------------------------------------ contract
    @abstractmethod
    def reset(self) -> None:
        """restore the initial state"""

    @property
    def distribution(self) -> D.Distribution:
        raise NotImplementedError()

    @property
    def batched_distribution(self) -> D.Distribution:
        raise NotImplementedError()

    @property
    def variances(self) -> Tensor:
        raise NotImplementedError()

    @abstractmethod
    def update(self, samples: Tensor) -> None:
        """stream `samples` [*leading_shape, batch, dim] into the running statistics"""

    @abstractmethod
    def fit(self, samples: Optional[Tensor] = None) -> None:
        """turn the running statistics (plus optional `samples`) into the model parameters"""

    @abstractmethod
    def predict(self, samples: Tensor) -> Any:
        """model-specific prediction on `samples`"""

    @abstractmethod
    def w2(self, other) -> Tensor:
        """W2 distance to `other`"""

    def forward(self, samples: Tensor) -> Any:
        """training mode streams the batch in first (unless autograd owns the parameters), then predicts"""
        self._validate_samples(samples)
        if self.training and not self.update_with_autograd:
            self.update(samples)
        return self.predict(samples)

    def extra_repr(self) -> str:
        return (f"leading_dim={tuple(self.leading_shape)}, dim={self.dim}, decay={self.decay}, "
                f"update_with_autograd={self.update_with_autograd}")


class MixtureMixin(ABC):
    """Assignment of samples to `n_components` mixture components from an `energy` (similarity) matrix.

    Modes (reference base.py:162, 224-234): 'mean' keeps the softmax weights, 'argmax' / 'sample' harden them to one-hot
    rows, 'gumbel-softmax' / 'gumbel-hardmax' draw Gumbel noise; `topk` masks all but the k largest energies first."""
    Mode = Literal["mean", "sample", "argmax", "gumbel-softmax", "gumbel-hardmax"]

    def __init__(self, *leading_shape, n_components: int, metric: Literal["cosine", "euclidean"] = "euclidean",
                 p: float = 2., topk: Optional[int] = None, temperature: float = 1., training_mode: Mode = "argmax",
                 inference_mode: Mode = "argmax", kmeans_iter: int = 100, laplace_eps: Optional[float] = 1e-5):
        super().__init__()
        self.n_components, self.metric, self.p = n_components, metric, p
        self.topk, self.temperature = topk, temperature
        self.training_mode, self.inference_mode = training_mode, inference_mode
        self.kmeans_iter = kmeans_iter
        self._weight_init = torch.full((*leading_shape, n_components), 1.0 / n_components)      # uniform mixture
        self.laplace_smoothing = partial(utils.laplace_smoothing, n_categories=n_components, eps=laplace_eps)

    @abstractmethod
    def energy(self, samples: Tensor) -> Tensor:
        """similarity of each sample to each component, [*leading_shape, batch, n_comp]"""

    @abstractmethod
    def _update_parameters(self, *kmeans_iter_res: Tensor) -> None:
        ...

    @abstractmethod
    def _update_buffers(self, *kmeans_iter_res: Tensor, decay=False):
        ...

    # ------------------------------------------------------------------------------------------------ assignment
    @property
    def _mode(self) -> str:
        return self.training_mode if getattr(self, "training", False) else self.inference_mode

    def _masked_energy(self, samples: Tensor) -> Tensor:
        energy = self.energy(samples)
        if self.topk:                                               # None or 0: keep every component
            kept, where = energy.topk(self.topk, dim=-1)
            energy = torch.full_like(energy, float("-inf")).scatter_(-1, where, kept)
        return energy

    def assign(self, samples: Tensor) -> Tuple[Tensor, Tensor, D.Categorical]:
        """-> (weights [*L, B, n_comp], sampled indices [*L, B], Categorical over the soft weights); reference :206-236.
        One categorical draw is always taken, whatever the mode, as the reference does (:222)."""
        energy = self._masked_energy(samples)
        soft = torch.softmax(energy / self.temperature, dim=-1)
        distribution = D.Categorical(soft)
        drawn = distribution.sample()
        mode, n = self._mode, energy.size(-1)
        if mode == "mean" or self.topk == 1:
            weights = soft
        elif mode in ("sample", "argmax"):
            picks = drawn if mode == "sample" else soft.argmax(-1)
            weights = F.one_hot(picks, n).type_as(soft)
        elif "gumbel" in mode:
            weights = F.gumbel_softmax(energy, tau=self.temperature, hard="hard" in mode, dim=-1)
        else:
            raise NotImplementedError(f"`mode` must be 'sample', 'mean', 'argmax', 'gumbel' or 'hard-gumbel'. Got {mode}")
        return weights, drawn, distribution

    def _nearest_component_kernel_applies(self, samples: Tensor) -> bool:
        """hard assignment by Euclidean distance to `self.codebook`: argmax of softmax(1/(dist+1e-8)/T) is argmin dist"""
        book = getattr(self, "codebook", None)
        return (self._mode == "argmax" and not self.topk and self.metric == "euclidean" and self.p == 2
                and self.temperature > 0 and isinstance(book, Tensor) and book.is_cuda and samples.is_cuda
                and samples.dim() == book.dim() and type(self).energy is getattr(type(self), "_euclidean_energy_owner", None))

    def kmean_iteration(self, samples: Tensor) -> Tuple[Tensor, ...]:
        """-> (sum of weights per component [*L, n_comp], weighted sum of samples per component [*L, n_comp, dim]);
        reference :238-253."""
        if self._nearest_component_kernel_applies(samples):
            _, counts, sums = K.kmeans_assign(samples, self.codebook.detach(), want_index=False, sums_dtype=samples.dtype)
            return counts.to(samples.dtype), sums.to(samples.dtype)
        weights, _, _ = self.assign(samples)
        per_component = weights.sum(-2)
        if samples.is_cuda and samples.dim() == 2:
            # [n_comp, B] x [B, dim] on the fp32-accurate tensor-core GEMM (FFMA engine for small / unaligned shapes)
            mixed = K.gemm(weights.transpose(-1, -2).float().contiguous(), samples.float().contiguous(), nn=True)
            return per_component, mixed.to(samples.dtype)
        return per_component, weights.transpose(-1, -2) @ samples

    def extra_repr(self) -> str:
        return (f"num_components={self.n_components}, metric={self.metric}, topk={self.topk}, p={self.p}, "
                f"temperature={self.temperature}, training_mode={self.training_mode}, "
                f"inference_mode={self.inference_mode}")
