"""Abstract streaming distribution model (mirror of reference ot/distribution_models/base.py:29-158).

`MixtureMixin` (k-means style assignment, reference :161-258) is control-heavy glue around the same kernels and is
kept in stock PyTorch; only what `DiscreteTransport` needs is provided here.
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

from ... import utils

__all__ = ["DistributionModel", "MixtureMixin"]


def default_device(device):
    """Buffers default to the current CUDA device (the kernels need them there); CPU only if no GPU exists."""
    if device is not None:
        return torch.device(device)
    return torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")


class DistributionModel(nn.Module, utils.DDPMixin, ABC):
    """`DistributionModel(*size, reduce_on_update=True, update_decay=None, update_with_autograd=False, device=None,
    dtype=None, **ddp_kwargs)`: `size = (*leading_shape, dim)`; one independent model per leading index."""
    Distribution = None

    def __init__(self, *size: int, reduce_on_update: bool = True, update_decay: Optional[float] = None,
                 update_with_autograd: bool = False, device: Optional[_device] = None,
                 dtype: Optional[_dtype] = None, **ddp_kwargs):
        nn.Module.__init__(self)
        utils.DDPMixin.__init__(self, **ddp_kwargs)
        self.leading_shape = torch.Size(size[:-1])
        self.dim = size[-1]
        self.reduce_on_update = reduce_on_update
        self.decay = update_decay
        self.ema_update = partial(utils.ema, decay=update_decay)
        device = default_device(device)
        self.register_buffer("vec_init", torch.randn(*self.vec_shape, dtype=dtype, device=device))
        self.register_buffer("mat_init", torch.randn(*self.vec_shape, self.dim, dtype=dtype, device=device))
        self.update_with_autograd = update_with_autograd

    @property
    def vec_shape(self):
        return *self.leading_shape, self.dim

    def _broadcastable(self, shape):
        if tuple(shape) == tuple(self.leading_shape):
            return True
        return torch.broadcast_shapes(shape, self.leading_shape) == self.leading_shape

    def _validate_samples(self, samples: Tensor) -> None:
        """The reference builds these errors but never raises them (base.py:76-79); kept non-raising for parity."""
        if not self._broadcastable(samples.shape[:-2]):
            ValueError(f"`samples` leading dimensions are expected to broadcast to {self.leading_shape}")
        if samples.size(-1) != self.dim:
            ValueError(f"`samples` are expected to have dimensionality {self.dim}")

    def _fit_warn(self):
        if self.update_with_autograd:
            self.warn("`self.update_with_autograd` is True: `fit` overrides the trained nn.Parameters with values "
                      "computed from the running statistics.")

    def _update_warn(self):
        if self.update_with_autograd:
            self.warn("`self.update_with_autograd` is True: the running statistics were not created in `__init__`, "
                      "`update` cannot use them.")

    @abstractmethod
    def reset(self) -> None:
        """reset internal model states"""

    @property
    def distribution(self) -> D.Distribution:
        raise NotImplementedError()

    @property
    def batched_distribution(self) -> D.Distribution:
        raise NotImplementedError()

    @property
    def variances(self) -> Tensor:
        raise NotImplementedError()

    def forward(self, samples: Tensor) -> Any:
        self._validate_samples(samples)
        if self.training and not self.update_with_autograd:
            self.update(samples)
        return self.predict(samples)

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

    def extra_repr(self) -> str:
        return (f"leading_dim={tuple(self.leading_shape)}, dim={self.dim}, decay={self.decay}, "
                f"update_with_autograd={self.update_with_autograd}")


class MixtureMixin(ABC):
    """Soft/hard assignment of samples to mixture components (reference base.py:161-258)."""
    Mode = Literal["mean", "sample", "argmax", "gumbel-softmax", "gumbel-hardmax"]

    def __init__(self, *leading_shape, n_components: int, metric: Literal["cosine", "euclidean"] = "euclidean",
                 p: float = 2., topk: Optional[int] = None, temperature: float = 1., training_mode: Mode = "argmax",
                 inference_mode: Mode = "argmax", kmeans_iter: int = 100, laplace_eps: Optional[float] = 1e-5):
        super().__init__()
        self.n_components = n_components
        self.metric = metric
        self.topk = topk
        self.temperature = temperature
        self.training_mode = training_mode
        self.inference_mode = inference_mode
        self.kmeans_iter = kmeans_iter
        self.p = p
        self._weight_init = torch.full((*leading_shape, n_components), 1.0 / n_components)
        self.laplace_smoothing = partial(utils.laplace_smoothing, n_categories=n_components, eps=laplace_eps)

    @abstractmethod
    def energy(self, samples: Tensor) -> Tensor:
        """similarity of each sample to each component, [*leading_shape, batch, n_comp]"""

    def assign(self, samples: Tensor) -> Tuple[Tensor, Tensor, D.Categorical]:
        energy = self.energy(samples)
        if self.topk is not None and self.topk > 0:
            val, idx = torch.topk(energy, self.topk, dim=-1)
            energy = torch.full_like(energy, float("-inf")).scatter_(-1, idx, val)
        weights = torch.softmax(energy / self.temperature, dim=-1)
        distribution = D.Categorical(weights)
        indices = distribution.sample()
        mode = self.training_mode if getattr(self, "training", False) else self.inference_mode
        if mode == "mean" or self.topk == 1:
            pass
        elif mode == "sample":
            weights = F.one_hot(indices, energy.size(-1)).type_as(weights)
        elif mode == "argmax":
            weights = F.one_hot(weights.argmax(-1), energy.size(-1)).type_as(weights)
        elif "gumbel" in mode:
            weights = F.gumbel_softmax(energy, tau=self.temperature, hard="hard" in mode, dim=-1)
        else:
            raise NotImplementedError(f"`mode` must be 'sample', 'mean', 'argmax', 'gumbel' or 'hard-gumbel'. Got {mode}")
        return weights, indices, distribution

    def kmean_iteration(self, samples: Tensor) -> Tuple[Tensor, ...]:
        weights, _, _ = self.assign(samples)
        return weights.sum(-2), weights.transpose(-1, -2) @ samples

    @abstractmethod
    def _update_parameters(self, *kmeans_iter_res: Tensor) -> None:
        ...

    @abstractmethod
    def _update_buffers(self, *kmeans_iter_res: Tensor, decay=False):
        ...

    def extra_repr(self) -> str:
        return (f"num_components={self.n_components}, metric={self.metric}, topk={self.topk}, p={self.p}, "
                f"temperature={self.temperature}, training_mode={self.training_mode}, "
                f"inference_mode={self.inference_mode}")
