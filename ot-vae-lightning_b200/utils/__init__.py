"""Host-side seam helpers of the latent-OT path (mirror of the names the reference's `ot/` code imports from
`ot_vae_lightning.utils`, reference utils/__init__.py:21-46, 190-218, 233-328).

The reference reaches `torch.distributed` through pytorch-lightning helpers; here the defaults talk to
`torch.distributed` directly (NCCL over NVLink on the GPU box, gloo in the CPU tests) and stay injectable
through the same three constructor arguments.
"""
from __future__ import annotations

import warnings
from typing import Callable, List, Optional, Sequence

import torch
import torch.distributed as dist
from torch import Tensor

__all__ = [
    "DDPMixin", "ddp_reduce_func_default", "ddp_gather_all_default", "ddp_warn_default",
    "gather_all_if_ddp_available", "ema", "ema_inplace", "laplace_smoothing", "unsqueeze_like",
    "permute_and_flatten", "unflatten_and_unpermute", "human_format",
]

DDPAllGather = Callable[[Tensor], List[Tensor]]
DDPAllReduce = Callable[[Tensor], Tensor]
DDPWarn = Callable[[str], None]


def _dist_on() -> bool:
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def ddp_reduce_func_default(t: Tensor) -> Tensor:
    """SUM all-reduce when a process group exists, identity otherwise (reference utils/__init__.py:32)."""
    if not _dist_on():
        return t
    out = t.clone()
    dist.all_reduce(out, op=dist.ReduceOp.SUM)
    return out


def gather_all_if_ddp_available(t: Tensor) -> List[Tensor]:
    """reference utils/__init__.py:24-27"""
    if not _dist_on():
        return [t]
    outs = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(outs, t.contiguous())
    return outs


ddp_gather_all_default = gather_all_if_ddp_available


def ddp_warn_default(msg: str) -> None:
    if not _dist_on() or dist.get_rank() == 0:
        warnings.warn(msg)


class DDPMixin(object):
    """reference utils/__init__.py:37-46: `reduce`, `gather`, `warn` callables, injectable per instance."""

    def __init__(self, ddp_reduce_func: Optional[DDPAllReduce] = ddp_reduce_func_default,
                 ddp_gather_func: Optional[DDPAllGather] = gather_all_if_ddp_available,
                 ddp_warn_func: Optional[DDPWarn] = ddp_warn_default):
        self.reduce = ddp_reduce_func or (lambda x: x)
        self.gather = ddp_gather_func or (lambda x: [x])
        self.warn = ddp_warn_func or warnings.warn


def ema(moving_avg, new, decay):
    """reference utils/__init__.py:204-206"""
    return moving_avg + new if decay is None else moving_avg * decay + new * (1 - decay)


def ema_inplace(moving_avg, new, decay):
    """reference utils/__init__.py:190-201"""
    if decay is None:
        moving_avg.add_(new)
    else:
        moving_avg.mul_(decay).add_(new, alpha=1 - decay)


def laplace_smoothing(x, n_categories, eps=1e-5):
    """reference utils/__init__.py:209-218"""
    if eps is None:
        return x
    total = x.sum(-1, keepdim=True)
    return (x + eps) / (total + n_categories * eps) * total


def unsqueeze_like(tensor: Tensor, like: Tensor) -> Tensor:
    """Append singleton dims until `tensor.ndim == like.ndim` (reference utils/__init__.py:314-328).
    A python number raises, as in the reference (it has no `.ndim`)."""
    missing = like.ndim - tensor.ndim
    if missing < 0:
        raise ValueError(f"tensor.ndim={tensor.ndim} > like.ndim={like.ndim}")
    return tensor if missing == 0 else tensor[(...,) + (None,) * missing]


def permute_and_flatten(x: Tensor, permute_dims: Sequence[int], batch_first: bool = True,
                        flatten_batch: bool = False) -> Tensor:
    """Layout the caller (`LatentTransport`, reference ot/transport_callback.py:36-43) feeds the operators with:
    the `permute_dims` are moved last and flattened into the feature axis (reference utils/__init__.py:233-267).

    The reference materialises the permuted tensor (`.contiguous()`, :260-261) before every update.  Here the result is
    a strided VIEW whenever the flattening allows one (e.g. ViT tokens [B, T, D] with per-token operators ->
    [T, B, D] with strides (D, T D, 1)): the statistics kernel reads such a layout in place through its TMA descriptor
    (`otk_stats_update` takes row and batch strides), which saves one read and one write of the latents per update.
    Layouts that admit no view are copied exactly as in the reference."""
    others = set(range(1, x.dim()))
    if not others:
        raise ValueError("`input` is expected to have at least 2 dimensions")
    if len(permute_dims) == 0:
        raise ValueError("`permute_dims` is expected to contain at least one dimension")
    if not set(permute_dims).issubset(others):
        raise ValueError("`permute_dims` is expected to be a subset of the `input` dimensions")
    rest = sorted(others - set(permute_dims))
    if not rest:
        return x.flatten(int(not flatten_batch))
    order = (0, *rest, *permute_dims) if batch_first else (*rest, 0, *permute_dims)
    y = x.permute(*order)                         # flatten() = reshape(): a view if the strides allow it, else one copy
    y = y.flatten(int(batch_first and not flatten_batch), len(rest) - int(not batch_first and not flatten_batch))
    return y.flatten(-len(permute_dims))


def unflatten_and_unpermute(xr: Tensor, orig_shape: Sequence[int], permute_dims: Sequence[int],
                            batch_first: bool = True, flatten_batch: bool = False) -> Tensor:
    """Inverse of `permute_and_flatten` (reference utils/__init__.py:270-311)."""
    rest = sorted(set(range(1, len(orig_shape))) - set(permute_dims))
    if not rest:
        return xr.view(*orig_shape)
    feat_shape = [orig_shape[d] for d in permute_dims]
    rest_shape = [orig_shape[d] for d in rest]
    y = xr
    if flatten_batch:
        n_rest = 1
        for r in rest_shape:
            n_rest *= r
        y = y.unflatten(0, [orig_shape[0], n_rest] if batch_first else [n_rest, orig_shape[0]])
    y = y.unflatten(-1, feat_shape)
    y = y.unflatten(int(batch_first), rest_shape)
    where = list(range(len(orig_shape)))
    if not batch_first:
        where[0] = len(rest)
    for dim in range(1, len(orig_shape)):
        where[dim] = rest.index(dim) + int(batch_first) if dim in rest else len(rest) + 1 + list(permute_dims).index(dim)
    return y.permute(*where).contiguous()


def human_format(num) -> str:
    """1234 -> '1.23K' (used by the reference's covariance test printout)."""
    num = float(f"{num:.3g}")
    mag = 0
    while abs(num) >= 1000:
        mag += 1
        num /= 1000.0
    return "{}{}".format(f"{num:f}".rstrip("0").rstrip("."), ["", "K", "M", "B", "T"][mag])
