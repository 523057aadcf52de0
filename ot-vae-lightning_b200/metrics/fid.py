"""Streaming Frechet Inception Distance statistics on B200 (mirror of the accumulation part of reference
metrics/fid.py:27-131: same state names, `update(generated=, samples=)` / `compute()`).

In scope: the fp64 running sum / correlation of the features (`features.T @ features`, reference
:103-104,115-121) - the same libotk statistics kernel as `GaussianModel` - and `mean_cov`.  Out of scope: the
Inception network (torchmetrics' `NoTrainInceptionV3`, a third-party dependency absent from this image) - pass any
feature extractor as `net`.  The final score (torchmetrics' `_compute_fid`, fid.py:130:
|m1-m2|^2 + tr C1 + tr C2 - 2 sum sqrt(eig(C1 C2))) is the Gelbrich distance and is evaluated on the device by
`otk_w2_gaussian`; covariances of fewer observations than features are exactly singular, which the fp64 Newton-Schulz
engine handles through its relative ridge (csrc/matfun.cu, NS_F64_REL_RIDGE).
"""
from __future__ import annotations

from typing import Any, Callable, Optional, Tuple

import torch
import torch.distributed as dist
import torch.nn as nn
from torch import Tensor

from .. import kernels as K
from ..ot.matrix_utils import mean_cov

__all__ = ["FrechetInceptionDistance"]


class FrechetInceptionDistance(nn.Module):
    higher_is_better = False

    def __init__(self, net: Optional[nn.Module] = None, feature_size: int = 2048, to_255: Optional[bool] = False,
                 data_range: Optional[Tuple[float, float]] = (0., 1.), compute_on_step: Optional[bool] = False,
                 dist_sync_on_step: Optional[bool] = False, process_group: Optional[Any] = None,
                 dist_sync_fn: Callable = None, device=None):
        super().__init__()
        if net is None:
            raise ValueError("torchmetrics' InceptionV3 is not part of this build: pass a feature extractor as `net` "
                             f"(output [B, {feature_size}])")
        self.net = net.eval()
        self.to_255 = to_255
        self.data_range = data_range[1] - data_range[0]
        self.data_low = data_range[0]
        self.process_group = process_group
        dev = torch.device(device) if device is not None else (
            torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu"))
        for kind in ("real", "fake"):
            self.register_buffer(f"{kind}_sum", torch.zeros(feature_size, dtype=torch.double, device=dev))
            self.register_buffer(f"{kind}_correlation",
                                 torch.zeros(feature_size, feature_size, dtype=torch.double, device=dev))
            self.register_buffer(f"num_{kind}_obs", torch.zeros(1, dtype=torch.long, device=dev))
        self.register_buffer("_count_scratch", torch.zeros((), dtype=torch.double, device=dev), persistent=False)

    def reset(self) -> None:
        for name, buf in self.named_buffers(recurse=False):
            buf.zero_()

    @torch.no_grad()
    def _features(self, img: Tensor) -> Tensor:
        if img.size(1) == 1:
            img = torch.cat([img, img, img], dim=1)
        if self.to_255:
            img = (255 * (img - self.data_low) / self.data_range).type(torch.uint8)
        return self.net(img).double().reshape(img.shape[0], -1)          # fp64 features, as fid.py:101

    @torch.no_grad()
    def _accumulate(self, img: Tensor, kind: str) -> None:
        feats = self._features(img.to(getattr(self, f"{kind}_sum").device))
        self._count_scratch.zero_()
        K.stats_update(feats, self._count_scratch, getattr(self, f"{kind}_sum"), getattr(self, f"{kind}_correlation"),
                       None)
        getattr(self, f"num_{kind}_obs").add_(img.shape[0])

    def update(self, generated: Optional[Tensor] = None, samples: Optional[Tensor] = None) -> None:
        """NB the reference stores `generated` under the `real_*` states and `samples` under `fake_*`
        (metrics/fid.py:113-122); kept as is."""
        if generated is not None:
            self._accumulate(generated, "real")
        if samples is not None:
            self._accumulate(samples, "fake")

    def _synced(self, t: Tensor) -> Tensor:
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.process_group) > 1:
            t = t.clone()
            dist.all_reduce(t, group=self.process_group)
        return t

    def compute(self) -> Tensor:
        n_real, n_fake = self._synced(self.num_real_obs), self._synced(self.num_fake_obs)
        if n_fake < 1e3 or n_real < 1000:
            return torch.ones(1) * float("inf")
        r_mean, r_cov = mean_cov(self._synced(self.real_sum), self._synced(self.real_correlation), n_real)
        f_mean, f_cov = mean_cov(self._synced(self.fake_sum), self._synced(self.fake_correlation), n_fake)
        return K.w2_gaussian(r_mean, f_mean, r_cov, f_cov).reshape(())

    def forward(self, *args, **kwargs):
        self.update(*args, **kwargs)
