"""Closed-form Gaussian W2 transport between streamed latent statistics (mirror of reference
ot/transport/gaussian_transport.py:23-98; Freirich-Michaeli-Meir eq. 17/19).

`compute()` for the deterministic full-covariance case is ONE libotk call (`otk_transport_operator`): the source
root / inverse root, the root of Cs^1/2 Ct Cs^1/2, T and W2^2 (whose trace term equals the reference's
tr((Ct^1/2 Cs Ct^1/2)^1/2), same eigenvalues) come out of two Newton-Schulz solves instead of ~14 `eigh`.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn.utils.parametrize as P
from torch import Tensor

from ... import kernels as K
from ..._native import NotConverged
from ..matrix_utils import STABILITY_CONST
from ..distribution_models.gaussian_model import GaussianModel
from ..w2_utils import W2Mixin
from .base import TransportOperator

__all__ = ["GaussianTransport"]


class GaussianTransport(TransportOperator, W2Mixin):
    def __init__(self, *size, source_cfg={}, target_cfg={}, transport_cfg={}, **kwargs):
        W2Mixin.__init__(self, **dict(transport_cfg))
        TransportOperator.__init__(
            self, *size,
            source_model=GaussianModel(*size, w2_cfg=dict(transport_cfg), **source_cfg),
            target_model=GaussianModel(*size, w2_cfg=dict(transport_cfg), **target_cfg),
            **kwargs)
        self.transport_operator = None
        self.cov_stochastic_noise = None

    def update(self, source_samples: Optional[Tensor] = None, target_samples: Optional[Tensor] = None) -> None:
        """reference ot/transport/base.py `update`: one `GaussianModel.update` per side.  A source and a target batch of the
        same shape in the latency regime (the reference's batches of 250) go to the kernel as ONE call
        (`otk_stats_update_pair`); everything else takes the two per-model updates."""
        if source_samples is not None and target_samples is not None and not (self.store_source or self.store_target) \
                and source_samples.dim() == 2 and source_samples.shape == target_samples.shape and source_samples.shape[0] <= 256:
            pa, pb = self.source_model._fast_plan(), self.target_model._fast_plan()
            if pa is not None and pb is not None:
                pair = self.__dict__.get("_pair_plan")
                if pair is None or pair[0] is not pa or pair[1] is not pb:
                    try:
                        pair = (pa, pb, K.StatsUpdatePairPlan(pa, pb))
                    except ValueError:
                        pair = (pa, pb, None)
                    self.__dict__["_pair_plan"] = pair
                if pair[2] is not None and pair[2](source_samples, target_samples):
                    return
        super().update(source_samples, target_samples)

    def reset(self) -> None:
        super().reset()
        self.transport_operator = None
        self.cov_stochastic_noise = None
        self._prepared = None

    def fit_models(self):
        """While `compute()` runs, `_fit_operands = (cov_s_buf, cov_t_buf, shift)` asks the fit kernels to write the
        symmetrised + shifted covariances the map computation consumes as well (one launch per model does the whole fit)."""
        self._prepared = None          # the means are about to change
        operands = getattr(self, "_fit_operands", None)
        src, tgt = self._stored_samples()
        reduced = src is None and tgt is None and self._joint_reduce()
        if operands is None:
            self.source_model.fit(src, already_reduced=reduced)
            self.target_model.fit(tgt, already_reduced=reduced)
            return
        self.source_model.fit(src, cov_operand=operands[0], operand_shift=operands[2], already_reduced=reduced)
        self.target_model.fit(tgt, cov_operand=operands[1], operand_shift=operands[2], already_reduced=reduced)
        self._operands_written = True

    def _joint_reduce(self) -> bool:
        """ONE all-reduce of [n | sum x | sum x x^T] of BOTH models (the reference: three per model and update,
        gaussian_model.py:153-156), written back into the running buffers in place.  False if the two models do not
        share the default reduction or a process group does not exist - each model then reduces for itself."""
        sm, tm = self.source_model, self.target_model
        if self.diag or sm.update_with_autograd or tm.update_with_autograd or sm.reduce is not tm.reduce \
                or not sm._reduce_is_active() or sm._running_sum.dtype != tm._running_sum.dtype:
            return False
        bufs = [sm._n_obs, sm._running_sum, sm._running_sum_cov, tm._n_obs, tm._running_sum, tm._running_sum_cov]
        dt = sm._running_sum_cov.dtype
        flat = sm.reduce(torch.cat([b.reshape(-1).to(dt) for b in bufs]))
        lo = 0
        for b in bufs:
            b.copy_(flat[lo:lo + b.numel()].view(b.shape))
            lo += b.numel()
        return True

    def _prepared_operator(self):
        """`kernels.PreparedTransport` of the current (means, T), rebuilt whenever one of them was replaced or written to
        (tensor identity + version counters), so assigning `transport_operator` or refitting a model cannot go stale."""
        # (looked up through the module / parameter dicts: `nn.Module.__getattr__` costs ~0.4 us per hop, and this runs once
        # per transported batch of 250 latents)
        mods = self._modules
        sm, tm = mods["source_model"], mods["target_model"]
        ms, mt = sm._parameters.get("mean"), tm._parameters.get("mean")
        if ms is None or mt is None:
            ms, mt = sm.mean, tm.mean
        T = self.transport_operator
        key = (id(T), T._version, id(ms), ms._version, id(mt), mt._version,
               sm.__dict__.get("_fit_generation", 0), tm.__dict__.get("_fit_generation", 0))
        cached = self.__dict__.get("_prepared")
        if cached is None or cached[0] != key:
            var_s = self.source_model.parametrizations.cov.original.diagonal(dim1=-2, dim2=-1)
            cached = (key, K.PreparedTransport(ms, mt, T, var_s), (T, ms, mt))   # keep the keyed tensors alive
            self._prepared = cached
        return cached[1]

    def compute(self) -> Tensor:
        """Fit both Gaussians, then W2^2 [*leading_shape] and the operators (reference :64-78)."""
        if not (self.diag or self.stochastic):
            fast = self._compute_full_deterministic()       # fits the models itself (fused with the operand preparation)
            if fast is not None:
                return fast
        else:
            self.fit_models()
        with P.cached():  # evaluate each `.cov` parametrization once for everything below
            mean_s, mean_t = self.source_model.mean, self.target_model.mean
            cov_s, cov_t = self.source_model.cov, self.target_model.cov
            if self.diag or self.stochastic:
                w2 = self.w2_gaussian(mean_s, mean_t, cov_s, cov_t)
                self.transport_operator, self.cov_stochastic_noise = self.compute_transport_operators(cov_s, cov_t)
                return w2
            # covariances read through the parametrization are symmetric and PD by construction, so the
            # reference's per-call validation (5 eigh) cannot fail here
            T, w2 = K.transport_operator(cov_s.to(self.dtype), cov_t.to(self.dtype), pg_star=float(self.pg_star),
                                         mean_s=mean_s.to(self.dtype), mean_t=mean_t.to(self.dtype))
        return self._store(T, w2, cov_s.device)

    def _store(self, T: Tensor, w2: Tensor, device) -> Tensor:
        self.transport_operator = T.to(device=device, dtype=self.dtype)
        zero = getattr(self, "_zero_noise", None)        # Cw = 0 (reference :768): one all-zero tensor, never written to
        if zero is None or zero.shape != T.shape or zero.dtype != self.transport_operator.dtype or zero.device != self.transport_operator.device:
            zero = self._zero_noise = torch.zeros_like(self.transport_operator)
        self.cov_stochastic_noise = zero
        return w2.to(device)

    def _compute_full_deterministic(self):
        """Fast path of the common case.  Reading `.cov` costs a smallest-eigenvalue solve per model only to learn
        that a covariance is PD, in which case the parametrization adds exactly 1e-8 I (reference
        matrix_utils.py:132-139).  The Newton-Schulz solves of the map converge only for PD input, so they are their
        own certificate: run them on triu-mirror(raw) + 1e-8 I and fall back to the eigenvalue repair (return None)
        only if they report an indefinite matrix."""
        sm, tm = self.source_model, self.target_model
        raw_s, raw_t = sm.parametrizations.cov.original, tm.parametrizations.cov.original
        if not raw_s.is_cuda:
            self.fit_models()
            return None
        # operands and results live in buffers that persist across calls: with the same pointers every time libotk
        # replays the whole map computation as one CUDA graph instead of ~100 launches
        buf = getattr(self, "_map_buffers", None)
        if buf is None or buf["cov_s"].shape != raw_s.shape or buf["cov_s"].device != raw_s.device or buf["cov_s"].dtype != self.dtype:
            lead = raw_s.shape[:-2]
            buf = dict(eps=torch.full(lead, STABILITY_CONST, dtype=raw_s.dtype, device=raw_s.device),
                       cov_s=torch.empty(raw_s.shape, dtype=self.dtype, device=raw_s.device),
                       cov_t=torch.empty(raw_s.shape, dtype=self.dtype, device=raw_s.device),
                       T=torch.empty(raw_s.shape, dtype=self.dtype, device=raw_s.device),
                       w2=torch.empty(lead, dtype=torch.float64, device=raw_s.device))
            self._map_buffers = buf
        self._fit_operands = (buf["cov_s"], buf["cov_t"], STABILITY_CONST) if raw_s.dtype == self.dtype else None
        self._operands_written = False
        try:
            self.fit_models()          # an overridden / replaced fit_models simply leaves the operands unwritten
        finally:
            self._fit_operands = None
        if self._operands_written:
            cov_s, cov_t = buf["cov_s"], buf["cov_t"]
        else:
            cov_s = K.symmetrize_shift(raw_s, buf["eps"], out=buf["cov_s"]).to(self.dtype)
            cov_t = K.symmetrize_shift(raw_t, buf["eps"], out=buf["cov_t"]).to(self.dtype)
        try:
            T, w2 = K.transport_operator(cov_s, cov_t, pg_star=float(self.pg_star), mean_s=sm.mean.to(self.dtype),
                                         mean_t=tm.mean.to(self.dtype), out=(buf["T"], buf["w2"]))
        except NotConverged:
            return None
        # hand out copies: the buffers are overwritten by the next compute()
        return self._store(T.clone(), w2.clone(), raw_s.device)

    def transport(self, inputs: Tensor) -> Tensor:
        """[*leading_shape, (B,) dim] -> same shape, dtype and device as `inputs` (reference :80-95)."""
        if inputs.size(-1) != self.dim:
            raise ValueError("`inputs` dimensionality must match the model dimensionality")
        lead = tuple(self.leading_shape)
        if tuple(inputs.shape[:-2]) != lead and tuple(inputs.shape[:-1]) != lead:
            raise ValueError("`inputs` leading dims must match the model batch_shape with optional trailing batch dimensions")
        is_batched = inputs.dim() == len(lead) + 2
        if not (self.diag or self.stochastic) and is_batched:
            # deterministic full map on a batch: straight to the GEMM kernel (Cw == 0, nothing to validate); the
            # operator-only preparation is done once per map, not per batch
            if self.transport_operator.is_cuda and tuple(inputs.shape[:-2]) == tuple(self.transport_operator.shape[:-2]):
                moved = self._prepared_operator().apply(inputs)
            else:
                moved = K.apply_transport(inputs, self.source_model.mean, self.target_model.mean, self.transport_operator)
            if moved.dtype is inputs.dtype and moved.device == inputs.device:
                return moved
            return moved.to(device=inputs.device, dtype=inputs.dtype)
        moved = self.apply_transport(inputs, self.source_model.mean, self.target_model.mean, self.transport_operator,
                                     self.cov_stochastic_noise, batch_dim=-2 if is_batched else None)
        return moved.type_as(inputs)

    def extra_repr(self) -> str:
        return super().extra_repr() + W2Mixin.__repr__(self)
