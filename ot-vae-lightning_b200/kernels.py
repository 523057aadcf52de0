"""Tensor-level wrappers over the C ABI (one function per libotk entry point).

Inputs may live on the host: they are copied to the compute device (H2D on the current stream) and the
caller decides where results go.  Nothing here computes on the CPU.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch
from torch import Tensor

from . import _native as N


def _dev_tensor(t: Tensor, device: torch.device, dtype: Optional[torch.dtype] = None) -> Tensor:
    if dtype is None:
        dtype = t.dtype if t.dtype in (torch.float32, torch.float64) else torch.float32
    if t.dtype == dtype and t.device == device and not t.requires_grad and t.is_contiguous():
        return t                                   # the common case costs nothing
    return t.detach().to(device=device, dtype=dtype, non_blocking=True).contiguous()


def _lead(shape, tail: int) -> Tuple[torch.Size, int]:
    lead = torch.Size(shape[:len(shape) - tail])
    return lead, int(lead.numel())


def _strided_latents(x: Tensor, dev: torch.device, lead: torch.Size, d: int) -> Tuple[Tensor, int, int]:
    """Latents [*lead, B, d] for the statistics kernel, WITHOUT a copy when they already are fp32 on `dev` with a unit
    feature stride, 16-byte aligned row / leading strides and leading dims that collapse into one (the strided views
    `utils.permute_and_flatten` hands over).  Returns (tensor, row_stride, batch_stride) in elements."""
    if x.dtype == torch.float32 and x.device == dev and not x.requires_grad and x.dim() >= 2 \
            and tuple(x.shape[:-2]) == tuple(lead) and x.shape[-1] == d and x.stride(-1) == 1 and x.shape[-2] > 0:
        try:
            v = x.view(-1, x.shape[-2], d) if x.dim() != 3 else x          # raises if the leading dims do not collapse
        except RuntimeError:
            v = None
        if v is not None and v.data_ptr() % 16 == 0 and v.stride(1) % 4 == 0 and v.stride(1) >= d \
                and (v.shape[0] == 1 or (v.stride(0) % 4 == 0 and v.stride(0) > 0)):
            return v, int(v.stride(1)), int(v.stride(0)) if v.shape[0] > 1 else int(v.shape[1] * v.stride(1))
    x = _dev_tensor(x, dev, torch.float32)
    if x.shape[:-2] != lead:
        x = x.expand(*lead, *x.shape[-2:]).contiguous()
    return x, d, x.shape[-2] * d


_stats_ws_bytes = {}      # (L, rows, d) -> otk_stats_update_workspace_bytes (one ctypes call less per update)


def stats_update(x: Tensor, n_obs: Tensor, run_sum: Tensor, run_cov: Tensor, decay: Optional[float]) -> None:
    """In-place update of the running buffers (all on one CUDA device) with latents x [*L, B, d] (fp32)."""
    dev = run_sum.device
    d = run_sum.shape[-1]
    lib = N.load()
    if x.dim() == 2 and run_sum.dim() == 1 and x.dtype == torch.float32 and x.device == dev and x.is_contiguous() \
            and not x.requires_grad and x.shape[1] == d and x.shape[0] > 0 and x.data_ptr() % 16 == 0:
        # the common call (one model, a contiguous fp32 batch on the model's device): nothing to normalise
        L, rows, row_stride, batch_stride, entry = 1, x.shape[0], d, x.shape[0] * d, lib.otk_stats_update
    else:
        lead, L = _lead(run_sum.shape, 1)
        if x.dtype == torch.float64:
            # fp64 latents keep fp64 products (otk_stats_update_f64): what the reference's einsum on
            # `samples.type_as(buffer)` / FID's `features.double()` computes; everything else is streamed as fp32 through the
            # tensor-core kernels
            x = _dev_tensor(x, dev, torch.float64)
            if x.shape[:-2] != lead:
                x = x.expand(*lead, *x.shape[-2:]).contiguous()
            row_stride, batch_stride, entry = d, x.shape[-2] * d, lib.otk_stats_update_f64
        else:
            x, row_stride, batch_stride = _strided_latents(x, dev, lead, d)
            entry = lib.otk_stats_update
        rows = x.shape[-2]
    if not (run_sum.is_cuda and n_obs.is_contiguous() and run_sum.is_contiguous() and run_cov.is_contiguous()):
        raise ValueError("running buffers must be contiguous CUDA tensors")
    key = (L, rows, d)
    need = _stats_ws_bytes.get(key)
    if need is None:
        need = _stats_ws_bytes[key] = lib.otk_stats_update_workspace_bytes(L, rows, d)
    with N.on_device(dev) as ctx:
        ws = ctx.workspace(need)
        st = entry(x.data_ptr(), L, rows, d, row_stride, batch_stride, -1.0 if decay is None else float(decay),
                   n_obs.data_ptr(), N.dtype_code(n_obs.dtype), run_sum.data_ptr(), run_cov.data_ptr(),
                   N.dtype_code(run_sum.dtype), ws.data_ptr(), ws.numel(), ctx.stream)
    if st != N.OK:
        N.check(st, "otk_stats_update")


class StatsUpdatePlan:
    """`stats_update` for one model with everything that does not change between batches bound once (buffer pointers,
    dtype codes, the library entry): the call that remains is a handful of checks on the batch and one ctypes call."""

    __slots__ = ("dev", "dev_index", "d", "decay", "n_ptr", "n_code", "s_ptr", "c_ptr", "b_code", "entry", "ws_query")

    def __init__(self, n_obs: Tensor, run_sum: Tensor, run_cov: Tensor, decay: Optional[float]):
        lib = N.load()
        self.dev, self.d = run_sum.device, run_sum.shape[-1]
        self.dev_index = run_sum.device.index if run_sum.device.index is not None else torch.cuda.current_device()
        self.decay = -1.0 if decay is None else float(decay)
        self.n_ptr, self.n_code = n_obs.data_ptr(), N.dtype_code(n_obs.dtype)
        self.s_ptr, self.c_ptr, self.b_code = run_sum.data_ptr(), run_cov.data_ptr(), N.dtype_code(run_sum.dtype)
        self.entry, self.ws_query = lib.otk_stats_update, lib.otk_stats_update_workspace_bytes

    def __call__(self, x: Tensor) -> bool:
        d = self.d
        if x.dtype != torch.float32 or x.device != self.dev or x.shape[1] != d or not x.is_contiguous() or x.requires_grad \
                or x.shape[0] == 0 or x.data_ptr() % 16 or torch.cuda.current_device() != self.dev_index:
            return False
        rows = x.shape[0]
        key = (1, rows, d)
        need = _stats_ws_bytes.get(key)
        if need is None:
            need = _stats_ws_bytes[key] = self.ws_query(1, rows, d)
        stream = N.raw_stream(self.dev_index)
        ws = N.workspace_for(self.dev_index, stream, need, self.dev)
        st = self.entry(x.data_ptr(), 1, rows, d, d, rows * d, self.decay, self.n_ptr, self.n_code, self.s_ptr, self.c_ptr,
                        self.b_code, ws.data_ptr(), ws.numel(), stream)
        if st != N.OK:
            N.check(st, "otk_stats_update")
        return True


class StatsUpdatePairPlan:
    """One `otk_stats_update_pair` call for the source and the target batch of a `TransportOperator.update` (two models of
    the same width, dtype and decay): in the latency regime both updates are ONE launch.  Built from the two models'
    `StatsUpdatePlan`s; `__call__` returns False when the batches do not qualify (the per-model path takes them)."""

    __slots__ = ("a", "b", "entry")

    def __init__(self, a: StatsUpdatePlan, b: StatsUpdatePlan):
        if (a.dev, a.d, a.decay, a.n_code, a.b_code) != (b.dev, b.d, b.decay, b.n_code, b.b_code):
            raise ValueError("the two models differ in device, width, decay or buffer dtypes")
        self.a, self.b, self.entry = a, b, N.load().otk_stats_update_pair

    def __call__(self, xa: Tensor, xb: Tensor) -> bool:
        a = self.a
        d = a.d
        if xa.shape != xb.shape or xa.dim() != 2 or xa.shape[1] != d or xa.shape[0] == 0:
            return False
        for x in (xa, xb):
            if x.dtype != torch.float32 or x.device != a.dev or not x.is_contiguous() or x.requires_grad or x.data_ptr() % 16:
                return False
        if torch.cuda.current_device() != a.dev_index:
            return False
        rows = xa.shape[0]
        key = (1, rows, d)
        need = _stats_ws_bytes.get(key)
        if need is None:
            need = _stats_ws_bytes[key] = a.ws_query(1, rows, d)
        stream = N.raw_stream(a.dev_index)
        ws = N.workspace_for(a.dev_index, stream, need, a.dev)
        st = self.entry(xa.data_ptr(), xb.data_ptr(), rows, d, d, a.decay, a.n_ptr, a.s_ptr, a.c_ptr, self.b.n_ptr, self.b.s_ptr,
                        self.b.c_ptr, a.n_code, a.b_code, ws.data_ptr(), ws.numel(), stream)
        if st != N.OK:
            N.check(st, "otk_stats_update_pair")
        return True


def mean_cov(run_sum: Tensor, run_cov: Tensor, n_obs: Tensor) -> Tuple[Tensor, Tensor]:
    dev = N.compute_device(run_sum)
    dt = run_sum.dtype if run_sum.dtype in (torch.float32, torch.float64) else torch.float32
    s, c = _dev_tensor(run_sum, dev, dt), _dev_tensor(run_cov, dev, dt)
    d = s.shape[-1]
    lead, L = _lead(s.shape, 1)
    n = _dev_tensor(n_obs, dev, torch.float64)
    # one count per leading index; a count of shape [1] next to a single [d] / [d, d] pair (the FID states,
    # reference metrics/fid.py:96-97,127-128) broadcasts the way the reference's `unsqueeze_like` division does
    n = (n.reshape(lead) if n.numel() == L else n.expand(lead)).contiguous()
    mean, cov = torch.empty_like(s), torch.empty_like(c)
    with torch.cuda.device(dev):
        st = N.load().otk_mean_cov(N.ptr(s), N.ptr(c), N.ptr(n), N.F64, L, d, N.ptr(mean), N.ptr(cov),
                                   N.dtype_code(dt), N.stream_ptr(dev))
    N.check(st, "otk_mean_cov")
    return mean, cov


def gaussian_fit(n_obs: Tensor, run_sum: Tensor, run_cov: Tensor, mean: Tensor, cov_raw: Tensor,
                 cov_sym: Optional[Tensor] = None, shift: float = 0.0) -> None:
    """In-place fit of mean / raw covariance (and optionally the symmetrised + shifted covariance) from the running buffers;
    all tensors contiguous on one CUDA device, mean / cov_raw / cov_sym of one dtype (otk_gaussian_fit)."""
    dev = run_sum.device
    d = run_sum.shape[-1]
    _, L = _lead(run_sum.shape, 1)
    with N.on_device(dev) as ctx:
        st = N.load().otk_gaussian_fit(run_sum.data_ptr(), run_cov.data_ptr(), N.dtype_code(run_sum.dtype), n_obs.data_ptr(),
                                       N.dtype_code(n_obs.dtype), L, d, mean.data_ptr(), cov_raw.data_ptr(),
                                       None if cov_sym is None else cov_sym.data_ptr(), float(shift),
                                       N.dtype_code(mean.dtype), ctx.stream)
    if st != N.OK:
        N.check(st, "otk_gaussian_fit")


def symmetrize_shift(a: Tensor, shift: Optional[Tensor], out: Optional[Tensor] = None) -> Tensor:
    dev = N.compute_device(a)
    a = _dev_tensor(a, dev)
    d = a.shape[-1]
    lead, L = _lead(a.shape, 2)
    sh = None if shift is None else _dev_tensor(shift, dev, a.dtype).expand(lead).contiguous()
    if out is None or out.shape != a.shape or out.dtype != a.dtype or out.device != a.device or not out.is_contiguous():
        out = torch.empty_like(a)
    with torch.cuda.device(dev):
        st = N.load().otk_symmetrize_shift(N.ptr(a), N.ptr(sh), L, d, N.ptr(out), N.dtype_code(a.dtype),
                                           N.stream_ptr(dev))
    N.check(st, "otk_symmetrize_shift")
    return out


def asymmetry(a: Tensor) -> Tensor:
    dev = N.compute_device(a)
    a = _dev_tensor(a, dev)
    d = a.shape[-1]
    lead, L = _lead(a.shape, 2)
    out = torch.empty(lead, dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        st = N.load().otk_asymmetry(N.ptr(a), L, d, N.dtype_code(a.dtype), N.ptr(out), N.stream_ptr(dev))
    N.check(st, "otk_asymmetry")
    return out


def min_eig(a: Tensor, steps: int = 0) -> Tensor:
    dev = N.compute_device(a)
    a = _dev_tensor(a, dev)
    d = a.shape[-1]
    lead, L = _lead(a.shape, 2)
    out = torch.empty(lead, dtype=torch.float64, device=dev)
    lib = N.load()
    with torch.cuda.device(dev):
        ws = N.workspace(lib.otk_min_eig_workspace_bytes(L, d, steps), dev)
        st = lib.otk_min_eig(N.ptr(a), L, d, N.dtype_code(a.dtype), steps, N.ptr(out), N.ptr(ws), ws.numel(),
                             N.stream_ptr(dev))
    N.check(st, "otk_min_eig")
    return out


def sqrtm_pair(a: Tensor, want_root: bool = True, want_iroot: bool = True, ridge: float = 0.0, iters: int = 0
               ) -> Tuple[Optional[Tensor], Optional[Tensor]]:
    dev = N.compute_device(a)
    a = _dev_tensor(a, dev)
    d = a.shape[-1]
    lead, L = _lead(a.shape, 2)
    root = torch.empty_like(a) if want_root else None
    iroot = torch.empty_like(a) if want_iroot else None
    lib = N.load()
    with torch.cuda.device(dev):
        ws = N.workspace(lib.otk_sqrtm_workspace_bytes(L, d), dev)
        st = lib.otk_sqrtm(N.ptr(a), L, d, N.dtype_code(a.dtype), float(ridge), int(iters), 0, N.ptr(root),
                           N.ptr(iroot), N.ptr(ws), ws.numel(), N.stream_ptr(dev))
    N.check(st, "otk_sqrtm")
    return root, iroot


def _bcast(t: Tensor, lead: torch.Size, tail: int) -> Tensor:
    return t.expand(*lead, *t.shape[t.dim() - tail:]).contiguous()


def w2_gaussian(mean_s: Tensor, mean_t: Tensor, cov_s: Tensor, cov_t: Tensor, iters: int = 0) -> Tensor:
    dev = N.compute_device(cov_s, cov_t, mean_s, mean_t)
    dt = torch.float64 if any(t.dtype == torch.float64 for t in (mean_s, mean_t, cov_s, cov_t)) else torch.float32
    ms, mt, cs, ct = (_dev_tensor(t, dev, dt) for t in (mean_s, mean_t, cov_s, cov_t))
    d = cs.shape[-1]
    lead = torch.broadcast_shapes(ms.shape[:-1], mt.shape[:-1], cs.shape[:-2], ct.shape[:-2])
    L = int(lead.numel())
    ms, mt, cs, ct = _bcast(ms, lead, 1), _bcast(mt, lead, 1), _bcast(cs, lead, 2), _bcast(ct, lead, 2)
    out = torch.empty(lead, dtype=torch.float64, device=dev)
    lib = N.load()
    with torch.cuda.device(dev):
        ws = N.workspace(lib.otk_w2_gaussian_workspace_bytes(L, d), dev)
        st = lib.otk_w2_gaussian(N.ptr(ms), N.ptr(mt), N.ptr(cs), N.ptr(ct), L, d, N.dtype_code(dt), int(iters), 0,
                                 N.ptr(out), N.ptr(ws), ws.numel(), N.stream_ptr(dev))
    N.check(st, "otk_w2_gaussian")
    return out


def transport_operator(cov_s: Tensor, cov_t: Tensor, pg_star: float = 0.0, mean_s: Optional[Tensor] = None,
                       mean_t: Optional[Tensor] = None, iters: int = 0, out: Optional[Tuple[Tensor, Tensor]] = None
                       ) -> Tuple[Tensor, Optional[Tensor]]:
    """T [*L,d,d] (dtype of the covariances) and, if means are given, W2^2 [*L] (fp64) from the same roots.
    `out = (T, w2)`: caller-owned result buffers.  A caller that passes the same operand and result tensors again lets
    libotk replay the whole call as one CUDA graph (its graph cache is keyed on the pointers)."""
    dev = N.compute_device(cov_s, cov_t)
    dt = torch.float64 if (cov_s.dtype == torch.float64 or cov_t.dtype == torch.float64) else torch.float32
    cs, ct = _dev_tensor(cov_s, dev, dt), _dev_tensor(cov_t, dev, dt)
    d = cs.shape[-1]
    lead = torch.broadcast_shapes(cs.shape[:-2], ct.shape[:-2])
    L = int(lead.numel())
    cs, ct = _bcast(cs, lead, 2), _bcast(ct, lead, 2)
    T = w2 = None
    if out is not None and out[0].shape == cs.shape and out[0].dtype == dt and out[0].device == dev and out[0].is_contiguous():
        T = out[0]
    if T is None:
        T = torch.empty_like(cs)
    ms = mt = None
    if mean_s is not None and mean_t is not None:
        ms, mt = _bcast(_dev_tensor(mean_s, dev, dt), lead, 1), _bcast(_dev_tensor(mean_t, dev, dt), lead, 1)
        if out is not None and out[1] is not None and out[1].shape == lead and out[1].dtype == torch.float64 and out[1].device == dev:
            w2 = out[1]
        else:
            w2 = torch.empty(lead, dtype=torch.float64, device=dev)
    lib = N.load()
    with torch.cuda.device(dev):
        ws = N.workspace(lib.otk_transport_operator_workspace_bytes(L, d), dev)
        st = lib.otk_transport_operator(N.ptr(cs), N.ptr(ct), L, d, N.dtype_code(dt), float(pg_star), int(iters), 0,
                                        N.ptr(T), N.ptr(ms), N.ptr(mt), N.ptr(w2), N.ptr(ws), ws.numel(),
                                        N.stream_ptr(dev))
    N.check(st, "otk_transport_operator")
    return T, w2


def transport_operator_stochastic(cov_s: Tensor, cov_t: Tensor, pg_star: float = 0.0) -> Tuple[Tensor, Tensor]:
    """(T, Cw) of eq. 19 [*L, d, d] in the dtype of the covariances (otk_transport_operator_stochastic)."""
    dev = N.compute_device(cov_s, cov_t)
    dt = torch.float64 if (cov_s.dtype == torch.float64 or cov_t.dtype == torch.float64) else torch.float32
    cs, ct = _dev_tensor(cov_s, dev, dt), _dev_tensor(cov_t, dev, dt)
    d = cs.shape[-1]
    lead = torch.broadcast_shapes(cs.shape[:-2], ct.shape[:-2])
    L = int(lead.numel())
    cs, ct = _bcast(cs, lead, 2), _bcast(ct, lead, 2)
    T, Cw = torch.empty_like(cs), torch.empty_like(cs)
    lib = N.load()
    with torch.cuda.device(dev):
        ws = N.workspace(lib.otk_transport_operator_stochastic_workspace_bytes(L, d), dev)
        st = lib.otk_transport_operator_stochastic(N.ptr(cs), N.ptr(ct), L, d, N.dtype_code(dt), float(pg_star), 0, 0,
                                                   N.ptr(T), N.ptr(Cw), N.ptr(ws), ws.numel(), N.stream_ptr(dev))
    N.check(st, "otk_transport_operator_stochastic")
    return T, Cw


def apply_transport(x: Tensor, mean_s: Tensor, mean_t: Tensor, T: Tensor) -> Tensor:
    """y = T (x - mean_s) + mean_t; x [*L, B, d] any float dtype/device -> fp32 result on the compute device."""
    dev = N.compute_device(T, mean_s, x)
    dt = T.dtype if T.dtype in (torch.float32, torch.float64) else torch.float32
    Td, ms, mt = _dev_tensor(T, dev, dt), _dev_tensor(mean_s, dev, dt), _dev_tensor(mean_t, dev, dt)
    xd = _dev_tensor(x, dev, torch.float32)
    d = xd.shape[-1]
    lead = torch.broadcast_shapes(xd.shape[:-2], Td.shape[:-2], ms.shape[:-1], mt.shape[:-1])
    L = int(lead.numel())
    rows = xd.shape[-2]
    xd = _bcast(xd, lead, 2)
    Td, ms, mt = _bcast(Td, lead, 2), _bcast(ms, lead, 1), _bcast(mt, lead, 1)
    y = torch.empty_like(xd)
    lib = N.load()
    with torch.cuda.device(dev):
        ws = N.workspace(lib.otk_apply_transport_workspace_bytes(L, rows, d), dev)
        st = lib.otk_apply_transport(N.ptr(xd), L, rows, d, N.ptr(ms), N.ptr(mt), N.ptr(Td), N.dtype_code(dt),
                                     N.ptr(y), N.ptr(ws), ws.numel(), N.stream_ptr(dev))
    N.check(st, "otk_apply_transport")
    return y


class PreparedTransport:
    """Operator-only work of `apply_transport` done once (reference flow: GaussianTransport.compute builds the map,
    transport() applies it to many batches): fp32 casts, TF32 and scaled-FP16 planes of T, per-feature input scales from
    the source variances.  `apply(x)` is then one kernel launch (+ its device-gated TF32 fallback)."""

    def __init__(self, mean_s: Tensor, mean_t: Tensor, T: Tensor, var_s: Tensor):
        dev = N.compute_device(T, mean_s)
        dt = T.dtype if T.dtype in (torch.float32, torch.float64) else torch.float32
        Td, ms, mt, vs = (_dev_tensor(t, dev, dt) for t in (T, mean_s, mean_t, var_s))
        self.lead = torch.broadcast_shapes(Td.shape[:-2], ms.shape[:-1], mt.shape[:-1], vs.shape[:-1])
        self.L, self.d, self.device = int(torch.Size(self.lead).numel()), Td.shape[-1], dev
        Td, ms, mt, vs = _bcast(Td, self.lead, 2), _bcast(ms, self.lead, 1), _bcast(mt, self.lead, 1), _bcast(vs, self.lead, 1)
        lib = N.load()
        with torch.cuda.device(dev):
            self.state = torch.empty(lib.otk_transport_prepared_bytes(self.L, self.d), dtype=torch.uint8, device=dev)
            st = lib.otk_transport_prepare(N.ptr(ms), N.ptr(mt), N.ptr(Td), N.ptr(vs), N.dtype_code(dt), self.L, self.d,
                                           N.ptr(self.state), self.state.numel(), N.stream_ptr(dev))
        N.check(st, "otk_transport_prepare")
        self._entry, self._state_ptr, self._state_bytes = lib.otk_apply_transport_prepared_strided, self.state.data_ptr(), self.state.numel()

    def apply(self, x: Tensor) -> Tensor:
        """x [*lead, B, d] (any float dtype / device) -> fp32 [*lead, B, d] on the compute device.  fp32 views with a unit
        feature stride (the strided token / channel views of `utils.permute_and_flatten`) are read in place."""
        if tuple(x.shape[:-2]) != tuple(self.lead) or x.shape[-1] != self.d:
            raise ValueError("PreparedTransport.apply: input does not match the operator's leading shape / dimension")
        if x.shape[-2] == 0:
            return torch.empty(x.shape, dtype=torch.float32, device=self.device)
        if x.dim() == 2 and x.dtype == torch.float32 and x.device == self.device and x.is_contiguous() and not x.requires_grad \
                and x.data_ptr() % 16 == 0:
            xd, row_stride, batch_stride = x, self.d, x.shape[0] * self.d          # the common call: nothing to normalise
        else:
            xd, row_stride, batch_stride = _strided_latents(x, self.device, torch.Size(self.lead), self.d)
        y = torch.empty(x.shape, dtype=torch.float32, device=self.device)
        with N.on_device(self.device) as ctx:
            st = self._entry(xd.data_ptr(), self.L, x.shape[-2], self.d, row_stride, batch_stride, self._state_ptr,
                             self._state_bytes, y.data_ptr(), ctx.stream)
        if st != N.OK:
            N.check(st, "otk_apply_transport_prepared_strided")
        return y


def sinkhorn_dense(a: Tensor, b: Tensor, Cm: Tensor, reg: float, max_iter: int, threshold: float,
                   want_plan: bool = True, poll_every: int = 16):
    """Returns (plan or None, u, v, iterations) on the compute device, dtype of C (fp32/fp64)."""
    dev = N.compute_device(Cm, a, b)
    dt = torch.float64 if Cm.dtype == torch.float64 else torch.float32
    Cd = _dev_tensor(Cm, dev, dt)
    n, m = Cd.shape[-2:]
    lead = torch.broadcast_shapes(Cd.shape[:-2], a.shape[:-1], b.shape[:-1])
    L = int(lead.numel())
    Cd = _bcast(Cd, lead, 2)
    ad, bd = _bcast(_dev_tensor(a, dev, dt), lead, 1), _bcast(_dev_tensor(b, dev, dt), lead, 1)
    u, v = torch.empty_like(ad), torch.empty_like(bd)
    plan = torch.empty_like(Cd) if want_plan else None
    iters = C.c_int(0)
    lib = N.load()
    with torch.cuda.device(dev):
        ws = N.workspace(lib.otk_sinkhorn_dense_workspace_bytes(L, n, m), dev)
        st = lib.otk_sinkhorn_dense(N.ptr(ad), N.ptr(bd), N.ptr(Cd), L, n, m, N.dtype_code(dt), float(reg),
                                    int(max_iter), float(threshold), int(poll_every), N.ptr(u), N.ptr(v), N.ptr(plan),
                                    C.byref(iters), N.ptr(ws), ws.numel(), N.stream_ptr(dev))
    N.check(st, "otk_sinkhorn_dense")
    return plan, u, v, iters.value


def sinkhorn_points(x: Tensor, y: Tensor, a: Tensor, b: Tensor, reg: float, max_iter: int, threshold: float = 0.0,
                    cost: int = N.COST_SQEUCLIDEAN, scale: Optional[float] = None, precision: int = 0,
                    poll_every: int = 16, want_summary: bool = True, want_iters: bool = True):
    """Point-cloud Sinkhorn (no N x M matrix in HBM on the fused engine).  scale=None -> 1/max cost.
    Returns dict(u, v, summary=[<C,pi>, mass, max row err, max col err] or None, row_marginal, col_marginal, iters)."""
    dev = N.compute_device(x, y)
    xd, yd = _dev_tensor(x, dev, torch.float32), _dev_tensor(y, dev, torch.float32)
    ad, bd = _dev_tensor(a, dev, torch.float32), _dev_tensor(b, dev, torch.float32)
    n, d = xd.shape
    m = yd.shape[0]
    u = torch.zeros(n, dtype=torch.float32, device=dev)
    v = torch.zeros(m, dtype=torch.float32, device=dev)
    summary = torch.zeros(4, dtype=torch.float64, device=dev) if want_summary else None
    row_marg = torch.empty(n, dtype=torch.float32, device=dev) if want_summary else None
    col_marg = torch.empty(m, dtype=torch.float32, device=dev) if want_summary else None
    iters = C.c_int(0)
    lib = N.load()
    with torch.cuda.device(dev):
        ws = N.workspace(lib.otk_sinkhorn_points_workspace_bytes(n, m, d, int(cost)), dev)
        st = lib.otk_sinkhorn_points(N.ptr(xd), N.ptr(yd), n, m, d, N.ptr(ad), N.ptr(bd), int(cost),
                                     1.0 if scale is None else float(scale), 1 if scale is None else 0, float(reg),
                                     int(max_iter), float(threshold), int(poll_every), int(precision), N.ptr(u),
                                     N.ptr(v), N.ptr(summary), N.ptr(row_marg), N.ptr(col_marg), C.byref(iters) if want_iters else None, N.ptr(ws),
                                     ws.numel(), N.stream_ptr(dev))
    N.check(st, "otk_sinkhorn_points")
    return dict(u=u, v=v, summary=summary, row_marginal=row_marg, col_marginal=col_marg, iters=iters.value)


def points_plan(x: Tensor, y: Tensor, u: Tensor, v: Tensor, scale: float, reg: float, cost: int = N.COST_SQEUCLIDEAN) -> Tensor:
    """plan [N, M] = exp(u_i + v_j - scale * cost(x_i, y_j) / reg) from the potentials of `sinkhorn_points` (fp32)."""
    dev = N.compute_device(x, y)
    xd, yd = _dev_tensor(x, dev, torch.float32), _dev_tensor(y, dev, torch.float32)
    n, d = xd.shape
    m = yd.shape[0]
    plan = torch.empty(n, m, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        ws = N.workspace((n + m) * 4 + 512, dev)
        st = N.load().otk_sinkhorn_points_plan(N.ptr(xd), N.ptr(yd), n, m, d, N.ptr(u), N.ptr(v), int(cost), float(scale),
                                               float(reg), N.ptr(plan), N.ptr(ws), ws.numel(), N.stream_ptr(dev))
    N.check(st, "otk_sinkhorn_points_plan")
    return plan


def cost_matrix(x: Tensor, y: Tensor, cost: int, scale: float = 1.0) -> Tensor:
    dev = N.compute_device(x, y)
    xd, yd = _dev_tensor(x, dev, torch.float32), _dev_tensor(y, dev, torch.float32)
    n, d = xd.shape
    m = yd.shape[0]
    out = torch.empty(n, m, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        ws = N.workspace(N.load().otk_cost_workspace_bytes(n, m, d), dev)
        st = N.load().otk_cost_matrix(N.ptr(xd), N.ptr(yd), n, m, d, int(cost), float(scale), N.ptr(out), N.ptr(ws),
                                      ws.numel(), N.stream_ptr(dev))
    N.check(st, "otk_cost_matrix")
    return out


def cost_max(x: Tensor, y: Tensor, cost: int) -> Tensor:
    dev = N.compute_device(x, y)
    xd, yd = _dev_tensor(x, dev, torch.float32), _dev_tensor(y, dev, torch.float32)
    n, d = xd.shape
    m = yd.shape[0]
    out = torch.zeros(1, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        ws = N.workspace(max((n + m) * 4 + 512, N.load().otk_sinkhorn_points_workspace_bytes(n, m, d, int(cost))
                             if cost == N.COST_SQEUCLIDEAN and d <= 128 and d % 4 == 0 else 0), dev)
        st = N.load().otk_cost_max(N.ptr(xd), N.ptr(yd), n, m, d, int(cost), N.ptr(out), N.ptr(ws), ws.numel(),
                                   N.stream_ptr(dev))
    N.check(st, "otk_cost_max")
    return out


def points_workspace(n: int, m: int, d: int, cost: int, device) -> Tensor:
    """A dedicated workspace for a sequence of colstep / rowstep calls that reuse the prepared operands."""
    return torch.empty(N.load().otk_sinkhorn_points_workspace_bytes(n, m, d, int(cost)), dtype=torch.uint8, device=device)


def colstep(x_local: Tensor, y: Tensor, u_local: Tensor, scale: float, reg: float, cost: int = N.COST_SQEUCLIDEAN,
            precision: int = 0, out: Optional[Tensor] = None, ws: Optional[Tensor] = None, reuse: int = 0
            ) -> Tuple[Tensor, Tensor]:
    """Partial column LSE over the local rows: (max, sumexp), written into `out` [2, M] if given.  `ws` + `reuse`: the
    caller's dedicated workspace still holds the operands prepared by an earlier colstep / rowstep on the same clouds
    (reuse = 1), and the previous call was the previous Sinkhorn iteration (reuse = 2: bounded-shift mode)."""
    dev = x_local.device
    n, d = x_local.shape
    m = y.shape[0]
    if out is None:
        out = torch.empty(2, m, dtype=torch.float32, device=dev)
    cm, cs = out[0], out[1]
    lib = N.load()
    with torch.cuda.device(dev):
        if ws is None:
            ws, reuse = N.workspace(lib.otk_sinkhorn_points_workspace_bytes(n, m, d, int(cost)), dev), 0
        st = lib.otk_sinkhorn_points_colstep(N.ptr(x_local), N.ptr(y), n, m, d, N.ptr(u_local), int(cost), float(scale),
                                             float(reg), int(precision), int(reuse), N.ptr(cm), N.ptr(cs), N.ptr(ws),
                                             ws.numel(), N.stream_ptr(dev))
    N.check(st, "otk_sinkhorn_points_colstep")
    return cm, cs


def colstep_push(x_local: Tensor, y: Tensor, u_local: Tensor, scale: float, reg: float, peers: Tensor, world: int, rank: int,
                 ctrl: Tensor, ws: Tensor, reuse: int = 0, cost: int = N.COST_SQEUCLIDEAN, precision: int = 0) -> None:
    """colstep + push of this rank's column partials into every peer's exchange buffer (`peers`: int64 device tensor [world] of
    the peers' symmetric-memory mappings; `ctrl`: int32 device tensor [4], zeroed once per solver instance)."""
    dev = x_local.device
    n, d = x_local.shape
    m = y.shape[0]
    with torch.cuda.device(dev):
        st = N.load().otk_sinkhorn_points_colstep_push(N.ptr(x_local), N.ptr(y), n, m, d, N.ptr(u_local), int(cost), float(scale),
                                                       float(reg), int(precision), int(reuse), N.ptr(peers), int(world),
                                                       int(rank), N.ptr(ctrl), N.ptr(ws), ws.numel(), N.stream_ptr(dev))
    N.check(st, "otk_sinkhorn_points_colstep_push")


def sharded_step(x_local: Tensor, y: Tensor, a_local: Tensor, b: Tensor, u_local: Tensor, v: Tensor, scale: float, reg: float,
                 stage: int, peers: Tensor, world: int, rank: int, xchg: Tensor, ctrl: Tensor, diffs: Optional[Tensor], ws: Tensor,
                 cost: int = N.COST_SQEUCLIDEAN, precision: int = 0) -> None:
    """one whole row-sharded iteration (otk_sinkhorn_points_sharded_step): five launches, exchange over peer memory"""
    dev = x_local.device
    n, d = x_local.shape
    m = y.shape[0]
    with torch.cuda.device(dev):
        st = N.load().otk_sinkhorn_points_sharded_step(N.ptr(x_local), N.ptr(y), n, m, d, N.ptr(a_local), N.ptr(b), N.ptr(u_local),
                                                       N.ptr(v), int(cost), float(scale), float(reg), int(precision), int(stage),
                                                       N.ptr(peers), int(world), int(rank), N.ptr(xchg), N.ptr(ctrl), N.ptr(diffs),
                                                       N.ptr(ws), ws.numel(), N.stream_ptr(dev))
    N.check(st, "otk_sinkhorn_points_sharded_step")


def lse_combine_wait(xchg: Tensor, world: int, b: Tensor, v: Tensor, diff: Optional[Tensor], ctrl: Tensor) -> None:
    """wait for every rank's partials of the current iteration in the local exchange buffer, then v = log(b + 1e-8) - LSE"""
    dev = v.device
    with torch.cuda.device(dev):
        st = N.load().otk_lse_combine_wait(N.ptr(xchg), int(world), v.shape[0], N.ptr(b), N.ptr(v), N.ptr(diff), N.ptr(ctrl),
                                           N.stream_ptr(dev))
    N.check(st, "otk_lse_combine_wait")


def points_fused_eligible(n: int, m: int, d: int, cost: int = N.COST_SQEUCLIDEAN, precision: int = 0) -> bool:
    """whether `sinkhorn_points` / the half-steps run on the fused tcgen05 engine for this shape (squared-Euclidean cost,
    d a multiple of 8 up to 128, precision 0) - the engine the peer-memory exchange is built into"""
    return precision == 0 and cost == N.COST_SQEUCLIDEAN and 8 <= d <= 128 and d % 8 == 0 and n >= 1 and m >= 1


def exchange_bytes(world: int, m: int) -> int:
    return int(N.load().otk_sinkhorn_exchange_bytes(int(world), int(m)))


def lse_combine(part_max: Tensor, part_sum: Tensor, b: Tensor, v: Tensor, diff: Optional[Tensor]) -> None:
    """v = log(b + 1e-8) - LSE over the parts; part_max / part_sum are [parts, M] views with a common row stride."""
    dev = v.device
    parts, m = part_max.shape
    stride = part_max.stride(0) if parts > 1 else m
    assert part_max.stride(1) == 1 and part_sum.stride(1) == 1 and (parts == 1 or part_sum.stride(0) == stride)
    with torch.cuda.device(dev):
        st = N.load().otk_lse_combine(N.ptr(part_max), N.ptr(part_sum), parts, stride, m, N.ptr(b), N.ptr(v), N.ptr(diff),
                                      N.stream_ptr(dev))
    N.check(st, "otk_lse_combine")


def rowstep(x_local: Tensor, y: Tensor, a_local: Tensor, v: Tensor, u_local: Tensor, diff: Optional[Tensor],
            scale: float, reg: float, cost: int = N.COST_SQEUCLIDEAN, precision: int = 0, ws: Optional[Tensor] = None,
            reuse: int = 0) -> None:
    dev = x_local.device
    n, d = x_local.shape
    m = y.shape[0]
    lib = N.load()
    with torch.cuda.device(dev):
        if ws is None:
            ws, reuse = N.workspace(lib.otk_sinkhorn_points_workspace_bytes(n, m, d, int(cost)), dev), 0
        st = lib.otk_sinkhorn_points_rowstep(N.ptr(x_local), N.ptr(y), n, m, d, N.ptr(a_local), N.ptr(v), int(cost),
                                             float(scale), float(reg), int(precision), int(reuse), N.ptr(u_local),
                                             N.ptr(diff), N.ptr(ws), ws.numel(), N.stream_ptr(dev))
    N.check(st, "otk_sinkhorn_points_rowstep")


def points_summary(x_local: Tensor, y: Tensor, a_local: Tensor, b: Tensor, u_local: Tensor, v: Tensor, scale: float,
                   reg: float, cost: int = N.COST_SQEUCLIDEAN, precision: int = 0, ws: Optional[Tensor] = None,
                   reuse: bool = False):
    """Plan statistics of a row shard (no plan in memory): returns (part [4] fp64 = <C,pi>, mass, max row error, local max
    |col_partial - b|; row_marginal [n_local]; col_partial [M] = sum over the local rows of pi_ij)."""
    dev = x_local.device
    n, d = x_local.shape
    m = y.shape[0]
    part = torch.zeros(4, dtype=torch.float64, device=dev)
    row_marg = torch.empty(n, dtype=torch.float32, device=dev)
    col_part = torch.empty(m, dtype=torch.float32, device=dev)
    lib = N.load()
    with torch.cuda.device(dev):
        if ws is None:
            ws, reuse = N.workspace(lib.otk_sinkhorn_points_workspace_bytes(n, m, d, int(cost)), dev), 0
        st = lib.otk_sinkhorn_points_summary(N.ptr(x_local), N.ptr(y), n, m, d, N.ptr(a_local), N.ptr(b), N.ptr(u_local),
                                             N.ptr(v), int(cost), float(scale), float(reg), int(precision), int(bool(reuse)),
                                             N.ptr(part), N.ptr(row_marg), N.ptr(col_part), N.ptr(ws), ws.numel(),
                                             N.stream_ptr(dev))
    N.check(st, "otk_sinkhorn_points_summary")
    return part, row_marg, col_part


def kmeans_assign(x: Tensor, codebook: Tensor, want_index: bool = True, want_sums: bool = True,
                  sums_dtype: Optional[torch.dtype] = None):
    """Nearest-codeword step (otk_kmeans_assign).  x [*L, B, d], codebook [*L, K, d] (leading dims broadcast) ->
    (index int64 [*L, B] or None, weights_sum [*L, K] or None, samples_sum [*L, K, d] or None)."""
    dev = N.compute_device(codebook, x)
    lead = torch.broadcast_shapes(x.shape[:-2], codebook.shape[:-2])
    L = int(torch.Size(lead).numel())
    xd = _bcast(_dev_tensor(x, dev, torch.float32), lead, 2)
    cd = _bcast(_dev_tensor(codebook, dev, torch.float32), lead, 2)
    B, d = xd.shape[-2:]
    Kc = cd.shape[-2]
    dt = sums_dtype if sums_dtype in (torch.float32, torch.float64) else torch.float32
    index = torch.empty(*lead, B, dtype=torch.int64, device=dev) if want_index else None
    wsum = torch.empty(*lead, Kc, dtype=dt, device=dev) if want_sums else None
    ssum = torch.empty(*lead, Kc, d, dtype=dt, device=dev) if want_sums else None
    lib = N.load()
    with torch.cuda.device(dev):
        ws = N.workspace(lib.otk_kmeans_workspace_bytes(L, B, Kc, d), dev)
        st = lib.otk_kmeans_assign(N.ptr(xd), L, B, Kc, d, N.ptr(cd), N.ptr(index), N.ptr(wsum), N.ptr(ssum), N.dtype_code(dt),
                                   N.ptr(ws), ws.numel(), N.stream_ptr(dev))
    N.check(st, "otk_kmeans_assign")
    return index, wsum, ssum


def gemm(A: Tensor, B: Tensor, alpha: float = 1.0, engine: int = 0, nn: bool = False) -> Tensor:
    """C = alpha * A @ B^T (nn=False, B [*, N, K]) or alpha * A @ B (nn=True, B [*, K, N]); fp32.
    engine: 0 auto, 1 FFMA, 2 tcgen05 3xTF32, 3 tcgen05 1xTF32.  Exported for the kernel unit tests."""
    dev = A.device
    A, B = A.contiguous(), B.contiguous()
    M, Kd = A.shape[-2:]
    Nn = B.shape[-1] if nn else B.shape[-2]
    batch = int(torch.Size(A.shape[:-2]).numel())
    out = torch.empty(*A.shape[:-2], M, Nn, dtype=torch.float32, device=dev)
    lib = N.load()
    fn = lib.otk_gemm_nn if nn else lib.otk_gemm_nt
    with torch.cuda.device(dev):
        ws = N.workspace(lib.otk_gemm_workspace_bytes(M, Nn, Kd, batch), dev)
        st = fn(N.ptr(A), N.ptr(B), N.ptr(out), M, Nn, Kd, Kd, B.shape[-1], Nn, batch, M * Kd, B.shape[-2] * B.shape[-1],
                M * Nn, float(alpha), 0.0, int(engine), N.ptr(ws), ws.numel(), N.stream_ptr(dev))
    N.check(st, "otk_gemm")
    return out
