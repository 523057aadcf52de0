"""ctypes binding of libotk.so (C ABI declared in include/otk.h).

There is no fallback: if the shared library is missing or the device is not a B200 (sm_100), every
compute entry point raises.  `load(require_gpu=False)` may be used on a CPU-only box to check that the
library loads and exports every declared symbol (tests/test_abi.py); no kernel can run there.
"""
from __future__ import annotations

import ctypes as C
import os
import re
import threading
from typing import Dict, Optional, Tuple

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "csrc", "libotk.so")
HEADER_PATH = os.path.join(os.path.dirname(HERE), "include", "otk.h")

F32, F64 = 0, 1
COST_SQEUCLIDEAN, COST_INV_EUCLIDEAN = 0, 1
OK, ERR_INVALID, ERR_WORKSPACE, ERR_CUDA, ERR_DEVICE, ERR_NOT_CONVERGED = 0, -1, -2, -3, -4, -5

_i64, _int, _dbl, _flt, _ptr, _sz = C.c_int64, C.c_int, C.c_double, C.c_float, C.c_void_p, C.c_size_t

# name -> (restype, argtypes); must list every symbol of include/otk.h (checked by tests/test_abi.py)
SIGNATURES = {
    "otk_abi_version": (_int, []),
    "otk_status_string": (C.c_char_p, [_int]),
    "otk_last_error": (C.c_char_p, []),
    "otk_device_supported": (_int, []),
    "otk_launch_count": (C.c_ulonglong, []),
    "otk_stats_update_workspace_bytes": (_sz, [_i64, _i64, _i64]),
    "otk_stats_update": (_int, [_ptr, _i64, _i64, _i64, _i64, _i64, _dbl, _ptr, _int, _ptr, _ptr, _int, _ptr, _sz, _ptr]),
    "otk_stats_update_f64": (_int, [_ptr, _i64, _i64, _i64, _i64, _i64, _dbl, _ptr, _int, _ptr, _ptr, _int, _ptr, _sz, _ptr]),
    "otk_stats_update_pair": (_int, [_ptr, _ptr, _i64, _i64, _i64, _dbl, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _int, _int, _ptr, _sz, _ptr]),
    "otk_mean_cov": (_int, [_ptr, _ptr, _ptr, _int, _i64, _i64, _ptr, _ptr, _int, _ptr]),
    "otk_gaussian_fit": (_int, [_ptr, _ptr, _int, _ptr, _int, _i64, _i64, _ptr, _ptr, _ptr, _dbl, _int, _ptr]),
    "otk_symmetrize_shift": (_int, [_ptr, _ptr, _i64, _i64, _ptr, _int, _ptr]),
    "otk_asymmetry": (_int, [_ptr, _i64, _i64, _int, _ptr, _ptr]),
    "otk_min_eig_workspace_bytes": (_sz, [_i64, _i64, _int]),
    "otk_min_eig": (_int, [_ptr, _i64, _i64, _int, _int, _ptr, _ptr, _sz, _ptr]),
    "otk_sqrtm_workspace_bytes": (_sz, [_i64, _i64]),
    "otk_sqrtm": (_int, [_ptr, _i64, _i64, _int, _dbl, _int, _int, _ptr, _ptr, _ptr, _sz, _ptr]),
    "otk_w2_gaussian_workspace_bytes": (_sz, [_i64, _i64]),
    "otk_w2_gaussian": (_int, [_ptr, _ptr, _ptr, _ptr, _i64, _i64, _int, _int, _int, _ptr, _ptr, _sz, _ptr]),
    "otk_transport_operator_workspace_bytes": (_sz, [_i64, _i64]),
    "otk_transport_operator": (_int, [_ptr, _ptr, _i64, _i64, _int, _dbl, _int, _int, _ptr, _ptr, _ptr, _ptr, _ptr, _sz, _ptr]),
    "otk_transport_operator_stochastic_workspace_bytes": (_sz, [_i64, _i64]),
    "otk_transport_operator_stochastic": (_int, [_ptr, _ptr, _i64, _i64, _int, _dbl, _int, _int, _ptr, _ptr, _ptr, _sz, _ptr]),
    "otk_apply_transport_workspace_bytes": (_sz, [_i64, _i64, _i64]),
    "otk_apply_transport": (_int, [_ptr, _i64, _i64, _i64, _ptr, _ptr, _ptr, _int, _ptr, _ptr, _sz, _ptr]),
    "otk_transport_prepared_bytes": (_sz, [_i64, _i64]),
    "otk_transport_prepare": (_int, [_ptr, _ptr, _ptr, _ptr, _int, _i64, _i64, _ptr, _sz, _ptr]),
    "otk_apply_transport_prepared": (_int, [_ptr, _i64, _i64, _i64, _ptr, _sz, _ptr, _ptr]),
    "otk_apply_transport_prepared_strided": (_int, [_ptr, _i64, _i64, _i64, _i64, _i64, _ptr, _sz, _ptr, _ptr]),
    "otk_sinkhorn_dense_workspace_bytes": (_sz, [_i64, _i64, _i64]),
    "otk_sinkhorn_dense": (_int, [_ptr, _ptr, _ptr, _i64, _i64, _i64, _int, _dbl, _int, _dbl, _int, _ptr, _ptr, _ptr,
                                  C.POINTER(_int), _ptr, _sz, _ptr]),
    "otk_sinkhorn_points_workspace_bytes": (_sz, [_i64, _i64, _i64, _int]),
    "otk_sinkhorn_points": (_int, [_ptr, _ptr, _i64, _i64, _i64, _ptr, _ptr, _int, _dbl, _int, _dbl, _int, _dbl, _int,
                                   _int, _ptr, _ptr, _ptr, _ptr, _ptr, C.POINTER(_int), _ptr, _sz, _ptr]),
    "otk_sinkhorn_points_colstep": (_int, [_ptr, _ptr, _i64, _i64, _i64, _ptr, _int, _dbl, _dbl, _int, _int, _ptr, _ptr,
                                           _ptr, _sz, _ptr]),
    "otk_lse_combine": (_int, [_ptr, _ptr, _i64, _i64, _i64, _ptr, _ptr, _ptr, _ptr]),
    "otk_sinkhorn_points_rowstep": (_int, [_ptr, _ptr, _i64, _i64, _i64, _ptr, _ptr, _int, _dbl, _dbl, _int, _int, _ptr,
                                           _ptr, _ptr, _sz, _ptr]),
    "otk_sinkhorn_points_summary": (_int, [_ptr, _ptr, _i64, _i64, _i64, _ptr, _ptr, _ptr, _ptr, _int, _dbl, _dbl, _int, _int,
                                           _ptr, _ptr, _ptr, _ptr, _sz, _ptr]),
    "otk_sinkhorn_points_plan": (_int, [_ptr, _ptr, _i64, _i64, _i64, _ptr, _ptr, _int, _dbl, _dbl, _ptr, _ptr, _sz, _ptr]),
    "otk_sinkhorn_exchange_bytes": (_sz, [_int, _i64]),
    "otk_sinkhorn_points_colstep_push": (_int, [_ptr, _ptr, _i64, _i64, _i64, _ptr, _int, _dbl, _dbl, _int, _int, _ptr, _int, _int,
                                                _ptr, _ptr, _sz, _ptr]),
    "otk_sinkhorn_points_sharded_step": (_int, [_ptr, _ptr, _i64, _i64, _i64, _ptr, _ptr, _ptr, _ptr, _int, _dbl, _dbl, _int, _int,
                                                _ptr, _int, _int, _ptr, _ptr, _ptr, _ptr, _sz, _ptr]),
    "otk_lse_combine_wait": (_int, [_ptr, _int, _i64, _ptr, _ptr, _ptr, _ptr, _ptr]),
    "otk_cost_max": (_int, [_ptr, _ptr, _i64, _i64, _i64, _int, _ptr, _ptr, _sz, _ptr]),
    "otk_cost_workspace_bytes": (_sz, [_i64, _i64, _i64]),
    "otk_cost_matrix": (_int, [_ptr, _ptr, _i64, _i64, _i64, _int, _dbl, _ptr, _ptr, _sz, _ptr]),
    "otk_kmeans_assign_workspace_bytes": (_sz, [_i64, _i64, _i64]),
    "otk_kmeans_workspace_bytes": (_sz, [_i64, _i64, _i64, _i64]),
    "otk_kmeans_assign": (_int, [_ptr, _i64, _i64, _i64, _i64, _ptr, _ptr, _ptr, _ptr, _int, _ptr, _sz, _ptr]),
    "otk_microbench_peak": (_int, [_int, C.POINTER(_dbl), _ptr]),
    "otk_gemm_workspace_bytes": (_sz, [_i64, _i64, _i64, _i64]),
    "otk_gemm_nt": (_int, [_ptr, _ptr, _ptr, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _flt, _flt,
                           _int, _ptr, _sz, _ptr]),
    "otk_gemm_nn": (_int, [_ptr, _ptr, _ptr, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _flt, _flt,
                           _int, _ptr, _sz, _ptr]),
}

_lib: Optional[C.CDLL] = None
_lock = threading.Lock()


class NativeError(RuntimeError):
    pass


class NotConverged(NativeError):
    """An iterative kernel (Newton-Schulz) did not converge: the input is not positive definite."""


def declared_symbols(header: str = HEADER_PATH):
    """Function names declared in include/otk.h."""
    with open(header) as f:
        text = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    return sorted(set(re.findall(r"\b(otk_[a-z0-9_]+)\s*\(", text)))


def load(require_gpu: bool = True) -> C.CDLL:
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise NativeError(
                    f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                    "(there is no CPU / PyTorch fallback for this path)")
            lib = C.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(lib, name)  # AttributeError if the .so lacks a declared symbol
                fn.restype, fn.argtypes = res, args
            if lib.otk_abi_version() != 1:
                raise NativeError("libotk.so ABI version mismatch")
            _lib = lib
    if require_gpu:
        if not torch.cuda.is_available():
            raise NativeError("no CUDA device: the latent-OT path is CUDA-only (sm_100a) and has no CPU fallback")
    return _lib


def check(status: int, what: str) -> None:
    if status == OK:
        return
    lib = load(require_gpu=False)
    detail = lib.otk_last_error().decode() or lib.otk_status_string(status).decode()
    if status in (ERR_INVALID, ERR_WORKSPACE):
        raise ValueError(f"{what}: {detail}")
    if status == ERR_NOT_CONVERGED:
        raise NotConverged(f"{what}: {detail}")
    raise NativeError(f"{what}: {lib.otk_status_string(status).decode()} ({detail})")


def dtype_code(dt: torch.dtype) -> int:
    if dt == torch.float32:
        return F32
    if dt == torch.float64:
        return F64
    raise ValueError(f"unsupported dtype {dt}: the native path takes float32 or float64")


def ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def raw_stream(device_index: int) -> int:
    """cudaStream_t of torch's current stream on `device_index` as an int (one C call, no Stream object)."""
    return torch._C._cuda_getCurrentRawStream(device_index)


def stream_ptr(device: torch.device):
    return C.c_void_p(raw_stream(device.index if device.index is not None else torch.cuda.current_device()))


class on_device:
    """`torch.cuda.device(dev)` that costs nothing when `dev` is already current (the per-call overhead matters for
    batches of a few hundred latents); also hands out the current stream once: `.stream` (handle) / `.workspace(n)`."""

    __slots__ = ("device", "_ctx", "stream")

    def __init__(self, device: torch.device):
        self.device = device
        self._ctx = None

    def __enter__(self):
        idx = self.device.index
        cur = torch.cuda.current_device()
        if idx is not None and idx != cur:
            self._ctx = torch.cuda.device(self.device)
            self._ctx.__enter__()
        self.stream = raw_stream(cur if idx is None else idx)
        return self

    def __exit__(self, *exc):
        if self._ctx is not None:
            self._ctx.__exit__(*exc)
        return False

    def stream_ptr(self):
        return C.c_void_p(self.stream)

    def workspace(self, nbytes: int) -> torch.Tensor:
        key = (self.device.index if self.device.index is not None else torch.cuda.current_device(), self.stream)
        buf = _workspaces.get(key)
        if buf is None or buf.numel() < nbytes:
            _workspaces[key] = buf = torch.empty(max(int(nbytes), 1 << 20), dtype=torch.uint8, device=self.device)
        return buf


_workspaces: Dict[Tuple[int, int], torch.Tensor] = {}


def workspace_for(device_index: int, stream: int, nbytes: int, device: torch.device) -> torch.Tensor:
    """`workspace` for callers that already hold the device index and the raw stream handle"""
    buf = _workspaces.get((device_index, stream))
    if buf is None or buf.numel() < nbytes:
        _workspaces[(device_index, stream)] = buf = torch.empty(max(int(nbytes), 1 << 20), dtype=torch.uint8, device=device)
    return buf


def workspace(nbytes: int, device: torch.device) -> torch.Tensor:
    """Per-(device, stream) scratch buffer handed to libotk (which never allocates)."""
    idx = device.index if device.index is not None else torch.cuda.current_device()
    key = (idx, raw_stream(idx))
    buf = _workspaces.get(key)
    if buf is None or buf.numel() < nbytes:
        _workspaces[key] = buf = torch.empty(max(int(nbytes), 1 << 20), dtype=torch.uint8, device=device)
    return buf


def compute_device(*tensors: torch.Tensor) -> torch.device:
    """Device the kernels run on: the first CUDA tensor's device, else the current CUDA device."""
    load(require_gpu=True)
    for t in tensors:
        if isinstance(t, torch.Tensor) and t.is_cuda:
            return t.device
    return torch.device("cuda", torch.cuda.current_device())
