"""Host-buffer streaming helpers: feed latents that live in (pinned) host memory through a `TransportOperator`
with the H2D copy of chunk i+1, the kernels of chunk i and the D2H copy of chunk i-1 overlapped on three CUDA streams.

This is plumbing around the public API (`op.update`, `op.transport`); the arithmetic is unchanged.  It is what
`bench.py` times for the `e2e` figure.
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import Tensor


class _Ring:
    """Two device staging buffers + events for a copy/compute pipeline."""

    def __init__(self, rows: int, dim: int, device: torch.device, n: int = 2):
        self.bufs = [torch.empty(rows, dim, dtype=torch.float32, device=device) for _ in range(n)]
        self.filled = [torch.cuda.Event() for _ in range(n)]
        self.freed = [torch.cuda.Event() for _ in range(n)]
        self.n = n


def stream_update(op, source: Optional[Tensor] = None, target: Optional[Tensor] = None, chunk: int = 1 << 16,
                  device: Optional[torch.device] = None) -> None:
    """`op.update(source_samples=..., target_samples=...)` over host tensors [N, d], chunk by chunk."""
    ref = source if source is not None else target
    dev = device or next(op.buffers()).device
    n, d = ref.shape
    copy = torch.cuda.Stream(dev)
    main = torch.cuda.current_stream(dev)
    rings = {k: _Ring(min(chunk, n), d, dev) for k, t in (("s", source), ("t", target)) if t is not None}
    # the rings come from the caching allocator on `main`: blocks it hands back may still be read by kernels queued on
    # `main` (e.g. the previous call's last update), so the copy stream may not touch them before `main` gets here
    copy.wait_stream(main)
    host = dict(s=source, t=target)
    n_chunks = (n + chunk - 1) // chunk

    def launch_copy(i):
        slot = i % 2
        lo, hi = i * chunk, min(n, (i + 1) * chunk)
        with torch.cuda.stream(copy):
            for k, ring in rings.items():
                if i >= 2:
                    copy.wait_event(ring.freed[slot])
                ring.bufs[slot][:hi - lo].copy_(host[k][lo:hi], non_blocking=True)
                ring.filled[slot].record(copy)

    launch_copy(0)
    for i in range(n_chunks):
        if i + 1 < n_chunks:
            launch_copy(i + 1)
        slot = i % 2
        rows = min(n, (i + 1) * chunk) - i * chunk
        kw = {}
        for k, ring in rings.items():
            main.wait_event(ring.filled[slot])
            kw["source_samples" if k == "s" else "target_samples"] = ring.bufs[slot][:rows]
        op.update(**kw)
        for ring in rings.values():
            ring.freed[slot].record(main)
    main.wait_stream(copy)      # every side-stream access to the rings is ordered before `main` frees them


def stream_transport(op, inputs: Tensor, out: Tensor, chunk: int = 1 << 16, device: Optional[torch.device] = None) -> Tensor:
    """out[...] = op.transport(inputs) for host tensors [N, d]; H2D / kernels / D2H overlapped.  Returns after the last
    D2H copy has landed in `out` (the D2H stream is synchronised), so the host may read it at once."""
    dev = device or next(op.buffers()).device
    n, d = inputs.shape
    h2d, d2h = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    main = torch.cuda.current_stream(dev)
    ring_in = _Ring(min(chunk, n), d, dev)
    outs = [None, None]
    done = [torch.cuda.Event() for _ in range(2)]
    stored = [torch.cuda.Event() for _ in range(2)]
    n_chunks = (n + chunk - 1) // chunk
    h2d.wait_stream(main)       # see stream_update: the ring blocks may still be in use on `main`
    d2h.wait_stream(main)

    def launch_copy(i):
        slot = i % 2
        lo, hi = i * chunk, min(n, (i + 1) * chunk)
        with torch.cuda.stream(h2d):
            if i >= 2:
                h2d.wait_event(ring_in.freed[slot])
            ring_in.bufs[slot][:hi - lo].copy_(inputs[lo:hi], non_blocking=True)
            ring_in.filled[slot].record(h2d)

    launch_copy(0)
    for i in range(n_chunks):
        if i + 1 < n_chunks:
            launch_copy(i + 1)
        slot = i % 2
        lo, hi = i * chunk, min(n, (i + 1) * chunk)
        main.wait_event(ring_in.filled[slot])
        if i >= 2:
            main.wait_event(stored[slot])          # the previous result in this slot has left the device
        outs[slot] = op.transport(ring_in.bufs[slot][:hi - lo])
        ring_in.freed[slot].record(main)
        done[slot].record(main)
        with torch.cuda.stream(d2h):
            d2h.wait_event(done[slot])
            out[lo:hi].copy_(outs[slot], non_blocking=True)
            stored[slot].record(d2h)
    main.wait_stream(h2d)
    main.wait_stream(d2h)
    d2h.synchronize()           # `out` is host memory: the caller reads it without any stream in between
    return out
