"""Multi-GPU drivers for the two parts of the path that shard: one process per GPU,
`torch.distributed` (NCCL over NVLink on the GPU box, gloo in the CPU tests) for the plumbing.

* sufficient statistics: every rank streams its own latents (`reduce_on_update=False`), `GaussianModel.fit()` does
  ONE packed all-reduce of [n | sum x | sum x x^T]  (see GaussianModel._packed_reduce);
* Sinkhorn: source rows are sharded, y / b / v are replicated; per iteration one all-gather of the
  column (max, sum-exp) partials [2, M] - the only data-path collective - and one scalar all-reduce for the stop rule.
"""
from __future__ import annotations

import os
from typing import Optional

import torch
import torch.distributed as dist
from torch import Tensor

from . import kernels as K


def _world(group=None) -> int:
    return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1


def shard_rows(n: int, rank: int, world: int):
    """Contiguous row range of `rank` (first ranks take the remainder)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def global_cost_scale(x_local: Tensor, y: Tensor, cost: int = 0, group=None) -> float:
    """1 / max_ij cost over ALL ranks' rows (the normalisation of reference w2_utils.py:265-266)."""
    mx = K.cost_max(x_local, y, cost)
    if _world(group) > 1:
        dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=group)
    return 1.0 / float(mx.item())


def sharded_sinkhorn(x_local: Tensor, y: Tensor, a_local: Tensor, b: Tensor, reg: float, max_iter: int,
                     threshold: float = 0.0, scale: Optional[float] = None, cost: int = 0, precision: int = 0,
                     poll_every: int = 16, group=None, kernels=K, use_graph: Optional[bool] = None,
                     plan: Optional[dict] = None):
    """Row-sharded log-domain Sinkhorn (same recurrences / stop rule as reference w2_utils.py:301-319).
    x_local [n_g, d], a_local [n_g] are this rank's rows; y [M, d], b [M] are replicated.
    Returns dict(u_local, v, iters, scale).  `kernels` is injectable for the CPU (gloo) tests.

    One iteration = column half-step on the local rows -> all-gather of the [2, M] (max, sumexp) partials (the only
    data-path collective) -> combine -> row half-step.  On CUDA the iteration (kernels + the NCCL all-gather) is
    captured once into a CUDA graph and replayed, so the host enqueues one launch per iteration instead of ~10.
    Capturing a graph that contains a collective costs tens of milliseconds (as much as the 100 iterations of the
    benchmark at 8 GPUs): the result carries the `plan` (graph + the buffers it is bound to), and a caller that solves
    the same problem again (same operand tensors and parameters) passes it back to skip the capture.  The plan is owned
    by the caller on purpose - NCCL cannot destroy a communicator while a graph that captured it is alive, so it must be
    dropped before `destroy_process_group()`; nothing is cached behind the caller's back."""
    world = _world(group)
    dev = x_local.device
    if scale is None:
        scale = global_cost_scale(x_local, y, cost, group) if kernels is K else kernels.global_cost_scale(x_local, y)
    m = y.shape[0]
    want_graph = use_graph
    if want_graph is None:   # capture costs about as much as a few iterations: only worth it for long runs
        want_graph = (kernels is K and dev.type == "cuda" and max_iter >= 32
                      and os.environ.get("OTK_SINKHORN_GRAPH", "1") != "0")
    key = (x_local.data_ptr(), y.data_ptr(), a_local.data_ptr(), b.data_ptr(), tuple(x_local.shape), tuple(y.shape),
           float(scale), float(reg), int(cost), int(precision), world, id(group), str(dev))
    if plan is not None and plan.get("key") != key:
        plan = None                      # a plan of another problem: ignore it
    if plan is None:
        plan = dict(u=torch.zeros(x_local.shape[0], dtype=torch.float32, device=dev),
                    v=torch.zeros(m, dtype=torch.float32, device=dev),
                    diffs=torch.zeros(2, dtype=torch.float32, device=dev),     # [sum|du| local, sum|dv| replicated]
                    part=torch.empty(2, m, dtype=torch.float32, device=dev),   # this rank's column (max, sumexp)
                    graph=None, refused=False, key=key)
        plan["gathered"] = (torch.empty(world, 2, m, dtype=torch.float32, device=dev) if world > 1
                            else plan["part"].unsqueeze(0))
        # the operands (FP16 planes, norms) are prepared by the first half-step and then reused from a dedicated workspace
        plan["ws"] = (kernels.points_workspace(x_local.shape[0], m, x_local.shape[1], cost, dev)
                      if hasattr(kernels, "points_workspace") else None)
        plan["peer"] = None
        if kernels is K and dev.type == "cuda":
            fused = K.points_fused_eligible(x_local.shape[0], m, x_local.shape[1], cost, precision)
            if world > 1:
                peer = _peer_exchange(world, m, dev, group)
                # the exchange only works if EVERY rank has it (and the fused engine takes the shape): agree once
                ok = torch.tensor([1 if (peer is not None and fused) else 0], device=dev)
                dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
                plan["peer"] = peer if int(ok.item()) == 1 else None
            elif fused and os.environ.get("OTK_SINKHORN_PEER", "1") != "0":
                # one rank: the same five-launch iteration with the "exchange" buffer in local memory
                xchg = torch.zeros(K.exchange_bytes(1, m), dtype=torch.uint8, device=dev)
                plan["peer"] = dict(xchg=xchg, hdl=None, ptrs=torch.tensor([xchg.data_ptr()], dtype=torch.int64, device=dev),
                                    ctrl=torch.zeros(8, dtype=torch.int32, device=dev))
    else:
        plan["u"].zero_()
        plan["v"].zero_()
    u, v, diffs, part, gathered, ws = (plan[k] for k in ("u", "v", "diffs", "part", "gathered", "ws"))
    peer = plan.get("peer")
    rank = dist.get_rank(group) if world > 1 else 0

    def iteration(stage: int) -> None:
        """stage 0: first iteration (the half-steps prepare the operands); 1: second iteration; 2: steady state - both
        half-steps find the previous iteration's biases and partial LSEs in `ws` and run in bounded-shift mode"""
        if peer is not None:
            # the whole iteration in five launches: the column partials go straight into every peer's exchange buffer from
            # the half-step's last kernel, the combine kernel waits for all ranks' flags in local memory - no collective
            kernels.sharded_step(x_local, y, a_local, b, u, v, scale, reg, min(stage, 1), peer["ptrs"], world, rank,
                                 peer["xchg"], peer["ctrl"], diffs, ws, cost=cost, precision=precision)
            return
        else:
            kernels.colstep(x_local, y, u, scale, reg, cost, precision, out=part, ws=ws, reuse=min(stage, 2) if stage else 0)
            if world > 1:
                dist.all_gather_into_tensor(gathered.view(-1), part.view(-1), group=group)
            diffs.zero_()
            kernels.lse_combine(gathered[:, 0], gathered[:, 1], b, v, diffs[1:2])
        kernels.rowstep(x_local, y, a_local, v, u, diffs[0:1], scale, reg, cost, precision, ws=ws, reuse=max(1, min(stage, 2)))

    def converged() -> bool:
        du = diffs[0:1].clone()
        if world > 1:
            dist.all_reduce(du, group=group)
        return float(du.item() + diffs[1].item()) < threshold

    done_iters = 0
    for it in range(max_iter):
        if want_graph and it == 2 and plan["graph"] is None and not plan["refused"]:
            plan["graph"] = _capture(lambda **kw: iteration(2), dev)
            plan["refused"] = plan["graph"] is None
        if it >= 2 and plan["graph"] is not None:
            plan["graph"].replay()
        else:
            iteration(min(it, 2))
        done_iters = it + 1
        if threshold > 0 and ((it + 1) % poll_every == 0 or it + 1 == max_iter) and converged():
            break
    # the plan's buffers are reused if the caller passes the plan back: hand out copies
    return dict(u_local=u.clone(), v=v.clone(), iters=done_iters, scale=scale, plan=plan)


def _peer_exchange(world: int, m: int, dev: torch.device, group=None) -> Optional[dict]:
    """Symmetric-memory exchange buffer of the fused column-partial exchange (otk_sinkhorn_points_colstep_push /
    otk_lse_combine_wait): allocated with torch's symmetric-memory allocator, mapped into every peer of the node by the
    rendezvous.  None if that is unavailable (other engines, OTK_SINKHORN_PEER=0, no P2P): the NCCL all-gather is used."""
    if os.environ.get("OTK_SINKHORN_PEER", "1") == "0":
        return None
    try:
        import torch.distributed._symmetric_memory as symm
        nbytes = kernels_exchange_bytes(world, m)
        xchg = symm.empty(nbytes, dtype=torch.uint8, device=dev)
        hdl = symm.rendezvous(xchg, dist.group.WORLD if group is None else group)
        xchg.zero_()
        ptrs = torch.tensor([int(p) for p in hdl.buffer_ptrs], dtype=torch.int64, device=dev)
        ctrl = torch.zeros(8, dtype=torch.int32, device=dev)
        torch.cuda.synchronize(dev)
        hdl.barrier()                   # every rank's flags are zero before anyone pushes
        return dict(xchg=xchg, hdl=hdl, ptrs=ptrs, ctrl=ctrl)
    except Exception as e:  # noqa: BLE001 - an optimisation: the collective path is always valid
        import warnings
        warnings.warn(f"sharded_sinkhorn: peer-memory exchange unavailable ({type(e).__name__}: {e}); using all_gather")
        return None


def kernels_exchange_bytes(world: int, m: int) -> int:
    return K.exchange_bytes(world, m)


def sharded_summary(x_local: Tensor, y: Tensor, a_local: Tensor, b: Tensor, u_local: Tensor, v: Tensor, scale: float,
                    reg: float, cost: int = 0, precision: int = 0, group=None, kernels=K) -> dict:
    """<C,pi>, total mass and the marginal errors of the row-sharded plan pi_ij = exp(u_i + v_j - C_ij/reg), which is never
    materialised: every rank evaluates its rows (`otk_sinkhorn_points_summary`), then ONE packed SUM all-reduce of
    [cost, mass | column partial marginals] and one MAX all-reduce of the row error.  Same numbers on every
    rank; at world size 1 they equal `kernels.sinkhorn_points(...)["summary"]`."""
    part, _, col_part = kernels.points_summary(x_local, y, a_local, b, u_local, v, scale, reg, cost, precision)
    packed = torch.cat([part[:2], col_part.double()])
    row_err = part[2:3].clone()
    if _world(group) > 1:
        dist.all_reduce(packed, group=group)
        dist.all_reduce(row_err, op=dist.ReduceOp.MAX, group=group)
    col_err = (packed[2:] - b.double()).abs().max()
    return dict(cost=float(packed[0]), mass=float(packed[1]), max_row_err=float(row_err), max_col_err=float(col_err))


def _capture(iteration, dev):
    """Capture one steady-state iteration (prepared operands reused) into a CUDA graph; None if capture is refused."""
    try:
        torch.cuda.synchronize(dev)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            iteration(first=False)
        return g
    except Exception as e:  # noqa: BLE001 - capture is an optimisation; the eager loop is always valid
        import warnings
        warnings.warn(f"sharded_sinkhorn: CUDA-graph capture unavailable ({type(e).__name__}: {e}); running eagerly")
        torch.cuda.synchronize(dev)
        return None
