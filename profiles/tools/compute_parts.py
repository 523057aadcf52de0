"""Where the time of `GaussianTransport.compute()` goes: fit_models(), the two symmetrize launches, the operator call."""
import ctypes
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch  # noqa: E402
from ot_vae_lightning_b200 import _native as N  # noqa: E402
from ot_vae_lightning_b200 import kernels as K  # noqa: E402
from ot_vae_lightning_b200.ot import GaussianTransport  # noqa: E402
from ot_vae_lightning_b200.synthetic import gaussian_latents  # noqa: E402

d = int(sys.argv[1]) if len(sys.argv) > 1 else 512
dev = torch.device("cuda", 0)
cfg = dict(dtype=torch.double, device=dev, reduce_on_update=False)
op = GaussianTransport(d, transport_cfg=dict(make_pd=True), source_cfg=dict(cfg), target_cfg=dict(cfg)).to(dev)
src = gaussian_latents(65536, d, seed=1234, device=dev)
tgt = gaussian_latents(65536, d, seed=4321, device=dev, shift=0.5, scale=1.5)
op.update(source_samples=src, target_samples=tgt)


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    host_enq = (time.perf_counter() - t0) / reps * 1e3
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, host_enq


op.compute()
buf = op._map_buffers
sm, tm = op.source_model, op.target_model
raw_s, raw_t = sm.parametrizations.cov.original, tm.parametrizations.cov.original
print(f"d={d}")
print("  compute()            %.3f ms (events)  %.3f ms (host enqueue)" % timed(op.compute))
print("  fit_models()         %.3f ms (events)  %.3f ms (host enqueue)" % timed(op.fit_models))
print("  2 x symmetrize_shift %.3f ms (events)  %.3f ms (host enqueue)" % timed(lambda: (K.symmetrize_shift(raw_s, buf["eps"], out=buf["cov_s"]), K.symmetrize_shift(raw_t, buf["eps"], out=buf["cov_t"]))))
oper = lambda: K.transport_operator(buf["cov_s"], buf["cov_t"], pg_star=0.0, mean_s=sm.mean, mean_t=tm.mean, out=(buf["T"], buf["w2"]))
print("  transport_operator   %.3f ms (events)  %.3f ms (host enqueue)" % timed(oper))
cnt = (ctypes.c_int * 4)()
N.load().otkdbg_fast_counters(cnt)
print("  fast path: accepted %d rejected %d graph replays %d eager %d" % tuple(cnt))
os.environ["OTK_OPERATOR_FAST"] = "0"
