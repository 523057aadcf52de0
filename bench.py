#!/usr/bin/env python
"""Benchmark of the latent optimal-transport hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Primary metric  "cov+W2-map latents/s" on BASELINE.json configs[1]:
    one step = stream 2^20 source + 2^20 target synthetic 512-d latents (chunks of 65536) through the statistics
    kernel, compute the Gaussian W2 map (`GaussianTransport.compute`), transport the 2^20 source latents.
    value = latents transported per second, whole job (all ranks; weak scaling: every rank has its own 2^20).
Secondary metric (same JSON line, key "sinkhorn"): log-domain Sinkhorn iterations/s at N=M=65536, d=128, eps=0.05
    (BASELINE.json configs[2]; rows sharded over the ranks, strong scaling).
`e2e` measures the same step through the public Python API with pinned HOST buffers (H2D of both latent sets and
D2H of the transported latents inside the timed region).  `--impl reference` drives the UNMODIFIED reference
(pip-installed into the git-ignored baseline/_ref by __graft_entry__.build(); torch fp64 on all host threads) through the
same step; the oracle port is only the fallback when baseline/_ref is missing.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

D_LAT, N_LAT, CHUNK = 512, 1 << 20, 1 << 16
SK_N, SK_D, SK_EPS = 65536, 128, 0.05
METRIC, UNIT = "cov+W2-map latents/s", "latents/s"
# dram__bytes_read.sum + dram__bytes_write.sum of one stats_h2_kernel<2> launch on a 65536 x 512 chunk
# (ncu --set full, profiles/prof_stats_r06.md): 134.89 MB + 13.29 MB (partial tiles); the algorithmic figure is 134.2 MB
STATS_TRAFFIC_BYTES_PER_LAUNCH = 148.2e6


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm=p["hbm_gbs"], bf16=p["bf16_tflops"], bf16_sustained=p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                    source="measured")
    return dict(hbm=6650.0, bf16=1590.0, bf16_sustained=1400.0, source="fallback")


class ClockSampler(threading.Thread):
    """Samples SM clocks / throttle reasons during the timed regions with NVML: a thread polls every 50 ms and `timed()`
    takes one more sample right after the last step has been enqueued (the GPU is still busy then, and the call sits
    outside the CUDA-event bracket).  The polling is deliberately sparse: at 10 ms per sample plus an `nvidia-smi -lms 50`
    loop per rank the host-driven row-sharded Sinkhorn loop lost half of its speed at 8 GPUs (2064 -> 911 it/s).
    `nvidia-smi -lms 200` (the profiling recipe's loop) is only started when NVML is unusable."""

    REASONS = (("nvmlClocksThrottleReasonHwSlowdown", "hw_slowdown"),
               ("nvmlClocksThrottleReasonHwThermalSlowdown", "hw_thermal_slowdown"),
               ("nvmlClocksThrottleReasonSwThermalSlowdown", "sw_thermal_slowdown"),
               ("nvmlClocksThrottleReasonSwPowerCap", "sw_power_cap"))

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        self.windows, self.errors, self.smi, self.smi_path = [], [], None, None
        try:
            if os.environ.get("OTK_BENCH_NO_CLOCKS") == "1":      # diagnostic: measure the sampler's own perturbation
                raise RuntimeError("disabled by OTK_BENCH_NO_CLOCKS")
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:
            self.nv = None
            self.errors.append(f"nvml init: {type(e).__name__}: {e}")
        try:
            if self.nv is not None or os.environ.get("OTK_BENCH_NO_CLOCKS") == "1":
                raise RuntimeError("nvml available")
            import subprocess
            import tempfile
            fd, self.smi_path = tempfile.mkstemp(prefix="otk_clocks_", suffix=".csv")
            q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                 "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
            self.smi = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                         "-lms", "200"], stdout=fd, stderr=subprocess.DEVNULL)
            os.close(fd)
        except Exception as e:
            if self.nv is None:
                self.errors.append(f"nvidia-smi: {type(e).__name__}: {e}")

    def mark(self, t0, t1):
        """a timed region (host clock) - samples inside the regions are the ones reported"""
        self.windows.append((t0, t1))

    def sample_now(self):
        if self.nv is None:
            return
        nv = self.nv
        try:
            mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
            mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            self.samples.append((time.perf_counter(), mhz, mask))
        except Exception as e:
            if len(self.errors) < 3:
                self.errors.append(f"nvml sample: {type(e).__name__}: {e}")

    def run(self):
        if self.nv is None:
            return
        self.names = [(getattr(self.nv, attr), name) for attr, name in self.REASONS if hasattr(self.nv, attr)]
        while not self.stop_flag:
            self.sample_now()
            time.sleep(0.05)

    def result(self):
        self.stop_flag = True
        if self.is_alive():
            self.join(timeout=1.0)
        names = getattr(self, "names", [])
        inside = [s for s in self.samples if any(a <= s[0] <= b for a, b in self.windows)] or self.samples
        mhz = sorted(s[1] for s in inside)
        reasons = {name for _, _, mask in inside for bit, name in names if mask & bit}
        smi_mhz = []
        if self.smi is not None:
            self.smi.terminate()
            try:
                self.smi.wait(timeout=2.0)
                for ln in open(self.smi_path):
                    f = [t.strip() for t in ln.split(",")]
                    if len(f) >= 6 and f[0].isdigit():
                        smi_mhz.append(int(f[0]))
                        self.max_mhz = self.max_mhz or int(f[1])
                        for flag, name in zip(f[2:6], ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")):
                            if flag == "Active" and not mhz:
                                reasons.add(name)
                os.unlink(self.smi_path)
            except Exception as e:
                self.errors.append(f"nvidia-smi parse: {type(e).__name__}: {e}")
        out = dict(sm_mhz=mhz[len(mhz) // 2] if mhz else None, sm_max_mhz=self.max_mhz, reasons=sorted(reasons),
                   samples=len(mhz), source="nvml: 50 ms polling + one sample per timed region right after its last step "
                                            "was enqueued; samples inside the timed regions" if mhz else None)
        if smi_mhz:
            smi_mhz.sort()
            # nvidia-smi samples cover the whole run (idle gaps included): the upper quartile is the under-load clock
            out["smi_sm_mhz_p75"] = smi_mhz[(3 * len(smi_mhz)) // 4]
            if not mhz:
                out.update(sm_mhz=out["smi_sm_mhz_p75"], samples=len(smi_mhz), source="nvidia-smi -lms 200, upper quartile of the run")
        if self.errors:
            out["sampler_errors"] = self.errors
        return out


# ------------------------------------------------------------------------------------------------ reference arm (CPU)

WORKLOAD = ("cfg2: streaming cov (2^20 source + 2^20 target 512-d latents, chunks of 65536) + Gaussian W2 map + transport of "
            "the 2^20 source latents, per rank")


def shared_config():
    """identical in both arms (the driver compares it): what one step is, nothing about how an arm computes it"""
    return dict(workload=WORKLOAD, dim=D_LAT, latents_per_rank=N_LAT, chunk=CHUNK, l2="inputs (2 x 2 GiB) exceed L2")


def load_reference_package():
    """The UNMODIFIED reference, pip-installed into baseline/_ref by `__graft_entry__.build()` (git-ignored, shipped to the
    GPU box), imported behind the third-party stubs of baseline/ref_loader.py.  None if it is not there."""
    ref_root = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref_root, "ot_vae_lightning")):
        return None
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    import ref_loader
    return ref_loader.load_reference(ref_root)


def ref_gaussian_step(ref, src, tgt, chunk, dim):
    """update(source, target) per chunk -> compute() -> transport(source) per chunk, through the reference's own classes on
    the host cores, with the reference's defaults for the path (fp64: `dtype=torch.double`)."""
    from ot_vae_lightning.ot.transport.gaussian_transport import GaussianTransport as RefGT
    cfg = dict(dtype=torch.double)
    op = RefGT(dim, transport_cfg=dict(diag=False, stochastic=False, make_pd=True, dtype=torch.double),
               source_cfg=dict(cfg), target_cfg=dict(cfg))
    n = src.shape[0]
    t0 = time.perf_counter()
    for lo in range(0, n, chunk):
        op.update(source_samples=src[lo:lo + chunk], target_samples=tgt[lo:lo + chunk])
    w2 = op.compute()
    last = None
    for lo in range(0, n, chunk):
        last = op.transport(src[lo:lo + chunk])
    dt = time.perf_counter() - t0
    assert torch.isfinite(last).all()
    return dt, float(w2)


def port_gaussian_step(src, tgt, chunk):
    from oracle import ot_oracle as O
    t0 = time.perf_counter()
    out = O.gaussian_transport_pipeline(src, tgt, chunk)
    dt = time.perf_counter() - t0
    assert torch.isfinite(out["moved"]).all()
    return dt, float(out["w2"])


def cpu_gaussian(sample_rows, threads, dim=D_LAT, chunk=CHUNK, reps=1):
    """(seconds per pass, kind, w2) of the CPU path on `sample_rows` source + target latents of the bench workload"""
    from ot_vae_lightning_b200.synthetic import gaussian_latents
    torch.set_num_threads(threads)
    src = gaussian_latents(sample_rows, dim, seed=1234, sample_seed=9001)
    tgt = gaussian_latents(sample_rows, dim, seed=4321, shift=0.5, scale=1.5, sample_seed=7001)
    ref = load_reference_package()
    times = []
    for _ in range(reps):
        if ref is not None:
            dt, w2 = ref_gaussian_step(ref, src, tgt, min(chunk, sample_rows), dim)
        else:
            dt, w2 = port_gaussian_step(src, tgt, min(chunk, sample_rows))
        times.append(dt)
    return sum(times) / len(times), ("reference" if ref is not None else "port"), w2


def cpu_sinkhorn(n, iters, threads):
    from oracle import ot_oracle as O
    from ot_vae_lightning_b200.synthetic import point_clouds
    torch.set_num_threads(threads)
    x, y = point_clouds(n, n, SK_D, seed=1234)
    C = O.sqeuclidean_cost(x, y)
    C = C / C.max()
    a = torch.full((n,), 1.0 / n)
    ref = load_reference_package()
    t0 = time.perf_counter()
    if ref is not None:
        from ot_vae_lightning.ot.w2_utils import sinkhorn_log as ref_sinkhorn
        ref_sinkhorn(a, a, C, reg=SK_EPS, max_iter=iters, threshold=0.0)
    else:
        O.sinkhorn_log(a, a, C, reg=SK_EPS, max_iter=iters, threshold=0.0)
    return (time.perf_counter() - t0) / iters


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    # the SAME workload as the B200 arm: 2^20 source + 2^20 target 512-d latents per step.  One step takes ~30-60 s on the
    # host cores, so at most two steps are timed (after one small warm-up pass): `steps` is what was timed.
    timed = max(1, min(args.steps, 2))
    if args.warmup > 0:
        cpu_gaussian(1 << 14, threads)
    t, kind, w2 = cpu_gaussian(N_LAT, threads, reps=timed)
    val = N_LAT / t
    sk_n = 4096
    sk_t = cpu_sinkhorn(sk_n, 3, threads)
    how = ("unmodified reference classes from baseline/_ref (GaussianTransport.update / compute / transport; fp64 einsum "
           "SYRK, eigh-based sqrtm, fp64 broadcast mat-vecs)" if kind == "reference" else
           "oracle port of the reference (baseline/_ref not installed)")
    line = dict(impl="reference", metric=METRIC, value=val, unit=UNIT, n_gpus=args.gpus, steps=timed,
                steps_requested=args.steps, warmup=min(args.warmup, 1), ms_per_step=t * 1e3, higher_is_better=True,
                scaling="weak", vs_baseline=None, dtype="f64", data="synthetic", config=shared_config(),
                cpu_baseline=dict(value=val, unit=UNIT, cores=threads, kind=kind,
                                  sample=f"the full step: {N_LAT} source + {N_LAT} target latents, d={D_LAT}, chunks of "
                                         f"{CHUNK}; {timed} timed step(s); {how}"),
                e2e=dict(value=val, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                check=dict(w2=w2),
                sinkhorn=dict(metric="Sinkhorn iters/s", value=1.0 / (sk_t * (SK_N / sk_n) ** 2), unit="iters/s",
                              measured_at=f"N=M={sk_n} fp32 ({1.0 / sk_t:.3f} it/s), EXTRAPOLATED to 65536^2 by N*M (the "
                                          f"reference materialises ~3 N x M buffers: 51 GB fp32 at 65536^2)",
                              cores=threads, kind=kind))
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ B200 arm

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--sinkhorn-iters", type=int, default=100)   # threshold=0, max_iter=100 for timing
    ap.add_argument("--skip-sinkhorn", action="store_true")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-e2e", action="store_true")
    ap.add_argument("--skip-sweeps", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    # stdout carries exactly one JSON line: anything a library prints there (NCCL's version banner) goes to stderr
    json_fd = os.dup(1)
    os.dup2(2, 1)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback exists)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from ot_vae_lightning_b200 import _native as NV
    from ot_vae_lightning_b200 import kernels as K
    from ot_vae_lightning_b200 import parallel
    from ot_vae_lightning_b200.ot import GaussianTransport
    from ot_vae_lightning_b200.synthetic import gaussian_latents, point_clouds
    lib = NV.load()
    peaks = load_peaks()

    def probe_peak(kind):
        import ctypes
        out = ctypes.c_double(0.0)
        st = lib.otk_microbench_peak(kind, ctypes.byref(out), ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
        return float(out.value) if st == 0 else None

    # denominators MEASURED_PEAKS.json does not hold ("builder must measure"), measured on this GPU now
    peaks_measured = dict(tf32_tflops=probe_peak(0), f16_tflops=probe_peak(1), mufu_ex2_tera_per_s=probe_peak(2),
                          dfma_tflops=probe_peak(3),
                          how="otk_microbench_peak: tcgen05 kind::tf32 / kind::f16 with CTA pairs (M = N = 256) on operands "
                              "resident in shared / tensor memory (issue-bound ceiling of the MMA pipe, no loads); "
                              "ex2.approx.ftz.f32, 8 independent chains per thread, 8 CTAs x 256 threads per SM; best of 3 "
                              "launches, CUDA events")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- workload: every rank owns its own 2^20 source / target latents (weak scaling)
    # of the SAME two distributions (the ranks differ in the draws, as the shards of one validation set would)
    src = gaussian_latents(N_LAT, D_LAT, seed=1234, device=dev, sample_seed=9001 + rank)
    tgt = gaussian_latents(N_LAT, D_LAT, seed=4321, device=dev, shift=0.5, scale=1.5, sample_seed=7001 + rank)
    n_chunks_all = N_LAT // CHUNK
    outs = [None] * n_chunks_all       # transport() returns a fresh tensor per chunk (reference semantics); all are kept
    cfg = dict(dtype=torch.double, device=dev, reduce_on_update=False)
    op = GaussianTransport(D_LAT, transport_cfg=dict(make_pd=True), source_cfg=dict(cfg), target_cfg=dict(cfg)).to(dev)

    def step_device():
        op.reset()
        for lo in range(0, N_LAT, CHUNK):
            op.update(source_samples=src[lo:lo + CHUNK], target_samples=tgt[lo:lo + CHUNK])
        w2 = op.compute()              # one packed all-reduce of the statistics inside fit() when world > 1
        for i, lo in enumerate(range(0, N_LAT, CHUNK)):
            outs[i] = op.transport(src[lo:lo + CHUNK])
        return w2

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(steps):
            res = fn()
        e1.record()
        if sampler is not None:
            sampler.sample_now()              # the queue is still draining: an under-load sample outside the event bracket
        barrier()
        if sampler is not None:
            sampler.mark(t0, time.perf_counter())
        return max_over_ranks(e0.elapsed_time(e1)), res

    sampler = None
    for _ in range(max(args.warmup, 3)):
        step_device()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = lib.otk_launch_count()
    profiling = os.environ.get("OTK_PROFILE_TIMED") == "1"     # ncu --profile-from-start off: only the timed region
    if profiling:
        torch.cuda.cudart().cudaProfilerStart()
    total_ms, w2 = timed(step_device, args.steps)
    if profiling:
        torch.cuda.cudart().cudaProfilerStop()
    launches = lib.otk_launch_count() - launches0
    ms_per_step = total_ms / args.steps
    value = world * N_LAT / (ms_per_step * 1e-3)

    # ---- per-kernel roofline: the statistics kernel over one rank's 2^20 x 512 latents
    def stats_only():
        op.source_model.reset()
        for lo in range(0, N_LAT, CHUNK):
            op.source_model.update(src[lo:lo + CHUNK])

    def apply_only():
        for i, lo in enumerate(range(0, N_LAT, CHUNK)):
            outs[i] = op.transport(src[lo:lo + CHUNK])

    stats_only()
    stats_ms, _ = timed(stats_only, 3)
    apply_ms, _ = timed(apply_only, 3)
    compute_ms, _ = timed(lambda: op.compute(), 3)
    clocks = sampler.result()
    flops = 2.0 * N_LAT * D_LAT * D_LAT
    tf32_peak = peaks["bf16_sustained"] / 2.0
    stats_tflops = flops / (stats_ms / 3 * 1e-3) / 1e12
    apply_tflops = flops / (apply_ms / 3 * 1e-3) / 1e12
    n_chunks = N_LAT // CHUNK
    # executed MMA flops of the statistics kernel: 3 FP16 MMAs per product (hi/lo split) on the 256x256 blocks (CTA-pair
    # kernel) of the upper block triangle (3 of 4 at d = 512); the kernel's tensor peak is the 16-bit dense one
    nb = -(-D_LAT // 256)
    tri = (nb * (nb + 1) / 2) / (nb * nb)
    f16_peak = peaks["bf16_sustained"]
    roofline = dict(kernel="stats_h2_kernel<2> (K1: sum x x^T, sum x, n; CTA pairs, cta_group::2), one launch per 65536 x 512 "
                           "fp32 chunk",
                    bound="tensor", achieved=stats_tflops, peak=f16_peak, unit="TFLOP/s", frac=stats_tflops / f16_peak,
                    traffic=STATS_TRAFFIC_BYTES_PER_LAUNCH,
                    algorithmic=dict(flops_per_launch=2.0 * CHUNK * D_LAT * D_LAT, bytes_per_launch=CHUNK * D_LAT * 4,
                                     launches_per_step=2 * n_chunks, avg_launch_ms=stats_ms / 3 / n_chunks),
                    executed=dict(tflops=3.0 * tri * stats_tflops, frac=3.0 * tri * stats_tflops / f16_peak,
                                  note="3 FP16 MMAs per fp32-accurate product, upper block triangle only: the ceiling of "
                                       f"`frac` for this scheme is 1/(3*{tri:.3f}) = {1 / (3 * tri):.2f}"),
                    note=f"achieved = algorithmic flops 2*N*d^2 / CUDA-event time of the update calls (kernel 67 us under ncu + "
                         f"its helper kernels: pivot/scale 4, partial-tile merge 14-17, gated fallback + clear 6 us, chained by "
                         f"programmatic dependent launch); peak = measured "
                         f"16-bit dense bf16_tflops_sustained ({peaks['source']}) - the kernel issues kind::f16 MMAs; ncu "
                         f"(profiles/prof_stats_r06.md): FP16 tensor ops 53 % of the nominal peak at the 1.77 GHz the "
                         f"kernel ran at, tensor pipe 59 % active; traffic = dram "
                         f"read+write bytes per launch from the same capture (algorithmic: {CHUNK * D_LAT * 4})",
                    vs_measured_f16_peak=dict(peak=peaks_measured.get("f16_tflops"),
                                              executed_frac=(3.0 * tri * stats_tflops / peaks_measured["f16_tflops"]) if peaks_measured.get("f16_tflops") else None,
                                              note="executed FP16 MMA flops against the tcgen05 kind::f16 ceiling measured by "
                                                   "otk_microbench_peak on this GPU"),
                    others=dict(apply_transport_tflops=apply_tflops, apply_frac=apply_tflops / f16_peak,
                                apply_executed_frac=3.0 * apply_tflops / f16_peak,
                                compute_map_ms=compute_ms / 3, stats_ms=stats_ms / 3, apply_ms=apply_ms / 3,
                                stats_gbs=N_LAT * D_LAT * 4 / (stats_ms / 3 * 1e-3) / 1e9,
                                apply_gbs=2 * N_LAT * D_LAT * 4 / (apply_ms / 3 * 1e-3) / 1e9, hbm_peak_gbs=peaks["hbm"]))

    # ---- d = 128 (the latent width north_star's target is quoted on): the two streaming kernels over 2^20 latents
    d128 = None
    try:
        x128 = gaussian_latents(N_LAT, 128, seed=77, device=dev)
        op128 = GaussianTransport(128, transport_cfg=dict(make_pd=True), source_cfg=dict(cfg), target_cfg=dict(cfg)).to(dev)
        op128.update(source_samples=x128, target_samples=x128 * 1.5 + 0.5)
        op128.compute()

        def stats128():
            op128.source_model.update(x128)       # accumulating calls: the timed region holds update() only

        def apply128():
            op128.transport(x128)

        apply128()
        a_ms, _ = timed(apply128, 5)        # before stats128, which resets the source model the map was prepared from
        op128.source_model.reset()
        stats128()
        s_ms, _ = timed(stats128, 5)
        s_ms, a_ms = s_ms / 5, a_ms / 5
        f128 = 2.0 * N_LAT * 128 * 128
        d128 = dict(latents=N_LAT, dim=128,
                    stats=dict(ms=s_ms, gbs=N_LAT * 128 * 4 / (s_ms * 1e-3) / 1e9, hbm_frac=N_LAT * 128 * 4 / (s_ms * 1e-3) / 1e9 / peaks["hbm"],
                               executed_tflops=3 * f128 / (s_ms * 1e-3) / 1e12,
                               executed_tensor_frac=3 * f128 / (s_ms * 1e-3) / 1e12 / peaks["bf16_sustained"],
                               note="one update() call (pivot/scale + FP16 hi/lo split kernel stats_h_kernel + gated fallback + "
                                    "record merge); HBM-bound: one read of X needs 82 us, the kernel alone takes 105 us "
                                    "(ncu, profiles/prof_stats_r04.md: 536.9 MB DRAM read = the algorithmic bytes); executed "
                                    "flops are FP16 MMAs (3 per product) against the measured 16-bit dense peak"),
                    apply=dict(ms=a_ms, gbs=2 * N_LAT * 128 * 4 / (a_ms * 1e-3) / 1e9,
                               hbm_frac=2 * N_LAT * 128 * 4 / (a_ms * 1e-3) / 1e9 / peaks["hbm"],
                               note="one transport() call; HBM-bound (read X, write Y)"))
        del x128, op128
    except Exception as e:  # secondary numbers must never take the primary line down
        d128 = dict(error=f"{type(e).__name__}: {e}")

    def guarded(fn):
        try:
            return fn()
        except Exception as e:  # secondary numbers must never take the primary line down
            torch.cuda.synchronize()
            return dict(error=f"{type(e).__name__}: {e}"[:300])

    def fast_counters():
        import ctypes
        cnt = (ctypes.c_int * 4)()
        lib.otkdbg_fast_counters(cnt)
        return list(cnt)

    def pipeline_parts(op_, xs, xt, batch, reps):
        """ms per step (reset, update both models per batch, compute, transport every source batch) and its three parts"""
        n_ = xs.shape[-2]
        keep = [None] * (-(-n_ // batch))

        def upd():
            op_.reset()
            for lo in range(0, n_, batch):
                op_.update(source_samples=xs[..., lo:lo + batch, :], target_samples=xt[..., lo:lo + batch, :])

        def mov():
            for i, lo in enumerate(range(0, n_, batch)):
                keep[i] = op_.transport(xs[..., lo:lo + batch, :])

        def step():
            upd()
            op_.compute()
            mov()

        step()
        step()
        ms, _ = timed(step, reps)
        u_ms, _ = timed(upd, reps)
        c_reps = max(reps, 10)
        f0 = fast_counters()
        c_ms, _ = timed(lambda: op_.compute(), c_reps)
        f1 = fast_counters()
        t_ms, _ = timed(mov, reps)
        # how the timed compute() calls ran: accepted / rejected by the single-graph operator path, graph replays / eager
        path = dict(zip(("accepted", "rejected", "graph_replays", "eager"), (b - a for a, b in zip(f0, f1))))
        return dict(ms_per_step=ms / reps, update_ms=u_ms / reps, compute_ms=c_ms / c_reps, transport_ms=t_ms / reps,
                    compute_path=path)

    # ---- cfg1: the reference's own operating point (README.md:54-57, tests/test_latent_transport.py:66-98):
    # 10 000 latents of width 128 in batches of 250 -> 40 update calls per model, compute(), 40 transport calls
    def bench_cfg1():
        n1, d1, b1 = 10000, 128, 250
        xs = gaussian_latents(n1, d1, seed=11, device=dev, sample_seed=100 + rank)
        xt = gaussian_latents(n1, d1, seed=12, device=dev, shift=0.5, scale=1.5, sample_seed=200 + rank)
        op1 = GaussianTransport(d1, transport_cfg=dict(make_pd=True), source_cfg=dict(cfg), target_cfg=dict(cfg)).to(dev)
        l0 = lib.otk_launch_count()
        out = pipeline_parts(op1, xs, xt, b1, 10)
        out.update(workload="cfg1: N=10000, d=128, batches of 250: update(source, target) x40 -> compute -> transport x40",
                   latents_per_s=world * n1 / (out["ms_per_step"] * 1e-3), per_update_call_us=out["update_ms"] * 1e3 / 40,
                   per_transport_call_us=out["transport_ms"] * 1e3 / 40, gpu_launches_per_step=int((lib.otk_launch_count() - l0) / 42),
                   note="latency-bound: 120 small calls + one map per step; per-call figures include the Python mirror")
        return out

    # ---- cfg4: conditional transport, one operator per class: GaussianTransport(10, 1024) (transport_callback.py:388-453)
    def bench_cfg4():
        L4, d4, n4, b4 = 10, 1024, 4096, 1024
        xs = torch.stack([gaussian_latents(n4, d4, seed=120 + k, device=dev) for k in range(L4)])
        xt = torch.stack([gaussian_latents(n4, d4, seed=140 + k, device=dev, shift=1.0, scale=0.8) for k in range(L4)])
        op4 = GaussianTransport(L4, d4, transport_cfg=dict(make_pd=True), source_cfg=dict(cfg), target_cfg=dict(cfg)).to(dev)
        out = pipeline_parts(op4, xs, xt, b4, 3)
        flop_map = L4 * (12 * 14 + 8) * float(d4) ** 3
        out.update(workload="cfg4: 10 classes x d=1024, 4096 latents per class (4 d), batches of 1024 per class",
                   latents_per_s=world * L4 * n4 / (out["ms_per_step"] * 1e-3),
                   map_tflops_algorithmic=flop_map / (out["compute_ms"] * 1e-3) / 1e12,
                   note="sample covariances of 4 d observations (condition number ~ 9 x the population's 1e2); "
                        "map flops counted as (12 K + 8) d^3 with K = 14 per operator (SURVEY 8d)")
        return out

    # ---- cfg5: sweeps d = 64 .. 4096 (statistics / map / transport) and Sinkhorn N = 4k .. 256k
    def bench_cfg5():
        rows_out = []
        for d5 in (64, 128, 256, 512, 1024, 2048, 4096):
            def one(d5=d5):
                n5 = 65536 if d5 <= 1024 else 16384
                xs = gaussian_latents(n5, d5, seed=300 + d5, device=dev)
                xt = (xs.roll(1, dims=1) * 1.5 + 0.5).contiguous()     # same spectrum, other eigenvectors: one QR per width
                op5 = GaussianTransport(d5, transport_cfg=dict(make_pd=True), source_cfg=dict(cfg), target_cfg=dict(cfg)).to(dev)
                r = pipeline_parts(op5, xs, xt, n5, 3)
                fl = 2.0 * n5 * d5 * d5
                return dict(d=d5, rows=n5, update_ms=r["update_ms"] / 2, compute_ms=r["compute_ms"], transport_ms=r["transport_ms"],
                            compute_path=r["compute_path"],
                            update_tflops=fl / (r["update_ms"] / 2 * 1e-3) / 1e12, update_gbs=n5 * d5 * 4 / (r["update_ms"] / 2 * 1e-3) / 1e9,
                            transport_tflops=fl / (r["transport_ms"] * 1e-3) / 1e12,
                            transport_gbs=2 * n5 * d5 * 4 / (r["transport_ms"] * 1e-3) / 1e9)
            res = guarded(one)
            res.setdefault("d", d5)
            rows_out.append(res)
            torch.cuda.empty_cache()
        sk_out = []
        for n5 in (4096, 16384, 65536, 262144):
            def one(n5=n5):
                x5, y5 = point_clouds(n5, n5, SK_D, seed=500 + (n5 >> 10), device=dev)
                a5 = torch.full((n5,), 1.0 / n5, device=dev)
                lo5, hi5 = parallel.shard_rows(n5, rank, world)
                it5 = 20
                if world == 1:
                    sc5 = 1.0 / float(K.cost_max(x5, y5, 0).item())
                    run5 = lambda: K.sinkhorn_points(x5, y5, a5, a5, reg=SK_EPS, max_iter=it5, threshold=0.0, scale=sc5,
                                                     want_summary=False, want_iters=False)
                else:
                    xl5, al5 = x5[lo5:hi5].contiguous(), a5[lo5:hi5].contiguous()
                    sc5 = parallel.global_cost_scale(xl5, y5)
                    st5 = {}

                    def run5():
                        # the plan (exchange buffer in symmetric memory, workspace) is set up once and handed back
                        out5 = parallel.sharded_sinkhorn(xl5, y5, al5, a5, reg=SK_EPS, max_iter=it5, threshold=0.0,
                                                         scale=sc5, use_graph=False, plan=st5.get("plan"))
                        st5["plan"] = out5["plan"]
                run5()
                ms5, _ = timed(run5, 1)
                return dict(N=n5, iters_per_s=it5 / (ms5 * 1e-3), ms_per_iter=ms5 / it5)
            res = guarded(one)
            res.setdefault("N", n5)
            sk_out.append(res)
            torch.cuda.empty_cache()
        return dict(workload="cfg5: one 65536-row chunk (16384 rows for d >= 2048) per width through update / compute / "
                             "transport; fused Sinkhorn at d=128, eps=0.05, 20 iterations per size, rows sharded over the ranks",
                    cov_map=rows_out, sinkhorn=sk_out)

    cfg1 = guarded(bench_cfg1)
    cfg4 = guarded(bench_cfg4)
    cfg5 = guarded(bench_cfg5) if not args.skip_sweeps else None
    torch.cuda.empty_cache()

    # ---- e2e: same step through the public API with pinned HOST buffers
    e2e = None
    if not args.skip_e2e:
        h_src, h_tgt = src.cpu().pin_memory(), tgt.cpu().pin_memory()
        h_out = torch.empty_like(h_src).pin_memory()

        from ot_vae_lightning_b200.streaming import stream_transport, stream_update

        def step_host():
            op.reset()
            stream_update(op, h_src, h_tgt, CHUNK, dev)        # H2D of chunk i+1 overlaps the kernels of chunk i
            op.compute()
            stream_transport(op, h_src, h_out, CHUNK, dev)      # H2D | kernels | D2H on three streams
            torch.cuda.synchronize()
            return float(h_out[0, 0])

        step_host()
        e2e_ms, _ = timed(step_host, max(1, min(args.steps, 3)))
        e2e_ms /= max(1, min(args.steps, 3))
        # the ceiling of this step: the same bytes in the same chunks over the same three streams, no kernels at all
        # (update phase: H2D of source + target chunks; transport phase: H2D of source chunks while the previous result
        # chunk goes D2H) - what the host memory system / PCIe delivers to this rank while all ranks copy at once
        def copies_only():
            h2d, d2h = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
            ring = [torch.empty(CHUNK, D_LAT, device=dev) for _ in range(4)]
            with torch.cuda.stream(h2d):
                for i, lo_ in enumerate(range(0, N_LAT, CHUNK)):
                    ring[(2 * i) % 4].copy_(h_src[lo_:lo_ + CHUNK], non_blocking=True)
                    ring[(2 * i + 1) % 4].copy_(h_tgt[lo_:lo_ + CHUNK], non_blocking=True)
            h2d.synchronize()
            for i, lo_ in enumerate(range(0, N_LAT, CHUNK)):
                with torch.cuda.stream(h2d):
                    ring[i % 2].copy_(h_src[lo_:lo_ + CHUNK], non_blocking=True)
                with torch.cuda.stream(d2h):
                    h_out[lo_:lo_ + CHUNK].copy_(ring[2 + i % 2], non_blocking=True)
            h2d.synchronize()
            d2h.synchronize()

        copies_only()
        barrier()
        t_c0 = time.perf_counter()
        copies_only()
        torch.cuda.synchronize()
        copy_ms = max_over_ranks((time.perf_counter() - t_c0) * 1e3)
        e2e = dict(value=world * N_LAT / (e2e_ms * 1e-3), unit=UNIT, h2d_bytes_per_step=3 * N_LAT * D_LAT * 4,
                   d2h_bytes_per_step=N_LAT * D_LAT * 4, ms_per_step=e2e_ms,
                   host_device_gbs_aggregate=world * 4 * N_LAT * D_LAT * 4 / (e2e_ms * 1e-3) / 1e9,
                   copies_only_ms_per_step=copy_ms, frac_of_copy_ceiling=copy_ms / e2e_ms,
                   note="pinned host latents -> update(src,tgt) -> compute -> transport(src) -> pinned host result; copies "
                        "double-buffered on side streams (streaming.py); PCIe-bound: `copies_only_ms_per_step` is the same "
                        "bytes in the same chunks with no kernel launched (host wall clock, max over ranks, all ranks "
                        "copying at once)")
        del h_src, h_tgt, h_out

    # ---- Sinkhorn secondary metric: N=M=65536, d=128, eps=0.05, rows sharded over the ranks
    sinkhorn = None
    sk_state = {}
    if not args.skip_sinkhorn:
        try:
            del src, tgt
            outs.clear()
            torch.cuda.empty_cache()
            x, y = point_clouds(SK_N, SK_N, SK_D, seed=99, device=dev)
            lo, hi = parallel.shard_rows(SK_N, rank, world)
            a = torch.full((SK_N,), 1.0 / SK_N, device=dev)
            iters = args.sinkhorn_iters
            if world == 1:
                scale = 1.0 / float(K.cost_max(x, y, 0).item())
                run = lambda: K.sinkhorn_points(x, y, a, a, reg=SK_EPS, max_iter=iters, threshold=0.0, scale=scale,
                                                want_summary=False, want_iters=False)
            else:
                scale = parallel.global_cost_scale(x[lo:hi], y)
                xl, al = x[lo:hi].contiguous(), a[lo:hi].contiguous()
                def run():
                    # the captured iteration graph (plan) is reused across calls; it is dropped below, before the
                    # process group goes away
                    out = parallel.sharded_sinkhorn(xl, y, al, a, reg=SK_EPS, max_iter=iters, threshold=0.0, scale=scale,
                                                    plan=sk_state.get("plan"))
                    sk_state["plan"] = out["plan"]
                    return out
            run()
            sampler = ClockSampler(local)          # the Sinkhorn kernel runs at the board power cap: its own clock line
            sampler.start()
            l0 = lib.otk_launch_count()
            sk_ms, sk_res = timed(run, 1)
            del sk_res
            sk_launches = lib.otk_launch_count() - l0
            sk_clocks = sampler.result()
            it_s = iters / (sk_ms * 1e-3)
            alg_gb = 2.0 * SK_N * SK_N * 4 / 1e9
            mufu_peak = peaks_measured.get("mufu_ex2_tera_per_s") or (148 * 16 * (clocks.get("sm_max_mhz") or 1965) * 1e6 / 1e12)
            mufu_src = "measured (otk_microbench_peak)" if peaks_measured.get("mufu_ex2_tera_per_s") else "148 SMs x 16/clk x max clock (estimate)"
            # correctness of what was timed: plan statistics WITHOUT the plan, on every rank count (the sharded rows'
            # <C,pi>, mass and column partials are summed with one all-reduce; same problem, so the numbers must agree
            # across n_gpus to the solver's accuracy)
            if world == 1:
                res = K.sinkhorn_points(x, y, a, a, reg=SK_EPS, max_iter=iters, threshold=0.0, scale=scale)
                s4 = res["summary"].cpu().tolist()
                check = dict(cost=s4[0], mass=s4[1], max_row_err=s4[2], max_col_err=s4[3])
                del res
            else:
                out = run()
                check = parallel.sharded_summary(xl, y, al, a, out["u_local"], out["v"], scale, SK_EPS)
                del out
            sinkhorn = dict(metric="Sinkhorn iters/s", value=it_s, unit="iters/s", ms_per_iter=sk_ms / iters, n_gpus=world,
                            scaling="strong", config=dict(N=SK_N, M=SK_N, d=SK_D, eps=SK_EPS, iters=iters, threshold=0.0,
                                                          cost="sqeuclidean / max", scale=scale),
                            gpu_launches=int(sk_launches), check=check,
                            roofline=dict(bound="hbm", achieved=alg_gb * it_s, peak=peaks["hbm"], unit="GB/s",
                                          frac=alg_gb * it_s / peaks["hbm"], traffic=34.1e6,
                                          note="algorithmic bytes 2*N*M*4 per iteration (one fp32 cost read per half-step: "
                                               "what a streamed cost matrix would move); the fused kernel never "
                                               "materialises the cost, so frac > 1; measured dram traffic per pass "
                                               "launch: 34 MB (profiles/prof_sinkhorn_r02.md)",
                                          binding=dict(bound="mufu", achieved=2.0 * SK_N * SK_N * it_s / 1e12,
                                                       frac=2.0 * SK_N * SK_N * it_s / 1e12 / (world * mufu_peak),
                                                       peak=world * mufu_peak, unit="T ex2/s", peak_source=mufu_src,
                                                       note="2*N*M exponentials per iteration against the MUFU.EX2 rate "
                                                            "measured on this GPU (peaks_measured); ncu: XU pipe 79 % active"),
                                          tensor=dict(achieved=4.0 * SK_N * SK_N * SK_D * it_s / 1e12 / world,
                                                      peak=peaks["bf16_sustained"], unit="TFLOP/s per GPU",
                                                      note="executed FP16 MMA flops 2 passes x 2*N*M*d")))
            sinkhorn["clocks"] = sk_clocks
            # the reference signature (`sinkhorn_log` on a MATERIALISED cost, w2_utils.py:276-319) at the same size: the
            # streaming kernels read the 17 GB fp32 cost once per half-step
            if world == 1:
                def dense_leg():
                    Cd = K.cost_matrix(x, y, 0, scale)
                    d_iters = 10
                    rund = lambda: K.sinkhorn_dense(a, a, Cd, SK_EPS, d_iters, 0.0, want_plan=False)
                    rund()
                    l0d = lib.otk_launch_count()
                    d_ms, outd = timed(rund, 1)
                    ud, vd = outd[1], outd[2]
                    ref10 = K.sinkhorn_points(x, y, a, a, reg=SK_EPS, max_iter=d_iters, threshold=0.0, scale=scale,
                                              want_summary=False, want_iters=False)
                    gbs = alg_gb * d_iters / (d_ms * 1e-3)
                    return dict(iters_per_s=d_iters / (d_ms * 1e-3), ms_per_iter=d_ms / d_iters, iters=d_iters,
                                gpu_launches=int(lib.otk_launch_count() - l0d),
                                roofline=dict(bound="hbm", achieved=gbs, peak=peaks["hbm"], unit="GB/s", frac=gbs / peaks["hbm"],
                                              traffic=None,
                                              note="algorithmic bytes 2*N*M*4 per iteration / CUDA-event time of the loop"),
                                max_abs_u_vs_fused=float((ud - ref10["u"]).abs().max()),
                                max_abs_v_vs_fused=float((vd - ref10["v"]).abs().max()),
                                note="sk_col_partial_kernel + sk_row_kernel (sinkhorn_dense.cu): 16-byte loads, grouped "
                                     "branch-free online LSE, persistent grids; potentials compared with the fused engine")
                sinkhorn["dense"] = guarded(dense_leg)
        except Exception as e:  # keep the primary line even if the secondary workload fails
            sinkhorn = dict(error=f"{type(e).__name__}: {e}")
        finally:                # NCCL cannot tear the communicator down while a graph that captured it is alive
            sk_state.clear()
            import gc
            gc.collect()
            torch.cuda.synchronize()

    # ---- CPU baseline beside it (rank 0, N=1 only): the oracle port on a bounded sample
    cpu = None
    if rank == 0 and world == 1 and not args.skip_cpu:
        threads = os.cpu_count() or 1
        sample = 1 << 18
        cpu_gaussian(1 << 13, threads)
        t, kind, _ = cpu_gaussian(sample, threads)
        cpu = dict(value=sample / t, unit=UNIT, cores=threads, kind=kind,
                   sample=f"{sample} source + {sample} target latents of the same workload (a quarter of a step), d={D_LAT}, "
                          f"chunks of {CHUNK}, fp64, one pass ({t:.1f} s); `--impl reference` times the full step")
        if isinstance(cfg1, dict) and "error" not in cfg1:
            t1, kind1, _ = cpu_gaussian(10000, threads, dim=128, chunk=250)
            cfg1["cpu_reference"] = dict(ms_per_step=t1 * 1e3, latents_per_s=10000 / t1, cores=threads, kind=kind1)

    if rank == 0:
        line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=max(args.warmup, 3),
                    ms_per_step=ms_per_step, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32",
                    data="synthetic",
                    config=shared_config(),
                    arithmetic=dict(gaussian="fp32-accurate products (FP16 hi/lo split with exact scales, TF32 split for the "
                                             "d x d matrix functions), fp64 running statistics",
                                    sinkhorn="FP16 operand planes (TF32-size mantissa), fp32 accumulation and softmax"),
                    check=dict(w2=float(w2)),
                    roofline=roofline, cpu_baseline=cpu, e2e=e2e, gpu_launches=int(launches), clocks=clocks,
                    peaks_measured=peaks_measured, d128=d128, cfg1=cfg1, cfg4=cfg4, cfg5=cfg5, sinkhorn=sinkhorn)
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
