import sys, os, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, pynvml
from ot_vae_lightning_b200 import kernels as K, _native as N
from ot_vae_lightning_b200.synthetic import point_clouds
lib = N.load()
pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
n = 65536
dev = torch.device('cuda', 0)
x, y = point_clouds(n, n, 128, seed=99, device=dev)
a = torch.full((n,), 1.0 / n, device=dev)
scale = 1.0 / float(K.cost_max(x, y, 0).item())
samples = []
stop = False
def sampler():
    while not stop:
        samples.append((time.time(), pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0,
                        pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h), pynvml.nvmlDeviceGetTemperature(h, 0)))
        time.sleep(0.02)
for fast in [0, 1]:
    lib.otkdbg_set_sinkhorn_fast(fast)
    run = lambda it: K.sinkhorn_points(x, y, a, a, reg=0.05, max_iter=it, threshold=0.0, scale=scale, want_summary=False, want_iters=False)
    run(5); torch.cuda.synchronize()
    samples.clear(); stop = False
    th = threading.Thread(target=sampler); th.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(300); e1.record(); torch.cuda.synchronize()
    stop = True; th.join()
    ms = e0.elapsed_time(e1) / 300
    clk = sorted(s[1] for s in samples); pw = sorted(s[2] for s in samples)
    reasons = 0
    for s in samples: reasons |= s[3]
    print(f"fast={fast}: {ms:.3f} ms/iter; SM clock median {clk[len(clk)//2]} min {clk[0]} max {clk[-1]} MHz; power median {pw[len(pw)//2]:.0f} max {pw[-1]:.0f} W; throttle mask {reasons:#x}; temp {samples[-1][4]} C; n={len(samples)}", flush=True)
    print("   clock trace:", [s[1] for s in samples[::max(1, len(samples)//20)]])
print("power limit W:", pynvml.nvmlDeviceGetEnforcedPowerLimit(h) / 1000.0)
