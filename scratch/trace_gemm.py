import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ot_vae_lightning_b200 import kernels as K, _native as N
lib = N.load()
buf = torch.zeros(64, dtype=torch.int64, device='cuda')
lib.otkdbg_set_gemm_trace.argtypes = [ctypes.c_void_p]
for (M, Nn, Kd, eng) in [(4096, 512, 512, 3), (4096, 512, 512, 2), (32768, 512, 512, 3)]:
    A = torch.randn(M, Kd, device='cuda'); B = torch.randn(Kd, Nn, device='cuda')
    K.gemm(A, B, engine=eng, nn=True); torch.cuda.synchronize()
    lib.otkdbg_set_gemm_trace(ctypes.c_void_p(buf.data_ptr()))
    buf.zero_()
    K.gemm(A, B, engine=eng, nn=True); torch.cuda.synchronize()
    lib.otkdbg_set_gemm_trace(None)
    t = buf.cpu().tolist(); t0 = t[0]
    print(f"M={M} N={Nn} K={Kd} engine={eng}: acc_full at {t[1]-t0}, epilogue end {t[2]-t0}, exit {t[3]-t0}")
    print("  MMA sees full[kt] at:", [t[8+i]-t0 for i in range(16)])
    print("  TMA sees empty[kt] at:", [t[40+i]-t0 for i in range(16)])
