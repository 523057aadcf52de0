import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ot_vae_lightning_b200.ot as ot
from ot_vae_lightning_b200 import kernels as K
torch.manual_seed(0)
d=16
x = torch.randn(650, d, device='cuda')*2+1
gm = ot.GaussianModel(d, w2_cfg=dict(make_pd=True), dtype=torch.double).cuda()
for lo in range(0,650,100): gm.update(x[lo:lo+100])
x64=x.double()
print('n', gm._n_obs.item(), 'sum err', (gm._running_sum - x64.sum(0)).abs().max().item(), 'cov err', (gm._running_sum_cov - x64.T@x64).abs().max().item())
gm.fit()
print('mean nan', gm.mean.isnan().any().item(), 'orig nan', gm.parametrizations.cov.original.isnan().any().item())
raw = gm.parametrizations.cov.original
sym = K.symmetrize_shift(raw, None)
print('sym nan', sym.isnan().any().item(), 'min_eig', K.min_eig(sym), 'true', torch.linalg.eigvalsh(sym).min().item())
cov = gm.cov
print('cov nan', cov.isnan().any().item())
r, ir = K.sqrtm_pair(cov)
print('sqrtm err', (r@r - cov).abs().max().item(), 'isqrt err', (ir@cov@ir - torch.eye(d,device='cuda',dtype=torch.double)).abs().max().item())
r, ir = K.sqrtm_pair(cov, ridge=1e-8)
print('sqrtm ridge err', (r@r - cov).abs().max().item())
cov2 = cov*1.5 + 0.1*torch.eye(d,device='cuda',dtype=torch.double)
try:
    print('w2', K.w2_gaussian(gm.mean, gm.mean+1, cov, cov2))
except Exception as e: print('w2 fail', e)
try:
    T, w2 = K.transport_operator(cov, cov2, mean_s=gm.mean, mean_t=gm.mean+1)
    print('T err', (T@cov@T - cov2).abs().max().item(), w2)
except Exception as e: print('T fail', e)
for it in (4, 8, 12, 20):
    try:
        T, w2 = K.transport_operator(cov, cov2, iters=it)
        print(it, 'T nan', T.isnan().any().item(), 'T err', (T@cov@T - cov2).abs().max().item())
    except Exception as e: print('T fail', it, e)
