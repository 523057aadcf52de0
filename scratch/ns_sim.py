import torch, math
def spd(d,kappa,seed=3):
    g=torch.Generator().manual_seed(seed)
    q,_=torch.linalg.qr(torch.randn(d,d,generator=g,dtype=torch.double))
    lam=torch.logspace(-math.log10(kappa),0,d,dtype=torch.double)
    return (q*lam)@q.T
def ns(A,dt,iters,nt=True):
    c=A.norm()
    Y=(A/c).to(dt); Z=torch.eye(A.shape[0],dtype=dt)
    I=torch.eye(A.shape[0],dtype=dt)
    res=[]
    for k in range(iters):
        P = Z@Y.T if nt else Z@Y
        res.append(((P-I).double().norm()**2).item())
        T=1.5*I-0.5*P
        Yn = Y@T.T if nt else Y@T
        Zn = T@Z.T if nt else T@Z
        Y,Z=Yn,Zn
    return Y.double()*c.sqrt(), Z.double()/c.sqrt(), res
for d,kappa in [(128,1e4),(16,6e4)]:
    A=spd(d,kappa)
    lam,V=torch.linalg.eigh(A); R=(V*lam.sqrt())@V.T
    for dt in (torch.float64,torch.float32):
      for nt in (True,False):
        Y,Z,res=ns(A,dt,60,nt)
        print(d,kappa,dt,nt,'final err',((Y-R).norm()/R.norm()).item())
        print('  res', ' '.join(f'{r:.1e}' for r in res[::3]))
