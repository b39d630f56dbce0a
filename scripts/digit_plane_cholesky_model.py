#!/usr/bin/env python
"""CPU model (numpy, exact integer digit products) of a blocked Cholesky whose rank-NB trailing updates use digit planes -- the
experiment behind the choice of digit radix, kept pair set and guard of csrc/ozaki_i8.cu (profiles/r02_radix256.md).

    python scripts/digit_plane_cholesky_model.py N NB fp64 old7 new6 new6diag new6all new7 old8 ...

modes: fp64 (plain), oldS (radix 128, S planes, round-1 scheme), newS (radix 256, S planes) + suffixes: diag (equal-plane pair
(S/2, S/2)), all (every pair of order S), p15 / p24 (single order-6 pair families), sign (Thue-Morse column signs).
Prints, against the fp64 run: relative error of logdet, of the quadratic form, of sum(Sigma^-1 d) (the mean-constant gradient) and
the residual max|L L^T - Sigma|.  Analysis script; uses the oracle's Gram only to build the test matrix."""
import numpy as np, sys, time
sys.path.insert(0,'/root/repo')
import oracle as o
import scipy.linalg as sla
def tm_sign(k): 
    return np.array([-1.0 if bin(i).count('1')&1 else 1.0 for i in range(k)])
def slice256(X, s, sign=None):
    mx=np.abs(X).max(1); mx=np.where(mx==0,1,mx)
    e=np.floor(np.log2(mx)).astype(int)+2
    e=e+(mx*2.0**(-e)>0.494)
    R=X*2.0**(-e)[:,None]
    if sign is not None: R=R*sign[None,:]
    I=np.rint(R*2.0**(8*s)).astype(np.int64); Q=[None]*s
    for p in range(s-1,-1,-1):
        d=((I+128)&255)-128; Q[p]=d.astype(np.float64); I=(I-d)>>8
    assert np.all(I==0)
    return Q, 2.0**e
def slice128(X, s):
    mx=np.abs(X).max(1); mx=np.where(mx==0,1,mx)
    e=np.floor(np.log2(mx)).astype(int)+2
    R=X*2.0**(-e)[:,None]; Q=[]
    for p in range(s):
        R=R*128; d=np.rint(R); Q.append(d); R=R-d
    return Q, 2.0**e
def xxt(X, mode):
    if mode=='fp64': return X@X.T
    if mode.startswith('old'):
        s=int(mode[3:]); Q,sc=slice128(X,s); beta=7; diag=False
    else:
        s=int(mode[3]); Q,sc=slice256(X,s, tm_sign(X.shape[1]) if 'sign' in mode else None); beta=8; diag='diag' in mode
    acc=0
    for t in range(s):
        P=sum(Q[p]@Q[t-p].T for p in range(t+1))
        acc=acc+P*2.0**(-beta*(t+2))
    if diag and s%2==0:
        acc=acc+(Q[s//2]@Q[s//2].T)*2.0**(-beta*(s+2))
    if 'p24' in mode:
        acc=acc+(Q[2]@Q[4].T+Q[4]@Q[2].T)*2.0**(-beta*(s+2))
    if 'p15' in mode:
        acc=acc+(Q[1]@Q[5].T+Q[5]@Q[1].T)*2.0**(-beta*(s+2))
    if 'all' in mode:
        P=sum(Q[p]@Q[s-p].T for p in range(1,s))
        acc=acc+P*2.0**(-beta*(s+2))
    return acc*sc[:,None]*sc[None,:]
def chol(S, NB, mode):
    A=S.copy(); N=A.shape[0]
    for j in range(0,N,NB):
        e=min(j+NB,N)
        A[j:e,j:e]=np.linalg.cholesky(A[j:e,j:e])
        if e<N:
            A[e:,j:e]=sla.solve_triangular(A[j:e,j:e],A[e:,j:e].T,lower=True).T
            A[e:,e:]-=xxt(A[e:,j:e],mode)
    return np.tril(A)
N=int(sys.argv[1]); NB=int(sys.argv[2]); D=8
rng=np.random.default_rng(8192); X=rng.uniform(-2,2,(N,D)); y=np.sin(X[:,:1])+0.1*rng.standard_normal((N,1))
S=o.gram("rbf",X,np.full(D,3.0),1.0)+(0.03**2+1e-6)*np.eye(N)
d=y[:,0]
res={}
for mode in sys.argv[3:]:
    t0=time.time(); L=chol(S,NB,mode)
    w=sla.solve_triangular(L,d,lower=True); alpha=sla.solve_triangular(L.T,w,lower=False)
    res[mode]=(2*np.log(np.diag(L)).sum(), w@w, alpha.sum(), np.abs(L@L.T-S).max())
    print(mode,'%.1fs'%(time.time()-t0),flush=True)
b=res['fp64']
for m,v in res.items():
    print(m,'logdet %.1e quad %.1e sum(alpha) %.1e resid %.1e'%(abs(v[0]-b[0])/abs(b[0]),abs(v[1]-b[1])/abs(b[1]),abs(v[2]-b[2])/abs(b[2]),v[3]))
