"""CPU study behind the re-do threshold of the INT8 WPE correlation build:
smallest Cholesky pivot ratio L_jj^2 / R_jj vs the condition number of the equilibrated normal
equations vs the change of the WPE output when the Gram matrix is perturbed at the level of the
INT8 path's measured error (2e-10 sqrt(R_ii R_jj)).  Developer tool, not part of the product."""
import numpy as np, sys
sys.path.insert(0, str(__import__('pathlib').Path(__file__).resolve().parent.parent))
from pb_chime5_b200 import synth
rng=np.random.default_rng(0)
D,L,delay=8,10,2
def gram(Y,inv):
    D,T=Y.shape; rows=[]
    for k in range(L):
        s=delay+k; r=np.zeros((D,T),complex); r[:,s:]=Y[:,:T-s]; rows.append(r)
    A=np.concatenate(rows,0); R=(A*inv)@A.conj().T; P=(A*inv)@Y.conj().T
    return R,P
def study(Y,label):
    lam=np.mean(np.abs(Y)**2,0); lam=np.maximum(lam,1e-10*lam.max()); inv=1/lam
    R,P=gram(Y,inv)
    d=np.sqrt(np.diag(R).real); 
    Lc=np.linalg.cholesky(R); ratio=(np.diag(Lc).real**2/np.diag(R).real).min()
    Re=R/np.outer(d,d); cond=np.linalg.cond(Re)
    G=np.linalg.solve(R,P)
    errs=[]
    for _ in range(3):
        E=(rng.standard_normal(R.shape)+1j*rng.standard_normal(R.shape))*2e-10*np.outer(d,d); E=(E+E.conj().T)/2
        EP=(rng.standard_normal(P.shape)+1j*rng.standard_normal(P.shape))*2e-10*np.outer(d,np.ones(P.shape[1]))*np.sqrt((np.abs(Y)**2*inv).sum(1))[None,:]
        G2=np.linalg.solve(R+E,P+EP)
        X=Y-G.conj().T@np.concatenate([np.pad(Y[:,:Y.shape[1]-(delay+k)],((0,0),(delay+k,0))) for k in range(L)],0)
        X2=Y-G2.conj().T@np.concatenate([np.pad(Y[:,:Y.shape[1]-(delay+k)],((0,0),(delay+k,0))) for k in range(L)],0)
        errs.append(np.abs(X2-X).max()/np.abs(X).max())
    print(f'{label:28s} min pivot ratio {ratio:9.2e}  cond(equil) {cond:9.2e}  X rel err from 2e-10 Gram noise {max(errs):8.1e}')
# synthetic benchmark-like
Obs,_=synth.make_utterance(3,D=D,T=941,F=4,K=3)
for f in range(2): study(Obs[:,:,f].astype(complex),f'synthetic bin {f}')
Obs2=Obs.copy(); Obs2[:,3:,:]+=0.4*Obs2[:,:-3,:]
study(Obs2[:,:,0].astype(complex),'synthetic + echo')
# reverberant with varying noise
import torch
for noise in (0.5,0.1,0.02,0.005):
    obs,_=synth.make_reverberant_audio(3,D=D,N=32000,K=3,noise=noise)
    from oracle import gss_oracle as o
    Y=o.stft(obs.astype(np.float64)) if hasattr(o,'stft') else None
    Y=np.asarray(Y)  # (D,T,F)
    for f in (40,300):
        study(Y[:,:,f],f'reverb noise={noise} bin {f}')
