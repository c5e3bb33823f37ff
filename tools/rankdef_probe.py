"""Developer probe (GPU): where does the device EM leave the oracle on a rank-deficient class?
Prints, per EM iteration count, the max posterior error and the error of the class models."""
import sys
from pathlib import Path
import numpy as np
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from oracle import gss_oracle as oracle
from pb_chime5_b200 import ops, synth

Obs, act = synth.make_utterance(31, D=8, T=200, F=6, K=3)
act[1] = False
act[1, 40:45] = True
O = Obs.astype(np.complex128)
dev = torch.device('cuda')
Y = ops.pack_dtf_to_fdt(torch.from_numpy(Obs).to(dev)[None])
A = torch.from_numpy(act.astype(np.uint8))[None].to(dev)
for it in (1, 2, 3, 4, 6, 8, 10):
    ref, models = oracle.gss_posteriors(O, act, it, return_models=True)
    post, model = ops.cacgmm(Y, A, it, return_model=True)
    got = ops.unpack_fkt_to_ktf(post)[0].cpu().numpy()
    err = np.abs(got - ref)
    k, t, f = np.unravel_index(err.argmax(), err.shape)
    weight, eigvec, eigval = models
    # oracle covariance scaled to unit trace, per (f, k)
    cov_o = np.einsum('fkde,fke,fkge->fkdg', eigvec, eigval, eigvec.conj())
    cov_o /= np.trace(cov_o, axis1=-2, axis2=-1).real[..., None, None]
    cov_d = model['covariance'][0].cpu().numpy()
    ce = np.abs(cov_d - cov_o).max(axis=(-1, -2))
    print(f'it {it:2d}: post err max {err.max():.2e} at k={k} t={t} f={f}  q99 {np.quantile(err, 0.99):.1e}  '
          f'cov err per class {ce.max(axis=0)}  weight err {np.abs(model["weight"][0].cpu().numpy() - weight[..., 0]).max():.1e}')
    if it == 10:
        print('eigval oracle f=%d:' % f, eigval[f])
        print('posterior at worst frame: got', got[:, t, f], 'ref', ref[:, t, f], 'act', act[:, t])
