"""One resident headline step on the Hermitian-coordinate path (for ncu): argv[1] = members."""
import os
import sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import qspectra_b200 as qb
from qspectra_b200 import systems, _capi, engine

E = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
model = qb.RedfieldModel(systems.fmo(), hilbert_subspace='e', unit_convert=qb.CM_FS, secular=False)
t = np.arange(0, 1000.0, model.time_step)
psi0 = np.eye(7)[0]
y0 = model.density_matrix_to_state_vector(np.outer(psi0, psi0).astype(complex), 'ee')
eom = model.ensemble_eom(E, False, 'ee', member0=0)
y0_dev = _capi.to_device(y0).reshape(1, -1).expand(E, -1).contiguous()
for _ in range(3):
    eom.__dict__.pop('_propagators', None)
    out = eom.propagate(y0_dev, t, generators=np.arange(E), return_device=True, hermitian_state=True, packed=True)
    mean = engine.reduce_members(out, 1.0 / E)
torch.cuda.synchronize()
print('trace error', float(np.abs(np.einsum('tii->t', mean.cpu().numpy().reshape(len(t), 7, 7, order='F')) - 1).max()))
