"""ZOFE FMO 'e' ensemble for ncu captures / timing: python tools/prof_zofe.py [members] [intervals]"""
import sys, numpy as np
sys.path.insert(0, '.')
import torch, qspectra_b200 as qb
from qspectra_b200 import systems
E = int(sys.argv[1]) if len(sys.argv) > 1 else 592
nint = int(sys.argv[2]) if len(sys.argv) > 2 else 4
model = qb.ZOFEModel(systems.fmo(bath='pseudomode'), hilbert_subspace='e', unit_convert=qb.CM_FS)
eom = model.ensemble_eom(E, False, 'ee')
y0 = model.density_matrix_to_state_vector(np.diag(np.eye(7)[0]).astype(complex), 'ee')
y0_dev = torch.from_numpy(y0).cuda().reshape(1, -1).expand(E, -1).contiguous()
t = model.time_step * np.arange(nint + 1)
for _ in range(2):
    eom.propagate(y0_dev, t, save=('ado0',), generators=np.arange(E), return_device=True)
    print(eom.last, flush=True)
