"""Small fixed workloads for ncu captures (see profiles/README.md).
    python tools/prof_run.py dense|heom8|heom4|build [repeat]"""
import sys
import numpy as np
sys.path.insert(0, '.')
import torch
import qspectra_b200 as qb
from qspectra_b200 import systems, engine, _capi

mode = sys.argv[1]
repeat = int(sys.argv[2]) if len(sys.argv) > 2 else 1
if mode in ('dense', 'build'):
    E = 4000
    model = qb.RedfieldModel(systems.fmo(), hilbert_subspace='e', unit_convert=qb.CM_FS, secular=False)
    t = np.arange(0, 1000.0, model.time_step)
    y0 = model.density_matrix_to_state_vector(np.diag(np.eye(7)[0]).astype(complex), 'ee')
    for _ in range(repeat):
        eom = model.ensemble_eom(E, False, 'ee')
        if mode == 'dense':
            y0_dev = _capi.to_device(y0).reshape(1, -1).expand(E, -1).contiguous()
            out = eom.propagate(y0_dev, t, generators=np.arange(E), return_device=True)
            print(eom.last)
else:
    depth = 8 if mode == 'heom8' else 4
    model = qb.HEOMModel(systems.fmo(), hilbert_subspace='e', unit_convert=qb.CM_FS, level_cutoff=depth, K=1)
    eom = model.equation_of_motion('ee')
    y0 = model.density_matrix_to_state_vector(np.diag(np.eye(7)[0]).astype(complex), 'ee')
    import os
    t = model.time_step * np.arange(int(os.environ.get('NPTS', 2 if depth == 8 else 11)))
    y0_dev = torch.from_numpy(y0).cuda().reshape(1, -1)
    for _ in range(repeat):
        eom.propagate(y0_dev, t, save=('ado0',), return_device=True)
        print(eom.last)
torch.cuda.synchronize()
