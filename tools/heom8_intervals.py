"""Depth-8 FMO HEOM: RHS rate of the default integrator for different run lengths (share of the
adaptive-Taylor pilot intervals) -- argv: interval counts."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import qspectra_b200 as qb
from qspectra_b200 import systems

model = qb.HEOMModel(systems.fmo(), hilbert_subspace='e', unit_convert=qb.CM_FS, level_cutoff=8, K=1)
eom = model.equation_of_motion('ee')
y0 = model.density_matrix_to_state_vector(np.diag(np.eye(7)[0]).astype(complex), 'ee')
y0_dev = torch.from_numpy(y0).cuda().reshape(1, -1)
eom.propagate(y0_dev, model.time_step * np.arange(2), save=('ado0',), return_device=True)
for n_int in [int(a) for a in sys.argv[1:]] or [12, 24, 48]:
    t = model.time_step * np.arange(n_int + 1)
    best = None
    for _ in range(2):
        eom.propagate(y0_dev, t, save=('ado0',), return_device=True)
        if best is None or eom.last['kernel_ms'] < best['kernel_ms']:
            best = dict(eom.last)
    rate = best['rhs'] / (best['kernel_ms'] * 1e-3)
    print('REPILOT=%s intervals %3d: %d RHS, %.2f ms, %.0f RHS/s, %.1f us per RHS, %.3f of 6544.7 GB/s'
          % (os.environ.get('QSX_HEOM_REPILOT', 'default'), n_int, best['rhs'], best['kernel_ms'], rate, 1e6 / rate,
             rate * 32 * eom.dim / 6544.7e9), flush=True)
