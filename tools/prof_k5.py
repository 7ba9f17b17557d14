"""simulate_dynamics of a 4000-member FMO Redfield ensemble from host objects (for ncu captures of K5)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import qspectra_b200 as qb
from qspectra_b200 import systems
E = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
model = qb.RedfieldModel(systems.fmo(), hilbert_subspace='e', unit_convert=qb.CM_FS, secular=False)
for _ in range(3):
    t, rho = qb.simulate_dynamics(model, np.eye(7)[0], 1000, ensemble_size=E)
torch.cuda.synchronize()
print('trace error', float(np.abs(np.einsum('tii->t', rho) - 1).max()))
