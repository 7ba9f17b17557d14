"""End-to-end simulate_dynamics of the headline ensemble from host objects: wall time per call."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import qspectra_b200 as qb
from qspectra_b200 import systems
E = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
model = qb.RedfieldModel(systems.fmo(), hilbert_subspace='e', unit_convert=qb.CM_FS, secular=False)
for _ in range(3):
    t, rho = qb.simulate_dynamics(model, np.eye(7)[0], 1000, ensemble_size=E)
torch.cuda.synchronize()
ts = []
for _ in range(8):
    t0 = time.perf_counter()
    t, rho = qb.simulate_dynamics(model, np.eye(7)[0], 1000, ensemble_size=E)
    torch.cuda.synchronize()
    ts.append(time.perf_counter() - t0)
print('e2e ms per call: min %.3f mean %.3f' % (1e3 * min(ts), 1e3 * np.mean(ts)))
