"""cProfile of the end-to-end ensemble call (host-side overhead hunt)."""
import sys, cProfile, pstats, io
sys.path.insert(0, '.')
import numpy as np, torch
import qspectra_b200 as qb
from qspectra_b200 import systems
model = qb.RedfieldModel(systems.fmo(), hilbert_subspace='e', unit_convert=qb.CM_FS, secular=False)
psi0 = np.eye(7)[0]
for _ in range(3):
    qb.simulate_dynamics(model, psi0, 1000.0, liouville_subspace='ee', ensemble_size=10000)
pr = cProfile.Profile()
pr.enable()
for _ in range(5):
    qb.simulate_dynamics(model, psi0, 1000.0, liouville_subspace='ee', ensemble_size=10000)
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats('tottime').print_stats(28)
print(s.getvalue()[:6000])
