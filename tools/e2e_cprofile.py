"""cProfile of the end-to-end ensemble call (host-side share of the 1e4-member simulate_dynamics)."""
import os, sys, cProfile, pstats
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import qspectra_b200 as qb
from qspectra_b200 import systems
model = qb.RedfieldModel(systems.fmo(), hilbert_subspace='e', unit_convert=qb.CM_FS, secular=False)
for _ in range(3):
    qb.simulate_dynamics(model, np.eye(7)[0], 1000, ensemble_size=10000)
pr = cProfile.Profile()
pr.enable()
for _ in range(20):
    qb.simulate_dynamics(model, np.eye(7)[0], 1000, ensemble_size=10000)
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(28)
