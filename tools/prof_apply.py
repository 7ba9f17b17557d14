import sys, numpy as np
sys.path.insert(0, '.')
import torch, qspectra_b200 as qb
from qspectra_b200 import systems
depth = int(sys.argv[1]) if len(sys.argv) > 1 else 8
model = qb.HEOMModel(systems.fmo(), hilbert_subspace='e', unit_convert=qb.CM_FS, level_cutoff=depth, K=1)
eom = model.equation_of_motion('ee')
rng = np.random.RandomState(0)
y = rng.randn(eom.dim) + 1j * rng.randn(eom.dim)
for _ in range(4):
    eom.apply(y[None])
torch.cuda.synchronize()
