"""FMO HEOM depth-4 disorder ensemble (config 3 batched): RHS/s and HBM fraction."""
import sys, time, numpy as np
sys.path.insert(0, '.')
import torch, qspectra_b200 as qb
from qspectra_b200 import systems
E = int(sys.argv[1]) if len(sys.argv) > 1 else 512
depth = int(sys.argv[2]) if len(sys.argv) > 2 else 4
model = qb.HEOMModel(systems.fmo(), hilbert_subspace='e', unit_convert=qb.CM_FS, level_cutoff=depth, K=1)
t0 = time.perf_counter()
eom = model.ensemble_eom(E, False, 'ee')
print('ensemble handle: %.1f ms, n_ado %d, dim %d' % (1e3 * (time.perf_counter() - t0), eom.n_ado, eom.dim))
y0 = model.density_matrix_to_state_vector(np.diag(np.eye(7)[0]).astype(complex), 'ee')
y0 = torch.from_numpy(y0).cuda().reshape(1, -1).expand(E, -1).contiguous()
t = model.time_step * np.arange(21)
for _ in range(3):
    out = eom.propagate(y0, t, save=('ado0',), generators=np.arange(E), return_device=True)
    last = eom.last
rhs_s = last['rhs'] / (last['kernel_ms'] * 1e-3)
print('E=%d depth %d: %.2f ms, %d rhs (columns x applications), %.3e RHS/s, %.1f GB/s algorithmic = %.1f %% of 6540.8' % (
    E, depth, last['kernel_ms'], last['rhs'], rhs_s, rhs_s * 32 * eom.dim / 1e9, rhs_s * 32 * eom.dim / 1e9 / 65.408))
