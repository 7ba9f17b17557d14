"""Per-RHS time of the depth-8 FMO HEOM Taylor kernel (experiment harness)."""
import sys, os, numpy as np
sys.path.insert(0, '.')
import torch, qspectra_b200 as qb
from qspectra_b200 import systems
depth = int(sys.argv[1]) if len(sys.argv) > 1 else 8
model = qb.HEOMModel(systems.fmo(), hilbert_subspace='e', unit_convert=qb.CM_FS, level_cutoff=depth, K=1)
eom = model.equation_of_motion('ee')
y0 = model.density_matrix_to_state_vector(np.diag(np.eye(7)[0]).astype(complex), 'ee')
y0 = torch.from_numpy(y0).cuda().reshape(1, -1)
def run(n):
    t = model.time_step * np.arange(n + 1)
    best = None
    for _ in range(2):
        eom.propagate(y0, t, save=('ado0',), return_device=True)
        if best is None or eom.last['kernel_ms'] < best['kernel_ms']:
            best = dict(eom.last)
    return best
a, b = run(2), run(12)
print('DBG=%s VARIANT=%s: %.1f us per RHS (marginal; %d rhs in %.2f ms, %d rhs in %.2f ms)' % (
    os.environ.get('QSX_HEOM_DBG', '0'), os.environ.get('QSX_HEOM_VARIANT', '-'),
    1e3 * (b['kernel_ms'] - a['kernel_ms']) / (b['rhs'] - a['rhs']), a['rhs'], a['kernel_ms'], b['rhs'], b['kernel_ms']))
