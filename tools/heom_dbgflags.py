"""Experiment: per-RHS time of the depth-8 kernel under QSX_HEOM_DBG switches (needs a
-DQSX_HEOM_DBG_FLAGS build): 1 = gathers redirected to the own (L2-hot) tile, 2 = no gathers,
4 = no H rho product."""
import sys, os, numpy as np
sys.path.insert(0, '.')
import torch, qspectra_b200 as qb
from qspectra_b200 import systems
depth = int(sys.argv[1]) if len(sys.argv) > 1 else 8
flags = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else '0,1,2,4,6').split(',')]
model = qb.HEOMModel(systems.fmo(), hilbert_subspace='e', unit_convert=qb.CM_FS, level_cutoff=depth, K=1)
variant = os.environ.get("QSX_HEOM_VARIANT", " ")
eom = model.equation_of_motion('ee')
y0 = model.density_matrix_to_state_vector(np.diag(np.eye(7)[0]).astype(complex), 'ee')
y0 = torch.from_numpy(y0).cuda().reshape(1, -1)
# fixed number of Taylor terms irrespective of what the switches do to the numbers: RK4 sub-steps
def run(n):
    t = model.time_step * np.arange(n + 1)
    best = None
    for _ in range(2):
        eom.propagate(y0, t, save=('ado0',), return_device=True)
        if best is None or eom.last['kernel_ms'] < best['kernel_ms']:
            best = dict(eom.last)
    return best
for f in flags:
    os.environ['QSX_HEOM_DBG'] = str(f)
    a, b = run(1), run(3)
    print('DBG=%d: %.1f us per RHS (%d rhs in %.2f ms, %d rhs in %.2f ms)' % (
        f, 1e3 * (b['kernel_ms'] - a['kernel_ms']) / (b['rhs'] - a['rhs']), a['rhs'], a['kernel_ms'], b['rhs'], b['kernel_ms']), flush=True)
