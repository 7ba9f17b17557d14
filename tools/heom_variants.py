"""A/B harness: per-RHS time of the FMO HEOM Taylor kernel for several tile variants
(QSX_HEOM_VARIANT letters) in one process, with a result cross-check against the first one.
usage: python tools/heom_variants.py [depth] [variants, e.g. mABCD]"""
import sys, os, numpy as np
sys.path.insert(0, '.')
import torch, qspectra_b200 as qb
from qspectra_b200 import systems
depth = int(sys.argv[1]) if len(sys.argv) > 1 else 8
variants = sys.argv[2] if len(sys.argv) > 2 else 'mABCD'
model = qb.HEOMModel(systems.fmo(), hilbert_subspace='e', unit_convert=qb.CM_FS, level_cutoff=depth, K=1)
eom = model.equation_of_motion('ee')
y0 = model.density_matrix_to_state_vector(np.diag(np.eye(7)[0]).astype(complex), 'ee')
y0 = torch.from_numpy(y0).cuda().reshape(1, -1)


def run(n):
    t = model.time_step * np.arange(n + 1)
    best, out = None, None
    for _ in range(2):
        out = eom.propagate(y0, t, save=('ado0',), return_device=True)
        if best is None or eom.last['kernel_ms'] < best['kernel_ms']:
            best = dict(eom.last)
    return best, out


ref = None
for v in variants:
    os.environ['QSX_HEOM_VARIANT'] = v
    a, _ = run(2)
    b, out = run(12)
    out = out.cpu().numpy()
    if ref is None:
        ref = out
    err = np.linalg.norm((out - ref).ravel()) / np.linalg.norm(ref.ravel())
    us = 1e3 * (b['kernel_ms'] - a['kernel_ms']) / (b['rhs'] - a['rhs'])
    print('depth %d VARIANT=%s: %.1f us per RHS (%d rhs in %.2f ms, %d rhs in %.2f ms) rel diff vs %s: %.2e'
          % (depth, v, us, a['rhs'], a['kernel_ms'], b['rhs'], b['kernel_ms'], variants[0], err), flush=True)
