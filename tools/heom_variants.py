"""A/B harness: per-RHS time of the FMO HEOM propagation kernel for several launch
configurations in one process, with a cross-check of the trajectories against the first one.

usage: python tools/heom_variants.py [depth] [spec spec ...]
spec = comma-separated KEY=VALUE settings, e.g.
    method=poly,QSX_HEOM_FLOW=0,QSX_HEOM_GRID=148
`method` picks the integrator (poly | taylor); every other key is exported to the
environment (QSX_HEOM_VARIANT, QSX_HEOM_FLOW, QSX_HEOM_GRID, ...)."""
import os
import sys

import numpy as np

sys.path.insert(0, '.')
import torch
import qspectra_b200 as qb
from qspectra_b200 import systems

depth = int(sys.argv[1]) if len(sys.argv) > 1 else 8
specs = sys.argv[2:] or ['method=poly', 'method=taylor', 'method=taylor,QSX_HEOM_VARIANT=b']
model = qb.HEOMModel(systems.fmo(), hilbert_subspace='e', unit_convert=qb.CM_FS,
                     level_cutoff=depth, K=1)
eom = model.equation_of_motion('ee')
y0 = model.density_matrix_to_state_vector(np.diag(np.eye(7)[0]).astype(complex), 'ee')
y0 = torch.from_numpy(y0).cuda().reshape(1, -1)
alg_bytes = 32.0 * eom.dim


def run(n, method):
    t = model.time_step * np.arange(n + 1)
    best, out = None, None
    for _ in range(2):
        out = eom.propagate(y0, t, save=('ado0',), method=method, return_device=True)
        if best is None or eom.last['kernel_ms'] < best['kernel_ms']:
            best = dict(eom.last)
    return best, out


ref = None
for spec in specs:
    for k in [k for k in os.environ if k.startswith('QSX_HEOM_')]:
        os.environ.pop(k)
    method = 'poly'
    for kv in spec.split(','):
        k, v = kv.split('=')
        if k == 'method':
            method = v
        else:
            os.environ[k] = v
    try:
        a, _ = run(3, method)
        b, out = run(13, method)
    except Exception as exc:          # keep the other configurations running
        print('depth %d %s: FAILED %r' % (depth, spec, exc), flush=True)
        continue
    out = out.cpu().numpy()
    if ref is None:
        ref = out
    err = np.linalg.norm((out - ref).ravel()) / np.linalg.norm(ref.ravel())
    us = 1e3 * (b['kernel_ms'] - a['kernel_ms']) / (b['rhs'] - a['rhs'])
    print('depth %d %-60s %6.1f us/RHS = %5.0f GB/s alg (%d rhs %.2f ms | %d rhs %.2f ms; %.2f ms per interval) '
          'rel diff vs first %.2e' % (depth, spec, us, alg_bytes / us / 1e3, a['rhs'], a['kernel_ms'],
                                      b['rhs'], b['kernel_ms'], (b['kernel_ms'] - a['kernel_ms']) / 10, err),
          flush=True)
