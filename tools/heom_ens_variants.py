"""Depth-4 FMO HEOM disorder ensemble (one generator per column) across tile variants:
time per batched RHS sweep and result cross-check against the first variant."""
import sys, os, numpy as np
sys.path.insert(0, '.')
import torch, qspectra_b200 as qb
from qspectra_b200 import systems
variants = sys.argv[1] if len(sys.argv) > 1 else 'mAP'
E = int(sys.argv[2]) if len(sys.argv) > 2 else 256
model = qb.HEOMModel(systems.fmo(), hilbert_subspace='e', unit_convert=qb.CM_FS, level_cutoff=4, K=1)
eom = model.ensemble_eom(E, False, 'ee')
y0 = model.density_matrix_to_state_vector(np.diag(np.eye(7)[0]).astype(complex), 'ee')
y0 = torch.from_numpy(y0).cuda().reshape(1, -1).expand(E, -1).contiguous()
t = model.time_step * np.arange(11)
ref = None
for v in variants:
    os.environ['QSX_HEOM_VARIANT'] = v
    for _ in range(2):
        out = eom.propagate(y0, t, save=('ado0',), generators=np.arange(E), return_device=True)
    last = dict(eom.last)
    out = out.cpu().numpy()
    if ref is None:
        ref = out
    err = np.linalg.norm((out - ref).ravel()) / np.linalg.norm(ref.ravel())
    rhs_s = last['rhs'] / (last['kernel_ms'] * 1e-3)
    print('ens E=%d VARIANT=%s: %.2f ms, %.3e RHS/s, %.1f GB/s algorithmic, rel diff vs %s: %.2e'
          % (E, v, last['kernel_ms'], rhs_s, rhs_s * 32 * eom.dim / 1e9, variants[0], err), flush=True)
# which variant agrees with single-member runs (one generator per launch: no member switch)?
os.environ['QSX_HEOM_VARIANT'] = 'b'
singles = []
for e in range(0, E, max(1, E // 8)):
    o = eom.propagate(y0[e:e + 1].contiguous(), t, save=('ado0',), generators=np.array([e]), return_device=True)
    singles.append((e, o.cpu().numpy()[0]))
for v in variants:
    os.environ['QSX_HEOM_VARIANT'] = v
    out = eom.propagate(y0, t, save=('ado0',), generators=np.arange(E), return_device=True).cpu().numpy()
    errs = [np.linalg.norm(out[e] - o) / np.linalg.norm(o) for e, o in singles]
    print('VARIANT=%s vs single-member launches: max rel diff %.2e' % (v, max(errs)), flush=True)
