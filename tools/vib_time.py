"""Vibronic (Jonas) dimer HEOM, 8 'e' states, level_cutoff 10 (715 ADOs): RHS/s of the shaped row
tile (default) against the generic tile, one trajectory and a batch of columns."""
import os, sys, numpy as np
sys.path.insert(0, '.')
import torch, qspectra_b200 as qb
from qspectra_b200 import systems
cut = int(sys.argv[1]) if len(sys.argv) > 1 else 10
model = qb.HEOMModel(systems.jonas_dimer(), hilbert_subspace='e', unit_convert=qb.CM_FS, level_cutoff=cut, K=1)
eom = model.equation_of_motion('ee')
psi = np.zeros(eom.M, dtype=complex); psi[0] = 1.0
y0 = torch.from_numpy(model._pad(psi)).cuda().reshape(1, -1)
t = model.time_step * np.arange(21)
ref = {}
for B in (1, 64, 512):
    yb = y0.expand(B, -1).contiguous()
    for variant in (' ', 'g'):
        os.environ.pop('QSX_HEOM_VARIANT', None)
        if variant != ' ':
            os.environ['QSX_HEOM_VARIANT'] = variant
        best = None
        for _ in range(3):
            out = eom.propagate(yb, t, save=('ado0',), return_device=True)
            if best is None or eom.last['kernel_ms'] < best['kernel_ms']:
                best = dict(eom.last)
        out = out.cpu().numpy()
        ref.setdefault(B, out)
        rhs_s = best['rhs'] / (best['kernel_ms'] * 1e-3)
        print('B=%4d variant %s method %-6s: %8.3f ms, %.3e RHS/s, %7.1f GB/s algorithmic (%.1f %% of 6544.7), diff %.1e'
              % (B, variant, best['method'], best['kernel_ms'], rhs_s, rhs_s * 32 * eom.dim / 1e9,
                 rhs_s * 32 * eom.dim / 1e9 / 65.447, np.abs(out - ref[B]).max()), flush=True)
