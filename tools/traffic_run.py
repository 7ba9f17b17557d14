"""One HEOM propagation for ncu DRAM-traffic captures: python tools/traffic_run.py case n_intervals
cases: d4 (FMO depth 4), ens512 (512-member depth-4 batch), vib (vibronic dimer), vib64 (64 columns).
Prints the RHS count; profiles/traffic.json holds (bytes(n2) - bytes(n1)) / (rhs(n2) - rhs(n1))."""
import sys, numpy as np
sys.path.insert(0, '.')
import torch, qspectra_b200 as qb
from qspectra_b200 import systems
case, nint = sys.argv[1], int(sys.argv[2])
if case in ('d4', 'ens512'):
    model = qb.HEOMModel(systems.fmo(), hilbert_subspace='e', unit_convert=qb.CM_FS, level_cutoff=4, K=1)
    y0 = model.density_matrix_to_state_vector(np.diag(np.eye(7)[0]).astype(complex), 'ee')
    if case == 'd4':
        eom, B, gens = model.equation_of_motion('ee'), 1, None
    else:
        eom, B, gens = model.ensemble_eom(512, False, 'ee'), 512, np.arange(512)
else:
    model = qb.HEOMModel(systems.jonas_dimer(), hilbert_subspace='e', unit_convert=qb.CM_FS, level_cutoff=10, K=1)
    eom = model.equation_of_motion('ee')
    psi = np.zeros(eom.M, dtype=complex); psi[0] = 1.0
    y0 = model._pad(psi)
    B, gens = (64 if case == 'vib64' else 1), None
yb = torch.from_numpy(y0).cuda().reshape(1, -1).expand(B, -1).contiguous()
t = model.time_step * np.arange(nint + 1)
eom.propagate(yb, t, save=('ado0',), generators=gens, return_device=True)
print('RHS', eom.last['rhs'], flush=True)
