"""HEOM 2D spectrum through the batched response layer: dimer, level_cutoff 3, disorder members x
isotropic average.  python tools/heom2d_time.py [members]"""
import sys, time, numpy as np
sys.path.insert(0, '.')
import torch, qspectra_b200 as qb
from qspectra_b200 import systems, _capi
E = int(sys.argv[1]) if len(sys.argv) > 1 else 16
m = qb.HEOMModel(systems.dimer(disorder=80), hilbert_subspace='gef', unit_convert=qb.CM_FS, level_cutoff=3, K=1)
kw = dict(population_times=np.linspace(0, 200, 3), geometry='-++', polarization='xxxx')
for iso in (False, True):
    run = lambda: qb.two_dimensional_spectra(m, 500, ensemble_size=E, exact_isotropic_average=iso, **kw)
    run()
    torch.cuda.synchronize(); l0 = _capi.kernel_launches(); t0 = time.perf_counter()
    run()
    torch.cuda.synchronize()
    print('dimer HEOM depth 3 2D spectrum, %d members, isotropic %s: %.1f ms, %d launches'
          % (E, iso, 1e3 * (time.perf_counter() - t0), _capi.kernel_launches() - l0), flush=True)
