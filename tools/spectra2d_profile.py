"""cProfile + device-time split of the FMO 2D-spectrum workload of bench.py (spectra2d leg, one GPU)."""
import sys, time, cProfile, pstats, io
sys.path.insert(0, '.')
import numpy as np, torch
import qspectra_b200 as qb
from qspectra_b200 import systems, parallel, _capi
E = int(sys.argv[1]) if len(sys.argv) > 1 else 32
model = qb.RedfieldModel(systems.fmo(), hilbert_subspace='gef', unit_convert=qb.CM_FS, secular=False)
kw = dict(population_times=np.linspace(0, 1000, 5), geometry='-++', polarization='xxxx',
          exact_isotropic_average=True, dst=0)
once = lambda: parallel.two_dimensional_spectra_sharded(model, 1000, E, **kw)
once(); once()
torch.cuda.synchronize()
l0 = _capi.kernel_launches()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
pr = cProfile.Profile()
t0 = time.perf_counter(); e0.record()
pr.enable()
once()
pr.disable()
e1.record(); torch.cuda.synchronize()
print('E=%d: wall %.1f ms, between events %.1f ms, %d launches' % (E, 1e3 * (time.perf_counter() - t0), e0.elapsed_time(e1), _capi.kernel_launches() - l0))
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats('tottime').print_stats(18)
print(s.getvalue()[:3500])
