"""Wall times of the public API on BASELINE configs 1 and 4 (ours vs CPU oracle)."""
import sys, time
import numpy as np
sys.path.insert(0, '.')
import torch
import oracle
import qspectra_b200 as qb
from qspectra_b200 import systems
CM_FS = qb.CM_FS
def timed(f, n=3):
    f(); best = 1e9
    for _ in range(n):
        torch.cuda.synchronize(); t0 = time.perf_counter(); r = f(); torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best, r
def rel(a, b): return np.linalg.norm((a - b).ravel()) / np.linalg.norm(b.ravel())
# config 1: dimer secular Redfield absorption
m = qb.RedfieldModel(systems.dimer(), hilbert_subspace='gef', unit_convert=CM_FS, discard_imag_corr=True)
o = oracle.OracleRedfield(systems.dimer(), hilbert_subspace='gef', unit_convert=CM_FS, discard_imag_corr=True)
t1, (f, X) = timed(lambda: qb.absorption_spectra(m, 10000))
t0 = time.perf_counter(); fo, Xo = oracle.absorption_spectra(o, 10000, **oracle.TIGHT); t2 = time.perf_counter() - t0
print('config1 dimer absorption: ours %.2f ms, oracle %.1f ms, rel-L2 %.1e' % (1e3 * t1, 1e3 * t2, rel(X, Xo)))
# config 4: dimer third-order, 50 t2 points, all pathways
t2pts = np.linspace(0, 1000, 50)
t1_, (_, S) = timed(lambda: qb.third_order_response(m, 1000, population_times=t2pts))
t0 = time.perf_counter(); _, So = oracle.third_order_response(o, 1000, population_times=t2pts, **oracle.TIGHT); t2 = time.perf_counter() - t0
print('config4 dimer third-order (103x50x103): ours %.2f ms, oracle %.1f ms, rel-L2 %.1e' % (1e3 * t1_, 1e3 * t2, rel(S, So)))
t1_, (_, S) = timed(lambda: qb.third_order_response(m, 1000, population_times=t2pts, exact_isotropic_average=True), 1)
print('  with exact isotropic average (21 polarisation configs): ours %.1f ms' % (1e3 * t1_))
# FMO third order, 5 t2 points
mf = qb.RedfieldModel(systems.fmo(), hilbert_subspace='gef', unit_convert=CM_FS)
of = oracle.OracleRedfield(systems.fmo(), hilbert_subspace='gef', unit_convert=CM_FS)
t2f = np.linspace(0, 1000, 5)
t1_, (_, S) = timed(lambda: qb.third_order_response(mf, 1000, population_times=t2f), 2)
t0 = time.perf_counter(); _, So = oracle.third_order_response(of, 1000, population_times=t2f, **oracle.TIGHT); t2 = time.perf_counter() - t0
print('config4 FMO third-order (197x5x197): ours %.1f ms, oracle %.1f ms, rel-L2 %.1e' % (1e3 * t1_, 1e3 * t2, rel(S, So)))
# HEOM dimer 2D
hm = qb.HEOMModel(systems.dimer(), hilbert_subspace='gef', unit_convert=CM_FS, level_cutoff=3, low_temp_corr=False)
ho = oracle.OracleHEOM(systems.dimer(), hilbert_subspace='gef', unit_convert=CM_FS, level_cutoff=3, low_temp_corr=False)
t1_, (_, S) = timed(lambda: qb.third_order_response(hm, 1000, population_times=t2pts), 2)
t0 = time.perf_counter(); _, So = oracle.third_order_response(ho, 1000, population_times=t2pts[:5], **oracle.TIGHT); t2 = time.perf_counter() - t0
print('dimer HEOM third-order (103x50x103): ours %.1f ms; oracle with 5 of 50 t2 points %.1f ms' % (1e3 * t1_, 1e3 * t2))
