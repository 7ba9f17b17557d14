"""Where the time of an FMO ensemble third-order response goes (builds vs stages)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import torch
import qspectra_b200 as qb
from qspectra_b200 import systems, engine

mf = qb.RedfieldModel(systems.fmo(), hilbert_subspace="gef", unit_convert=qb.CM_FS)
t2f = np.linspace(0, 1000, 5)
for E in (4, 32):
    qb.third_order_response(mf, 1000, population_times=t2f, ensemble_size=E)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    qb.third_order_response(mf, 1000, population_times=t2f, ensemble_size=E)
    torch.cuda.synchronize()
    print("FMO third-order 197x5x197, E=%d: %.1f ms" % (E, 1e3 * (time.perf_counter() - t0)))
    for ss in ("ge", "gg", "ee", "fe", "eg", "ef"):
        try:
            torch.cuda.synchronize(); t0 = time.perf_counter()
            eom = mf.ensemble_eom(E, False, ss)
            torch.cuda.synchronize(); tb = time.perf_counter() - t0
            dim = eom.dim
            y0 = np.ones((E * 197, dim), complex)
            t0 = time.perf_counter()
            eom.propagate(y0, t2f, t0=0, generators=np.repeat(np.arange(E), 197), return_device=True)
            torch.cuda.synchronize(); tp = time.perf_counter() - t0
            y1 = np.ones((E, dim), complex)
            t0 = time.perf_counter()
            eom.propagate(y1, np.arange(0, 1000, mf.time_step), generators=np.arange(E), return_device=True)
            torch.cuda.synchronize(); t1 = time.perf_counter() - t0
            print("  %s dim %d: build %.1f ms, t2-stage %.1f ms, t1-stage %.1f ms" % (ss, dim, 1e3 * tb, 1e3 * tp, 1e3 * t1))
        except Exception as e:
            print("  %s: %r" % (ss, e))
