import sys, time
import numpy as np
sys.path.insert(0, '.')
import torch
import qspectra_b200 as qb
from qspectra_b200 import systems, engine, _capi

E = 10000
model = qb.RedfieldModel(systems.fmo(), hilbert_subspace='e', unit_convert=qb.CM_FS, secular=False)
psi0 = np.eye(7)[0]
def sync(): torch.cuda.synchronize(); return time.perf_counter()
for it in range(6):
    t0 = sync()
    shifts = model.hamiltonian.sampled_site_shifts(E)
    t1 = sync()
    eom = model.ensemble_eom(E, False, 'ee')
    t2 = sync()
    t = np.arange(0, 1000.0, model.time_step)
    y0 = model.density_matrix_to_state_vector(np.outer(psi0, psi0).astype(complex), 'ee')
    y0_dev = _capi.to_device(y0).reshape(1, -1).expand(E, -1).contiguous()
    prop = eom.propagator(t[1] - t[0])
    t3 = sync()
    out = eom.propagate(y0_dev, t, generators=np.arange(E), return_device=True)
    t4 = sync()
    mean = engine.reduce_members(out, 1.0 / E)
    res = mean.cpu().numpy()
    t5 = sync()
    full0 = sync()
    _, rho = qb.simulate_dynamics(model, psi0, 1000.0, ensemble_size=E)
    full1 = sync()
    print('iter %d: sample %.1f  build+wrap %.1f  expm %.1f  stepping %.1f  reduce+d2h %.1f | simulate_dynamics total %.1f ms'
          % (it, 1e3*(t1-t0), 1e3*(t2-t1), 1e3*(t3-t2), 1e3*(t4-t3), 1e3*(t5-t4), 1e3*(full1-full0)))
