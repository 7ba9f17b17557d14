"""2-GPU check of the sharded third-order response against the single-process result.
torchrun --nproc-per-node 2 tools/mg_check.py"""
import os, sys
import numpy as np
sys.path.insert(0, '.')
import torch, torch.distributed as dist
import qspectra_b200 as qb
from qspectra_b200 import systems, parallel
local = int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
model = qb.RedfieldModel(systems.dimer(disorder=80), hilbert_subspace='gef', unit_convert=qb.CM_FS, discard_imag_corr=True)
t2 = np.linspace(0, 200, 3)
ticks, S = parallel.third_order_response_sharded(model, 300, 5, population_times=t2)
t, rho = parallel.simulate_dynamics_sharded(qb.RedfieldModel(systems.fmo(), hilbert_subspace='e', unit_convert=qb.CM_FS, secular=False),
                                            np.eye(7)[0], 300, ensemble_size=7)
(f1, _, f3), X2d = parallel.two_dimensional_spectra_sharded(model, 300, 5, population_times=t2)
import time
mf = qb.RedfieldModel(systems.fmo(), hilbert_subspace='gef', unit_convert=qb.CM_FS)
E = 64 * dist.get_world_size()
for _ in range(2):
    torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
    parallel.two_dimensional_spectra_sharded(mf, 1000, E, population_times=np.linspace(0, 1000, 5))
    torch.cuda.synchronize(); dist.barrier(); dt_fmo = time.perf_counter() - t0
if dist.get_rank() == 0:
    print('FMO 2D spectrum 197x5x197, %d members on %d GPU(s): %.1f ms (%.2f ms/member)' % (E, dist.get_world_size(), 1e3 * dt_fmo, 1e3 * dt_fmo / E))
    _, ref2d = qb.two_dimensional_spectra(model, 300, population_times=t2, ensemble_size=5)
    print('2D spectra sharded vs serial rel-L2: %.2e' % (np.linalg.norm(X2d - ref2d) / np.linalg.norm(ref2d)))
    _, ref = qb.third_order_response(model, 300, population_times=t2, ensemble_size=5)
    print('third-order sharded vs serial rel-L2: %.2e' % (np.linalg.norm(S - ref) / np.linalg.norm(ref)))
    _, ref2 = qb.simulate_dynamics(qb.RedfieldModel(systems.fmo(), hilbert_subspace='e', unit_convert=qb.CM_FS, secular=False),
                                   np.eye(7)[0], 300, ensemble_size=7)
    print('dynamics sharded vs serial rel-L2: %.2e' % (np.linalg.norm(rho - ref2) / np.linalg.norm(ref2)))
dist.destroy_process_group()
