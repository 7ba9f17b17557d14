import os, sys, time, numpy as np
sys.path.insert(0, '/root/repo')
import torch, torch.distributed as dist
import qspectra_b200 as qb
from qspectra_b200 import systems, engine, _capi
rank = int(os.environ['RANK']); world = int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
dist.init_process_group('nccl')
E = 10000
model = qb.RedfieldModel(systems.fmo(), hilbert_subspace='e', unit_convert=qb.CM_FS, secular=False)
t = np.arange(0, 1000.0, model.time_step)
psi0 = np.eye(7)[0]
y0 = model.density_matrix_to_state_vector(np.outer(psi0, psi0).astype(complex), 'ee')
eom = model.ensemble_eom(E, False, 'ee', member0=rank * E)
y0_dev = _capi.to_device(y0).reshape(1, -1).expand(E, -1).contiguous()
gens = np.arange(E)
def local_step():
    eom.__dict__.pop('_propagators', None)
    out = eom.propagate(y0_dev, t, generators=gens, return_device=True)
    return engine.reduce_members(out, 1.0 / (E * world))
for mode in ('noreduce', 'reduce', 'reduce_keepalive', 'reduce_sync'):
    engine.PropagationStats.flush()
    engine.PropagationStats.keep_alive = mode == 'reduce_keepalive'
    for _ in range(3):
        r = local_step()
    torch.cuda.synchronize(); dist.barrier()
    evs = [torch.cuda.Event(True) for _ in range(11)]
    t0 = time.perf_counter()
    evs[0].record()
    for i in range(10):
        r = local_step()
        if mode != 'noreduce':
            dist.reduce(torch.view_as_real(r), dst=0)
        if mode == 'reduce_sync':
            torch.cuda.synchronize()
        evs[i + 1].record()
    host = time.perf_counter() - t0
    torch.cuda.synchronize()
    print(rank, mode, 'host enqueue %.1f ms' % (1e3 * host), ['%.1f' % evs[i].elapsed_time(evs[i + 1]) for i in range(10)], 'mem %.1f GB' % (torch.cuda.max_memory_allocated() / 1e9), flush=True)
    engine.PropagationStats.flush()
dist.destroy_process_group()
