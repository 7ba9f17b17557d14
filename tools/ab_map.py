"""A/B run of the propagator-stepping kernels on the headline workload (FMO 'ee', M = 49):
QSX_MAP_SPLIT=0 (four threads per row, shuffles), 2 (five column groups x 25 row pairs, default),
1 (five column groups, one row per thread).  Prints the stepping time per launch (CUDA events
around the propagate call with the propagators already built) and the difference of the
trajectories from variant 0."""
import os
import sys
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import qspectra_b200 as qb
    from qspectra_b200 import systems, _capi
    E = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
    model = qb.RedfieldModel(systems.fmo(), hilbert_subspace='e', unit_convert=qb.CM_FS, secular=False)
    t = np.arange(0, 1000.0, model.time_step)
    psi0 = np.eye(7)[0]
    y0 = model.density_matrix_to_state_vector(np.outer(psi0, psi0).astype(complex), 'ee')
    eom = model.ensemble_eom(E, False, 'ee', member0=0)
    y0_dev = _capi.to_device(y0).reshape(1, -1).expand(E, -1).contiguous()
    gens = np.arange(E)
    ref = None
    for variant in sys.argv[2:] or ['0', '2', '1']:
        os.environ['QSX_MAP_SPLIT'] = variant
        for _ in range(3):
            out = eom.propagate(y0_dev, t, generators=gens, return_device=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        n = 10
        for _ in range(n):
            out = eom.propagate(y0_dev, t, generators=gens, return_device=True)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        o = out.cpu().numpy()
        if ref is None:
            ref = o
        print('QSX_MAP_SPLIT=%s  %.3f ms per %d-member stepping call  max|diff vs first| %.2e  finite %s'
              % (variant, ms, E, np.abs(o - ref).max(), np.isfinite(o).all()), flush=True)


if __name__ == '__main__':
    main()
