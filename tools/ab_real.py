"""A/B run of the resident headline step (FMO 'ee', M = 49, 1 ps / 197 points): complex path
(QSX_NO_HERMITIAN_FORM=1) against the Hermitian-coordinate path for QSX_REXPM_BLOCKS = 2, 3, 4.
Per variant: ms per step (propagators rebuilt every step + stepping + member mean), the build and
stepping kernel times, and the difference of the ensemble mean from the complex path."""
import os
import sys
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import qspectra_b200 as qb
    from qspectra_b200 import systems, _capi, engine
    E = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
    model = qb.RedfieldModel(systems.fmo(), hilbert_subspace='e', unit_convert=qb.CM_FS, secular=False)
    t = np.arange(0, 1000.0, model.time_step)
    psi0 = np.eye(7)[0]
    y0 = model.density_matrix_to_state_vector(np.outer(psi0, psi0).astype(complex), 'ee')
    eom = model.ensemble_eom(E, False, 'ee', member0=0)
    y0_dev = _capi.to_device(y0).reshape(1, -1).expand(E, -1).contiguous()
    gens = np.arange(E)

    def step():
        eom.__dict__.pop('_propagators', None)
        out = eom.propagate(y0_dev, t, generators=gens, return_device=True, hermitian_state=True, packed=True)
        return engine.reduce_members(out, 1.0 / E)

    ref = None
    variants = [('complex', {'QSX_NO_HERMITIAN_FORM': '1'}), ('real (default)', {}),
                ('real, two kernels', {'QSX_HERMITIAN_TWO_KERNELS': '1'}),
                ('real, shuffle stepping', {'QSX_RMAP_SHUFFLE': '1'})]
    for name, env in variants:
        for k in ('QSX_NO_HERMITIAN_FORM', 'QSX_REXPM_BLOCKS', 'QSX_HERMITIAN_TWO_KERNELS', 'QSX_RMAP_SHUFFLE', 'QSX_RMAP_ROWS'):
            os.environ.pop(k, None)
        os.environ.update(env)
        for _ in range(3):
            mean = step()
        torch.cuda.synchronize()
        engine.PropagationStats.reset()
        engine.PropagationStats.keep_alive = True
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        n = 10
        e0.record()
        for _ in range(n):
            mean = step()
        e1.record()
        torch.cuda.synchronize()
        st = engine.PropagationStats
        st.flush()
        st.keep_alive = False
        m = mean.cpu().numpy()
        if ref is None:
            ref = m
        print('%-18s %.3f ms per step | build %.3f ms (%d products) stepping %.3f ms | rel diff vs complex %.2e'
              % (name, e0.elapsed_time(e1) / n, st.expm_ms / n, st.expm_gemms // n, st.kernel_ms / n,
                 np.linalg.norm(m - ref) / np.linalg.norm(ref)), flush=True)


if __name__ == '__main__':
    main()
