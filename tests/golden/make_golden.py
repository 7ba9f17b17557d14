"""
Generate the committed golden fixtures by running the REAL reference
(whaley-group-berkeley/qspectra at /root/reference) in the development
container.  The reference cannot travel to the GPU box, the .npz files do.

    PYTHONPATH=/root/reference:/root/repo python tests/golden/make_golden.py

Every array here is an output of unmodified reference code; the only shim is
``inspect.getargspec`` (removed in Python 3.11, needed by
qspectra/simulate/decorators.py:30).  Integrator settings for trajectories:
rtol=1e-10, atol=1e-12, nsteps=100000 ("tight") unless the key says default.
"""
import collections
import inspect
import os
import sys
import warnings

import numpy as np

warnings.simplefilter('ignore')
if not hasattr(inspect, 'getargspec'):
    _AS = collections.namedtuple('ArgSpec', 'args varargs keywords defaults')
    inspect.getargspec = lambda f: _AS(*inspect.getfullargspec(f)[:4])

sys.path.insert(0, '/root/reference')
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import qspectra as qs                                            # noqa: E402
from qspectra.dynamics.liouville_space import liouville_subspace_index  # noqa
from qspectra.dynamics.heom import ADO_mappings                 # noqa: E402
from qspectra_b200 import systems                               # noqa: E402

TIGHT = dict(rtol=1e-10, atol=1e-12, nsteps=100000)
CM_FS = qs.CM_FS


def save(name, **arrays):
    path = os.path.join(HERE, name + '.npz')
    np.savez_compressed(path, **arrays)
    print('%-22s %7.1f KB  %d arrays' % (name, os.path.getsize(path) / 1e3,
                                         len(arrays)))


def maps():
    out = {}
    cases = [('eg,ge', 'ge', 2, 1), ('eg,fe', 'gef', 2, 1),
             ('gg,ee,ff', 'gef', 2, 1), ('gg,ge', 'ge', 1, 2),
             ('eg', 'ge', 1, 2), ('ee', 'e', 7, 1), ('fe', 'gef', 7, 1),
             ('gg,ge,eg,ee', 'gef', 7, 1), ('fg', 'gef', 7, 1),
             ('ee', 'gef', 2, 4), ('eg,fe', 'gef', 3, 2)]
    for i, (ls, hs, n, nv) in enumerate(cases):
        out['lsi_%d' % i] = liouville_subspace_index(ls, hs, n, nv)
    out['lsi_cases'] = np.array(['|'.join(map(str, c)) for c in cases])
    ado_cases = [(2, 1, 3), (2, 2, 4), (3, 0, 5), (7, 1, 4), (2, 1, 1)]
    for i, (N, K, Lc) in enumerate(ado_cases):
        ind_to_mat, mat_to_ind = ADO_mappings(N, K, Lc)
        table = np.array([m.ravel() for m in ind_to_mat], dtype=np.int64)
        up = -np.ones(table.shape, dtype=np.int64)
        down = -np.ones(table.shape, dtype=np.int64)
        for n, m in enumerate(ind_to_mat):
            for b in range(table.shape[1]):
                e = np.zeros(table.shape[1], dtype=int)
                e[b] = 1
                e = e.reshape(m.shape)
                p, q = mat_to_ind(m + e), mat_to_ind(m - e)
                up[n, b] = -1 if p is None else p
                down[n, b] = -1 if q is None else q
        out['ado_%d' % i], out['up_%d' % i], out['down_%d' % i] = table, up, down
    out['ado_cases'] = np.array(ado_cases)
    save('maps', **out)


def redfield():
    out = {}
    ham = systems.dimer(qs)
    for sec in (0, 1):
        for dic in (0, 1):
            m = qs.RedfieldModel(ham, hilbert_subspace='gef', unit_convert=CM_FS,
                                 secular=bool(sec), discard_imag_corr=bool(dic))
            out['dimer_L_sec%d_dic%d' % (sec, dic)] = m.evolution_super_operator
    m = qs.RedfieldModel(ham, hilbert_subspace='gef', unit_convert=CM_FS,
                         discard_imag_corr=True)
    f, X = qs.absorption_spectra(m, 10000)
    out['dimer_abs_f_default'], out['dimer_abs_X_default'] = f, X
    f, X = qs.absorption_spectra(m, 10000, **TIGHT)
    out['dimer_abs_f'], out['dimer_abs_X'] = f, X
    t, x = qs.linear_response(m, 'gg->eg->gg', 2000, polarization='xy',
                              exact_isotropic_average=True, **TIGHT)
    out['dimer_lin_iso_xy'] = x
    out['dimer_time_step'] = np.array(m.time_step)
    out['dimer_rw_freq'] = np.array(m.rw_freq)
    me = qs.RedfieldModel(ham, hilbert_subspace='gef', unit_convert=CM_FS,
                          evolve_basis='eigen', sparse_matrix=True)
    t, rho = qs.simulate_dynamics(me, me.hamiltonian.transform_operator_to_eigenbasis(
        np.diag([1., 0]), 'e'), 500, **TIGHT)
    out['dimer_eigen_dyn'] = rho

    fmo = systems.fmo(qs)
    m = qs.RedfieldModel(fmo, hilbert_subspace='e', unit_convert=CM_FS,
                         secular=False)
    out['fmo_L_ee'] = m.evolution_super_operator
    t, rho = qs.simulate_dynamics(m, np.eye(7)[0], 1000, **TIGHT)
    out['fmo_t'], out['fmo_rho_1ps'] = t, rho
    for n, member in enumerate(m.sample_ensemble(3)):
        out['fmo_member%d_H' % n] = member.hamiltonian.H('e')
        out['fmo_member%d_L' % n] = member.evolution_super_operator
    t, rho = qs.simulate_dynamics(m, np.eye(7)[0], 300, ensemble_size=4, **TIGHT)
    out['fmo_ens4_rho_300fs'] = rho
    ms = qs.RedfieldModel(fmo, hilbert_subspace='gef', unit_convert=CM_FS)
    t, rho = qs.simulate_dynamics(ms, np.eye(7)[0], 100000)
    out['fmo_secular_pop_100ps'] = np.einsum('ii->i', rho[-1].reshape(7, 7)).real
    out['fmo_secular_L_ee'] = ms.evolution_super_operator[np.ix_(
        ms.liouville_subspace_index('ee'), ms.liouville_subspace_index('ee'))]
    out['fmo_thermal_e'] = fmo.thermal_state('e')
    f, X = qs.absorption_spectra(ms, 2000, exact_isotropic_average=True, **TIGHT)
    out['fmo_abs_iso_f'], out['fmo_abs_iso_X'] = f, X
    f, X = qs.absorption_spectra(ms, 2000, ensemble_size=3,
                                 ensemble_random_orientations=True, **TIGHT)
    out['fmo_abs_ens3_ro_X'] = X
    save('redfield', **out)


def heom():
    out = {}
    ham = systems.dimer(qs)
    m = qs.HEOMModel(ham, hilbert_subspace='gef', unit_convert=CM_FS,
                     level_cutoff=3, low_temp_corr=False)
    for ss in ('ee', 'eg', 'fe', 'gg'):
        A = (CM_FS * m.HEOM_tensor(ss)).tocsr()
        A.sort_indices()
        out['dimer_%s_data' % ss] = A.data
        out['dimer_%s_indices' % ss] = A.indices
        out['dimer_%s_indptr' % ss] = A.indptr
    f, X = qs.absorption_spectra(m, 10000, **TIGHT)
    out['dimer_abs_f'], out['dimer_abs_X'] = f, X
    eom = m.equation_of_motion('ee')
    y0 = m.density_matrix_to_state_vector(np.diag([1., 0]).astype(complex), 'ee')
    t = np.arange(0, 500, m.time_step)
    out['dimer_dyn_t'] = t
    out['dimer_dyn'] = qs.integrate(eom, y0, t, **TIGHT)

    rng = np.random.RandomState(7)
    for tag, kw in [('k2', dict(level_cutoff=3, K=2)),
                    ('mod', dict(level_cutoff=4, K=1, modified_HEOM=True))]:
        mm = qs.HEOMModel(ham, hilbert_subspace='ge', unit_convert=CM_FS, **kw)
        for ss in ('ee', 'eg'):
            D = mm.ado_count * mm.lspace_model.liouville_subspace_index(ss).size
            y = rng.randn(D) + 1j * rng.randn(D)
            out['dimer_%s_%s_y' % (tag, ss)] = y
            out['dimer_%s_%s_Ly' % (tag, ss)] = mm.equation_of_motion(ss)(0, y)
            out['dimer_%s_%s_LTy' % (tag, ss)] = mm.equation_of_motion(ss, True)(0, y)

    vib = systems.jonas_dimer(qs)
    mv = qs.HEOMModel(vib, hilbert_subspace='ge', unit_convert=CM_FS,
                      level_cutoff=3, K=1)
    for ss in ('ee', 'eg'):
        D = mv.ado_count * mv.lspace_model.liouville_subspace_index(ss).size
        y = rng.randn(D) + 1j * rng.randn(D)
        out['vib_%s_y' % ss] = y
        out['vib_%s_Ly' % ss] = mv.equation_of_motion(ss)(0, y)

    fmo = systems.fmo(qs)
    for depth in (3, 4):
        mf = qs.HEOMModel(fmo, hilbert_subspace='e', unit_convert=CM_FS,
                          level_cutoff=depth, K=1)
        eom = mf.equation_of_motion('ee')
        D = mf.ado_count * 49
        y = np.random.RandomState(depth).randn(D) + 1j * np.random.RandomState(depth + 10).randn(D)
        out['fmo_d%d_Ly' % depth] = eom(0, y)
        y0 = mf.density_matrix_to_state_vector(
            np.diag(np.eye(7)[0]).astype(complex), 'ee')
        t = np.arange(0, 1000 if depth == 3 else 200, mf.time_step)
        traj = qs.integrate(eom, y0, t, **TIGHT)
        out['fmo_d%d_t' % depth] = t
        out['fmo_d%d_rho' % depth] = traj[:, :49]
        print('  fmo depth', depth, 'done')
    mg = qs.HEOMModel(fmo, hilbert_subspace='ge', unit_convert=CM_FS,
                      level_cutoff=2, K=1)
    f, X = qs.absorption_spectra(mg, 1000, **TIGHT)
    out['fmo_d2_abs_f'], out['fmo_d2_abs_X'] = f, X
    save('heom', **out)


def zofe():
    out = {}
    h3 = systems.fmo(qs, bath='pseudomode', n_sites=3)
    rng = np.random.RandomState(11)
    for hh in (0, 1):
        for rh in (0, 1):
            m = qs.ZOFEModel(h3, hilbert_subspace='ge', unit_convert=CM_FS,
                             ham_hermit=bool(hh), rho_hermit=bool(rh))
            D = 16 + 16 * 3 * 16
            y = rng.randn(D) + 1j * rng.randn(D)
            out['fmo3_y_%d%d' % (hh, rh)] = y
            out['fmo3_dy_%d%d' % (hh, rh)] = m.equation_of_motion('ee')(0, y)
    h7 = systems.fmo(qs, bath='pseudomode')
    m = qs.ZOFEModel(h7, hilbert_subspace='e', unit_convert=CM_FS)
    t, rho = qs.simulate_dynamics(m, np.eye(7)[0], 150, **TIGHT)
    out['fmo7_t'], out['fmo7_rho'] = t, rho
    hd = systems.dimer(qs, bath='pseudomode')
    md = qs.ZOFEModel(hd, hilbert_subspace='ge', unit_convert=CM_FS)
    f, X = qs.absorption_spectra(md, 1500, **TIGHT)
    out['dimer_abs_f'], out['dimer_abs_X'] = f, X
    save('zofe', **out)


def response():
    out = {}
    ham = systems.dimer(qs)
    red = qs.RedfieldModel(ham, hilbert_subspace='gef', unit_convert=CM_FS,
                           discard_imag_corr=True)
    t2 = np.linspace(0, 200, 3)
    for geom in ('-++', '+-+', '++-'):
        (t1, _, t3), S = qs.third_order_response(
            red, 300, population_times=t2, geometry=geom, **TIGHT)
        out['red_%s' % geom] = S
    out['t1'] = t1
    (_, _, _), S = qs.third_order_response(
        red, 300, population_times=t2, polarization='xxyy',
        exact_isotropic_average=True, **TIGHT)
    out['red_iso_xxyy'] = S
    dham = systems.dimer(qs, disorder=80)
    dred = qs.RedfieldModel(dham, hilbert_subspace='gef', unit_convert=CM_FS,
                            discard_imag_corr=True)
    (_, _, _), S = qs.third_order_response(
        dred, 300, population_times=t2, ensemble_size=3, include_signal='GSB,ESE',
        **TIGHT)
    out['red_ens3_gsb_ese'] = S
    (f1, _, f3), X = qs.two_dimensional_spectra(red, 300, population_times=t2,
                                                **TIGHT)
    out['red_2d_f1'], out['red_2d_f3'], out['red_2d'] = f1, f3, X
    hm = qs.HEOMModel(ham, hilbert_subspace='gef', unit_convert=CM_FS,
                      level_cutoff=3, low_temp_corr=False)
    (_, _, _), S = qs.third_order_response(hm, 200, population_times=t2[:2],
                                           **TIGHT)
    out['heom_-++'] = S
    pump = qs.GaussianPulse(12800, 40, scale=1e-3, freq_convert=CM_FS)
    t, st = qs.simulate_pump(red, pump, 'x', time_extra=200, **TIGHT)
    out['pump_t'], out['pump_states'] = t, st
    t, st = qs.simulate_pump(red, pump, 'x', time_extra=100,
                             exact_isotropic_average=True, **TIGHT)
    out['pump_iso_states'] = st
    f, X = qs.impulsive_probe(red, st, 500, exact_isotropic_average=True, **TIGHT)
    out['probe_f'], out['probe_X'] = f, X
    # vibronic systems
    jd = qs.RedfieldModel(systems.jonas_dimer(qs), hilbert_subspace='gef',
                          unit_convert=CM_FS, discard_imag_corr=True)
    t, rho = qs.simulate_dynamics(jd, qs.unit_vec(0, 8), 300, **TIGHT)
    out['jonas_dyn'] = rho
    f, X = qs.absorption_spectra(jd, 2000, **TIGHT)
    out['jonas_abs_f'], out['jonas_abs_X'] = f, X
    mono = qs.UnitaryModel(systems.vibronic_monomer(qs), hilbert_subspace='ge',
                           unit_convert=CM_FS)
    t, rho = qs.simulate_dynamics(mono, qs.unit_vec(0, 5), 300, **TIGHT)
    out['mono_dyn'] = rho
    f, X = qs.absorption_spectra(mono, 3000, correlation_decay_time=1000, **TIGHT)
    out['mono_abs_f'], out['mono_abs_X'] = f, X
    save('response', **out)


def round2():
    """Round-2 additions: the depth-8 ADO table (hash + sampled rows), FMO 'gef' Redfield
    generator blocks of a sampled member, FMO third-order response per pathway, and an
    eigen-basis disorder ensemble of third-order responses."""
    import hashlib
    out = {}
    # ---- BASELINE config 5: ADO_mappings(7, 1, 8), 116 280 ADOs
    ind_to_mat, mat_to_ind = ADO_mappings(7, 1, 8)
    table = np.array([m.ravel() for m in ind_to_mat], dtype=np.int64)
    out['ado8_shape'] = np.array(table.shape)
    out['ado8_sha256'] = np.array(hashlib.sha256(np.ascontiguousarray(table).tobytes()).hexdigest())
    rng = np.random.RandomState(8)
    rows = np.sort(rng.choice(len(table), 1000, replace=False))
    up = -np.ones((len(rows), table.shape[1]), dtype=np.int64)
    down = -np.ones((len(rows), table.shape[1]), dtype=np.int64)
    for i, n in enumerate(rows):
        m = ind_to_mat[n]
        for b in range(table.shape[1]):
            e = np.zeros(table.shape[1], dtype=int)
            e[b] = 1
            e = e.reshape(m.shape)
            p, q = mat_to_ind(m + e), mat_to_ind(m - e)
            up[i, b] = -1 if p is None else p
            down[i, b] = -1 if q is None else q
    out['ado8_rows'], out['ado8_index'] = rows, table[rows]
    out['ado8_up'], out['ado8_down'] = up, down
    # ---- FMO 'gef' non-secular Redfield: generator blocks of disorder member 3
    fmo = systems.fmo(qs)
    m = qs.RedfieldModel(fmo, hilbert_subspace='gef', unit_convert=CM_FS, secular=False)
    member = list(m.sample_ensemble(4))[3]
    L = member.evolution_super_operator
    for ss in ('fe', 'eg', 'ee'):
        idx = member.liouville_subspace_index(ss)
        out['fmo_gef_member3_L_%s' % ss] = L[np.ix_(idx, idx)]
    # ---- FMO third-order response, one pathway each (400 fs coherence times, 3 waiting times)
    t2 = np.array([0., 100., 300.])
    for sig in ('GSB', 'ESE', 'ESA'):
        (t1, _, t3), S = qs.third_order_response(m, 400, population_times=t2,
                                                 include_signal=sig, **TIGHT)
        out['fmo_gef_%s' % sig] = S
    out['fmo_gef_t1'], out['fmo_gef_t2'] = t1, t2
    # ---- eigen-basis model with disorder: members have their own eigenbases
    dham = systems.dimer(qs, disorder=80)
    de = qs.RedfieldModel(dham, hilbert_subspace='gef', unit_convert=CM_FS,
                          discard_imag_corr=True, evolve_basis='eigen')
    (_, _, _), S = qs.third_order_response(de, 300, population_times=np.linspace(0, 200, 3),
                                           ensemble_size=3, **TIGHT)
    out['dimer_eigen_ens3'] = S
    save('round2', **out)


def vibronic():
    """BASELINE config 5, second half: HEOM of the vibronic (Jonas) dimer with explicit modes --
    trajectory of the system density matrix from the lowest vibronic 'e' state (plain and
    modified hierarchy) and the RHS on a seeded vector in the 'e' Hilbert subspace."""
    out = {}
    vib = systems.jonas_dimer(qs)
    for tag, kw in (('plain', {}), ('mod', dict(modified_HEOM=True))):
        m = qs.HEOMModel(vib, hilbert_subspace='e', unit_convert=CM_FS, level_cutoff=5, K=1, **kw)
        eom = m.equation_of_motion('ee')
        psi = np.zeros(8, dtype=complex)
        psi[0] = 1.0
        y0 = m.density_matrix_to_state_vector(np.outer(psi, psi.conj()), 'ee')
        t = np.arange(0, 300, m.time_step)
        traj = qs.integrate(eom, y0, t, **TIGHT)
        out['vib_%s_t' % tag] = t
        out['vib_%s_rho' % tag] = np.array(traj)[:, :64]
        rng = np.random.RandomState(21)
        y = rng.randn(y0.size) + 1j * rng.randn(y0.size)
        out['vib_%s_y' % tag] = y
        out['vib_%s_Ly' % tag] = eom(0, y)
    save('vibronic', **out)


if __name__ == '__main__':
    which = sys.argv[1:] or ['maps', 'redfield', 'heom', 'zofe', 'response', 'round2', 'vibronic']
    for name in which:
        globals()[name]()
