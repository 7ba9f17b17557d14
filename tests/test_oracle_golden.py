"""
Pin the ORACLE (oracle/) against (a) the reference's own golden vectors and
(b) outputs of the real reference recorded in tests/golden/*.npz by
tests/golden/make_golden.py.  CPU only.
"""
import numpy as np
import pytest

import oracle
import qspectra_b200 as qb
from qspectra_b200 import systems
from conftest import rel_l2

CM_FS = qb.CM_FS
TIGHT = oracle.TIGHT


# -- (a) golden vectors held by the reference's own tests -------------------
def test_reference_test_vectors_subspace_index():
    # reference tests/test_liouville_space.py:11-29
    assert list(oracle.subspace_index('eg,ge', 'ge', 2)) == [1, 2, 3, 6]
    assert list(oracle.subspace_index('eg,fe', 'gef', 2)) == [1, 2, 7, 11]
    assert list(oracle.subspace_index('gg,ee,ff', 'gef', 2)) == [0, 5, 6, 9, 10, 15]
    assert list(oracle.subspace_index('gg', 'ge', 1, 2)) == [0, 1, 4, 5]
    assert list(oracle.subspace_index('eg', 'ge', 1, 2)) == [2, 3, 6, 7]
    with pytest.raises(KeyError):
        oracle.subspace_index('ef', 'ge', 2)


def test_reference_test_vectors_super_operators():
    # reference tests/test_liouville_space.py:46-82
    R = np.random.RandomState(0).rand(3, 3, 3, 3)
    S = oracle.tensor_to_super_matrix(R)
    for i, j, k, l in np.ndindex(3, 3, 3, 3):
        assert R[i, j, k, l] == S[i + 3 * j, k + 3 * l]
    X = np.random.RandomState(1).rand(3, 3)
    rho = np.random.RandomState(2).rand(3, 3)
    np.testing.assert_allclose(oracle.super_left(X) @ oracle.ket_vec(rho),
                               oracle.ket_vec(X @ rho))
    np.testing.assert_allclose(oracle.super_right(X) @ oracle.ket_vec(rho),
                               oracle.ket_vec(rho @ X))
    np.testing.assert_allclose(oracle.super_commutator(X) @ oracle.ket_vec(rho),
                               oracle.ket_vec(X @ rho - rho @ X))
    assert list(oracle.ket_vec([[1, 2], [3, 4]])) == [1, 3, 2, 4]
    assert list(oracle.bra_vec([[1, 2], [3, 4]])) == [1, 2, 3, 4]


def test_reference_test_vectors_operator_known_answers():
    # reference tests/test_liouville_space.py:125-138
    idx = lambda ss: oracle.subspace_index(ss, 'ge', 1)
    X = np.array([[1, 2], [3, 4]])
    op = oracle.OracleOperator(X, 'gg,eg,ge,ee->gg', idx)
    rho = np.array([1, 10, 100, 1000])
    np.testing.assert_allclose(op.left_multiply(rho), [21])
    np.testing.assert_allclose(op.right_multiply(rho), [301])
    np.testing.assert_allclose(op.commutator(rho), [-280])
    np.testing.assert_allclose(op.expectation_value(rho), 21)
    op = oracle.OracleOperator(X, 'ee->gg,ee', idx)
    np.testing.assert_allclose(op.left_multiply([1]), [0, 4])
    np.testing.assert_allclose(op.expectation_value([1]), 4)
    op = oracle.OracleOperator(X, 'ee->gg,ge,eg,ee', idx)
    np.testing.assert_allclose(op.expectation_value([1]), 4)


def test_notebook_golden_fmo_populations():
    # examples/FMO dynamics with Redfield theory.ipynb:142-143 (default zvode
    # tolerances; printed to 8 digits)
    ham = systems.fmo()
    m = oracle.OracleRedfield(ham, hilbert_subspace='gef', unit_convert=CM_FS)
    L = m.generator('ee')
    import scipy.linalg
    y0 = np.zeros(49, complex)
    y0[0] = 1
    t_end = np.arange(0, 100000, m.time_step)[-1]
    pops = (scipy.linalg.expm(L * t_end) @ y0).reshape(7, 7).diagonal().real
    printed = [0.02394464, 0.01493232, 0.68821011, 0.22268562, 0.02374941,
               0.00193277, 0.02454513]
    np.testing.assert_allclose(pops, printed, atol=2e-6)


def test_multichoose_order():
    assert oracle.multichoose(3, 2) == [[0, 0, 2], [0, 1, 1], [0, 2, 0],
                                        [1, 0, 1], [1, 1, 0], [2, 0, 0]]


# -- (b) recorded outputs of the real reference ------------------------------
def test_maps_bit_exact(golden):
    g = golden('maps')
    for i, case in enumerate(g['lsi_cases']):
        ls, hs, n, nv = case.split('|')
        got = oracle.subspace_index(ls, hs, int(n), int(nv))
        assert got.dtype == g['lsi_%d' % i].dtype
        assert np.array_equal(got, g['lsi_%d' % i]), case
    for i, (N, K, Lc) in enumerate(g['ado_cases']):
        table = oracle.ado_table(N, K, Lc)
        assert np.array_equal(table, g['ado_%d' % i])
        up, down = oracle.ado_neighbours(table)
        assert np.array_equal(up, g['up_%d' % i])
        assert np.array_equal(down, g['down_%d' % i])


def test_redfield_generators(golden):
    g = golden('redfield')
    ham = systems.dimer()
    for sec in (0, 1):
        for dic in (0, 1):
            m = oracle.OracleRedfield(ham, hilbert_subspace='gef',
                                      unit_convert=CM_FS, secular=bool(sec),
                                      discard_imag_corr=bool(dic))
            ref = g['dimer_L_sec%d_dic%d' % (sec, dic)]
            assert np.abs(m.full_generator() - ref).max() <= 1e-15 * np.abs(ref).max()
    fmo = systems.fmo()
    m = oracle.OracleRedfield(fmo, hilbert_subspace='e', unit_convert=CM_FS,
                              secular=False)
    assert np.abs(m.full_generator() - g['fmo_L_ee']).max() <= 1e-15
    for n, member in enumerate(m.sample_ensemble(3)):
        assert np.array_equal(member.hamiltonian.H('e'), g['fmo_member%d_H' % n])
        ref = g['fmo_member%d_L' % n]
        assert np.abs(member.full_generator() - ref).max() <= 2e-15 * np.abs(ref).max()


def test_redfield_trajectories(golden):
    g = golden('redfield')
    m = oracle.OracleRedfield(systems.dimer(), hilbert_subspace='gef',
                              unit_convert=CM_FS, discard_imag_corr=True)
    f, X = oracle.absorption_spectra(m, 10000)
    assert np.array_equal(f, g['dimer_abs_f_default'])
    assert rel_l2(X, g['dimer_abs_X_default']) < 1e-13
    fmo = oracle.OracleRedfield(systems.fmo(), hilbert_subspace='e',
                                unit_convert=CM_FS, secular=False)
    t, rho = oracle.simulate_dynamics(fmo, np.eye(7)[0], 1000, **TIGHT)
    assert np.array_equal(t, g['fmo_t'])
    assert rel_l2(rho, g['fmo_rho_1ps']) < 1e-12
    t, rho = oracle.ensemble_average(
        lambda mm: oracle.simulate_dynamics(mm, np.eye(7)[0], 300, **TIGHT),
        fmo, 4)
    assert rel_l2(rho, g['fmo_ens4_rho_300fs']) < 1e-12


def test_heom_generator_and_rhs(golden):
    g = golden('heom')
    import scipy.sparse as sp
    m = oracle.OracleHEOM(systems.dimer(), hilbert_subspace='gef',
                          unit_convert=CM_FS, level_cutoff=3, low_temp_corr=False)
    for ss in ('ee', 'eg', 'fe', 'gg'):
        A = m.generator(ss)
        ref = sp.csr_matrix((g['dimer_%s_data' % ss], g['dimer_%s_indices' % ss],
                             g['dimer_%s_indptr' % ss]), shape=A.shape)
        assert abs(A - ref).max() <= 1e-15 * abs(ref).max()
    for tag, kw in [('k2', dict(level_cutoff=3, K=2)),
                    ('mod', dict(level_cutoff=4, K=1, modified_HEOM=True))]:
        mm = oracle.OracleHEOM(systems.dimer(), hilbert_subspace='ge',
                               unit_convert=CM_FS, **kw)
        for ss in ('ee', 'eg'):
            y = g['dimer_%s_%s_y' % (tag, ss)]
            assert rel_l2(mm.generator(ss) @ y, g['dimer_%s_%s_Ly' % (tag, ss)]) < 1e-14
            assert rel_l2(mm.generator(ss, True) @ y,
                          g['dimer_%s_%s_LTy' % (tag, ss)]) < 1e-14
    mv = oracle.OracleHEOM(systems.jonas_dimer(), hilbert_subspace='ge',
                           unit_convert=CM_FS, level_cutoff=3, K=1)
    for ss in ('ee', 'eg'):
        assert rel_l2(mv.generator(ss) @ g['vib_%s_y' % ss], g['vib_%s_Ly' % ss]) < 1e-14
    for depth in (3, 4):
        mf = oracle.OracleHEOM(systems.fmo(), hilbert_subspace='e',
                               unit_convert=CM_FS, level_cutoff=depth, K=1)
        D = mf.n_ado * 49
        y = (np.random.RandomState(depth).randn(D)
             + 1j * np.random.RandomState(depth + 10).randn(D))
        assert rel_l2(mf.generator('ee') @ y, g['fmo_d%d_Ly' % depth]) < 1e-14


def test_heom_vibronic_dimer(golden):
    """vibronic (Jonas) dimer HEOM with explicit modes: oracle generator against the reference's RHS"""
    g = golden('vibronic')
    for tag, kw in (('plain', {}), ('mod', dict(modified_HEOM=True))):
        mv = oracle.OracleHEOM(systems.jonas_dimer(), hilbert_subspace='e',
                               unit_convert=CM_FS, level_cutoff=5, K=1, **kw)
        assert rel_l2(mv.generator('ee') @ g['vib_%s_y' % tag], g['vib_%s_Ly' % tag]) < 1e-14


def test_heom_trajectory(golden):
    g = golden('heom')
    m = oracle.OracleHEOM(systems.dimer(), hilbert_subspace='gef',
                          unit_convert=CM_FS, level_cutoff=3, low_temp_corr=False)
    y0 = m.density_matrix_to_state_vector(np.diag([1., 0]).astype(complex), 'ee')
    traj = oracle.integrate(m.equation_of_motion('ee'), y0, g['dimer_dyn_t'], **TIGHT)
    assert rel_l2(traj, g['dimer_dyn']) < 1e-12
    f, X = oracle.absorption_spectra(m, 10000, **TIGHT)
    assert rel_l2(X, g['dimer_abs_X']) < 1e-11


def test_zofe(golden):
    g = golden('zofe')
    h3 = systems.fmo(bath='pseudomode', n_sites=3)
    for hh in (0, 1):
        for rh in (0, 1):
            m = oracle.OracleZOFE(h3, hilbert_subspace='ge', unit_convert=CM_FS,
                                  ham_hermit=bool(hh), rho_hermit=bool(rh))
            dy = m.equation_of_motion('ee')(0, g['fmo3_y_%d%d' % (hh, rh)])
            assert rel_l2(dy, g['fmo3_dy_%d%d' % (hh, rh)]) < 1e-14
    m7 = oracle.OracleZOFE(systems.fmo(bath='pseudomode'), hilbert_subspace='e',
                           unit_convert=CM_FS)
    t, rho = oracle.simulate_dynamics(m7, np.eye(7)[0], 150, **TIGHT)
    assert rel_l2(rho, g['fmo7_rho']) < 1e-11


def test_response(golden):
    g = golden('response')
    red = oracle.OracleRedfield(systems.dimer(), hilbert_subspace='gef',
                                unit_convert=CM_FS, discard_imag_corr=True)
    t2 = np.linspace(0, 200, 3)
    for geom in ('-++', '+-+', '++-'):
        (t1, _, _), S = oracle.third_order_response(
            red, 300, population_times=t2, geometry=geom, **TIGHT)
        assert rel_l2(S, g['red_%s' % geom]) < 1e-12
    _, S = oracle.iso4(lambda p: oracle.third_order_response(
        red, 300, population_times=t2, polarization=p, **TIGHT), 'xxyy')
    assert rel_l2(S, g['red_iso_xxyy']) < 1e-12
    dred = oracle.OracleRedfield(systems.dimer(disorder=80), hilbert_subspace='gef',
                                 unit_convert=CM_FS, discard_imag_corr=True)
    _, S = oracle.ensemble_average(lambda mm: oracle.third_order_response(
        mm, 300, population_times=t2, include_signal='GSB,ESE', **TIGHT), dred, 3)
    assert rel_l2(S, g['red_ens3_gsb_ese']) < 1e-12
    pump = qb.GaussianPulse(12800, 40, scale=1e-3, freq_convert=CM_FS)
    t, st = oracle.simulate_with_fields(red, [pump, pump], '-+', 'xx',
                                        time_extra=200, **TIGHT)
    assert np.array_equal(t, g['pump_t'])
    assert rel_l2(st, g['pump_states']) < 1e-11


def test_oracle_round2_fixtures(golden):
    """FMO 'gef' generator blocks of a sampled member and per-pathway third-order responses
    (reference outputs, tests/golden/make_golden.py: round2)."""
    g = golden('round2')
    ham = systems.fmo()
    m = oracle.OracleRedfield(ham, hilbert_subspace='gef', unit_convert=CM_FS, secular=False)
    member = m.__class__.__new__(m.__class__)
    member.__dict__.update(m.__dict__)
    member.hamiltonian = m.hamiltonian.sample(3)
    for ss in ('fe', 'eg', 'ee'):
        assert rel_l2(member.generator(ss), g['fmo_gef_member3_L_%s' % ss]) < 1e-12, ss
    (t1, _, _), S = oracle.third_order_response(m, 400, population_times=g['fmo_gef_t2'],
                                                include_signal='ESE', **TIGHT)
    assert np.array_equal(t1, g['fmo_gef_t1'])
    assert rel_l2(S, g['fmo_gef_ESE']) < 1e-8
