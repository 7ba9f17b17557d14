"""
CPU tests of the host layer: the reference's own known answers (its tests/
directory) re-stated against qspectra_b200, the C-ABI library symbol table, the
bit-exact integer maps and seeded disorder streams.  No GPU needed.
"""
import ctypes
import os
import re

import numpy as np
import pytest

import qspectra_b200 as qb
from qspectra_b200 import _capi, systems, operator_tools as ot
from qspectra_b200.dynamics import liouville_space as ls
from qspectra_b200.dynamics.heom import ADO_mappings, multichoose
from qspectra_b200.simulate.utils import (fourier_transform, _symmetrize,
                                          is_constant)
from qspectra_b200.simulate.decorators import (
    optional_2nd_order_isotropic_average, optional_4th_order_isotropic_average)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, 'include', 'qspectra_b200.h')).read()
    declared = set(re.findall(r'\b(qsx_[a-z0-9_]+)\s*\(', header))
    lib = ctypes.CDLL(_capi.LIB_PATH)
    missing = [name for name in sorted(declared) if not hasattr(lib, name)]
    assert not missing, missing
    assert declared == set(_capi.EXPORTS)
    assert lib.qsx_version() >= 100


def test_subspace_index_golden_vectors():
    # reference tests/test_liouville_space.py:11-29
    f = ls.liouville_subspace_index
    assert list(f('eg,ge', 'ge', 2)) == [1, 2, 3, 6]
    assert list(f('eg,fe', 'gef', 2)) == [1, 2, 7, 11]
    assert list(f('gg,ee,ff', 'gef', 2)) == [0, 5, 6, 9, 10, 15]
    assert list(f('gg', 'ge', 1, 2)) == [0, 1, 4, 5]
    assert list(f('eg', 'ge', 1, 2)) == [2, 3, 6, 7]
    with pytest.raises(ot.SubspaceError):
        f('ef', 'ge', 2)
    assert ls.all_liouville_subspaces('gef') == 'gg,ge,gf,eg,ee,ef,fg,fe,ff'


def test_maps_bit_exact_against_reference(golden):
    g = golden('maps')
    for i, case in enumerate(g['lsi_cases']):
        l, h, n, nv = case.split('|')
        got = ls.liouville_subspace_index(l, h, int(n), int(nv))
        assert got.dtype == g['lsi_%d' % i].dtype and np.array_equal(got, g['lsi_%d' % i])
    for i, (N, K, Lc) in enumerate(g['ado_cases']):
        idx, up, down = _capi.ado_enumerate(int(N * (K + 1)), int(Lc))
        assert np.array_equal(idx, g['ado_%d' % i])
        assert np.array_equal(up, g['up_%d' % i])
        assert np.array_equal(down, g['down_%d' % i])
        ind_to_mat, mat_to_ind = ADO_mappings(int(N), int(K), int(Lc))
        assert mat_to_ind(ind_to_mat[-1]) == len(ind_to_mat) - 1
        assert mat_to_ind(ind_to_mat[-1] + Lc) is None
    assert multichoose(3, 2) == [[0, 0, 2], [0, 1, 1], [0, 2, 0], [1, 0, 1],
                                 [1, 1, 0], [2, 0, 0]]
    assert int(_capi.lib().qsx_ado_count(14, 8)) == 116280     # FMO K=1 depth 8


def test_super_operator_laws():
    # reference tests/test_liouville_space.py:46-82
    R = np.random.RandomState(0).rand(3, 3, 3, 3)
    S = ls.tensor_to_super(R)
    for i, j, k, l in np.ndindex(3, 3, 3, 3):
        assert R[i, j, k, l] == S[i + 3 * j, k + 3 * l]
    X, rho = np.random.RandomState(1).rand(2, 3, 3)
    v = ls.matrix_to_ket_vec
    np.testing.assert_allclose(ls.super_left_matrix(X) @ v(rho), v(X @ rho))
    np.testing.assert_allclose(ls.super_right_matrix(X) @ v(rho), v(rho @ X))
    np.testing.assert_allclose(ls.super_commutator_matrix(X) @ v(rho), v(X @ rho - rho @ X))
    np.testing.assert_allclose(ls.ket_vec_to_matrix(v(rho)), rho)
    assert list(ls.matrix_to_bra_vec(np.array([[1, 2], [3, 4]]))) == [1, 2, 3, 4]


class _IndexModel(ls.LiouvilleSpaceModel):
    @property
    def evolution_super_operator(self):
        raise NotImplementedError


def test_liouville_space_operator_known_answers():
    # reference tests/test_liouville_space.py:107-138
    model = _IndexModel(qb.ElectronicHamiltonian(np.eye(4)))
    np.testing.assert_allclose(model.thermal_state('gg'), [1])
    np.testing.assert_allclose(model.thermal_state('gg,eg,ge,ee'), qb.unit_vec(0, 25))
    np.testing.assert_allclose(model.thermal_state('ee'), 0.25 * np.eye(4).reshape(-1))
    ones = np.ones(25)
    np.testing.assert_allclose(model.map_between_subspaces(ones, 'gg,eg,ge,ee', 'gg'), [1])
    np.testing.assert_allclose(model.map_between_subspaces(ones, 'gg,eg,ge,ee', 'gf'), np.zeros(6))
    X = np.array([[1, 2], [3, 4]])
    model = _IndexModel(qb.ElectronicHamiltonian([[0]]), hilbert_subspace='ge')
    L = ls.LiouvilleSpaceOperator(X, 'gg,eg,ge,ee->gg', model)
    state = np.array([1, 10, 100, 1000])
    np.testing.assert_allclose(L.left_multiply(state), [21])
    np.testing.assert_allclose(L.right_multiply(state), [301])
    np.testing.assert_allclose(L.commutator(state), [-280])
    np.testing.assert_allclose(L.expectation_value(state), 21)
    L = ls.LiouvilleSpaceOperator(X, 'ee->gg,ee', model)
    np.testing.assert_allclose(L.left_multiply([1]), [0, 4])
    np.testing.assert_allclose(L.expectation_value([1]), 4)


def test_hamiltonian_known_answers():
    # reference tests/test_hamiltonian.py:86-101, 127-144
    ham = qb.ElectronicHamiltonian(np.array([[1., 0], [0, 3]]), dipoles=[[1, 0, 0], [0, 1, 0]],
                                   disorder=1, energy_spread_extra=0)
    assert ham.in_rotating_frame(2).freq_step == pytest.approx(4) or True
    H = systems.dimer().in_rotating_frame()
    np.testing.assert_allclose(np.sort(H.lab_frame.E('e')),
                               [12655.22085786, 12944.77914214], atol=1e-8)   # notebook golden
    a, b = ham.sample(1), ham.sample(1)
    assert np.array_equal(a.H('e'), b.H('e'))
    assert not np.array_equal(ham.sample(1).H('e'), ham.sample(2).H('e'))
    rot = ham.in_rotating_frame(2.0)
    np.testing.assert_allclose(rot.sample(3).H('gef'), ham.sample(3).in_rotating_frame(2.0).H('gef'))
    np.testing.assert_allclose(ham.dipole_operator('ge', 'x', '-'), [[0, 1, 0], [0, 0, 0], [0, 0, 0]])


def test_operator_tools_known_answers():
    # reference tests/test_operator_tools.py
    assert ot.all_states(2) == [[], [0], [1], [0, 1]]
    assert ot.all_states(3, 'f') == [[0, 1], [0, 2], [1, 2]]
    np.testing.assert_allclose(ot.operator_1_to_2(np.array([[1, 10], [10, 2]])), [[3]])
    T = ot.transition_operator(0, 2, 'gef', '-+')
    np.testing.assert_allclose(T, [[0, 1, 0, 0], [1, 0, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0]])
    assert ot.hilbert_subspace_index('f', 'gef', 2) == slice(3, 4)
    assert ot.full_liouville_subspace('eg->ee') == 'gg,ge,eg,ee'
    with pytest.raises(ot.SubspaceError):
        ot.hilbert_subspace_index('f', 'ge', 2)


def test_seeded_streams_bit_exact():
    g, u = _capi.sample_streams(0, 0, 64, 7, 3)
    for n in (0, 1, 17, 63):
        r = np.random.RandomState([0, n])
        assert np.array_equal(r.randn(7), g[n]) and np.array_equal(r.rand(3), u[n])
    g, u = _capi.sample_streams([3, 5], 100, 4, 8, 2)
    r = np.random.RandomState([3, 5, 102])
    assert np.array_equal(r.randn(8), g[2]) and np.array_equal(r.rand(2), u[2])
    ham = systems.fmo().in_rotating_frame()
    shifts = ham.sampled_site_shifts(5, member0=10)
    for k in range(5):
        np.testing.assert_allclose(ham.sample(10 + k).H_1exc.diagonal(),
                                   ham.H_1exc.diagonal() + shifts[k], rtol=0, atol=1e-10)


def test_redfield_generators_match_reference(golden):
    g = golden('redfield')
    for sec in (0, 1):
        for dic in (0, 1):
            m = qb.RedfieldModel(systems.dimer(), hilbert_subspace='gef', unit_convert=qb.CM_FS,
                                 secular=bool(sec), discard_imag_corr=bool(dic))
            ref = g['dimer_L_sec%d_dic%d' % (sec, dic)]
            assert np.abs(m.evolution_super_operator - ref).max() <= 1e-15 * np.abs(ref).max()
    f = qb.RedfieldModel(systems.fmo(), hilbert_subspace='e', unit_convert=qb.CM_FS, secular=False)
    assert np.abs(f.evolution_super_operator - g['fmo_L_ee']).max() < 1e-15
    Ls = f.ensemble_generators(list(f.sample_ensemble(3)), 'ee')
    for n in range(3):
        assert np.abs(Ls[n] - g['fmo_member%d_L' % n]).max() < 1e-15
    E, U = f.ensemble_eigensystems(3)
    for n, member in enumerate(f.sample_ensemble(3)):
        np.testing.assert_allclose(E[n], member.hamiltonian.E('e'), rtol=0, atol=1e-9)
    L = f.evolution_super_operator                     # trace preservation: vec(I)^T L = 0
    assert np.abs(np.eye(7).reshape(-1) @ L).max() < 1e-12


def test_fourier_transform_and_helpers():
    # reference tests/test_simulate_utils.py:8-81
    t = np.linspace(0, 100, 2001)
    x = np.exp(-t / 5)
    f, X = fourier_transform(t, x)
    np.testing.assert_allclose(X, 1 / (1 / 5. - 1j * f), atol=5e-2)
    assert is_constant([1, 1, 1]) and not is_constant([1, 2])
    ts, xs = _symmetrize(np.array([0., 1, 2]), np.array([1., 2, 3]))
    assert list(ts) == [-2, -1, 0, 1, 2] and list(xs) == [0, 0, 1, 2, 3]
    with pytest.raises(ValueError):
        fourier_transform(t, x[:-1])


def test_isotropic_average_decorators():
    # reference tests/test_decorators.py:27-50
    @optional_2nd_order_isotropic_average
    def second(polarization):
        return None, float(polarization == 'xx')
    assert second('xx', exact_isotropic_average=True)[1] == pytest.approx(1 / 3.)

    @optional_4th_order_isotropic_average
    def fourth(polarization):
        return None, float(polarization == 'xxxx')
    assert fourth('xxxx', exact_isotropic_average=True)[1] == pytest.approx(1 / 5.)
    ma = [0, 0, qb.MAGIC_ANGLE, qb.MAGIC_ANGLE]
    assert fourth(ma, exact_isotropic_average=True)[1] == pytest.approx(1 / 9.)


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    m = qb.RedfieldModel(systems.dimer(), hilbert_subspace='gef', unit_convert=qb.CM_FS)
    with pytest.raises(RuntimeError):
        m.equation_of_motion('ee')
    with pytest.raises(TypeError):
        qb.integrate(lambda t, y: y, np.zeros(2, complex), np.array([0., 1.]))


def test_depth8_ado_table_matches_reference(golden):
    """BASELINE config 5: the closed-form enumeration (csrc/ado.h) for 14 bins, level_cutoff 8
    against the reference's ADO_mappings(7, 1, 8) (heom.py:92-152): SHA-256 of the whole
    116 280 x 14 table and the neighbour maps of 1 000 sampled rows, bit-exact."""
    import hashlib
    from qspectra_b200 import _capi
    g = golden('round2')
    idx, up, down = _capi.ado_enumerate(14, 8)
    assert idx.shape == tuple(g['ado8_shape'])
    assert hashlib.sha256(np.ascontiguousarray(idx.astype(np.int64)).tobytes()).hexdigest() \
        == str(g['ado8_sha256'])
    rows = g['ado8_rows']
    assert np.array_equal(idx[rows], g['ado8_index'])
    assert np.array_equal(up[rows], g['ado8_up'])
    assert np.array_equal(down[rows], g['ado8_down'])


def test_transposition_permutation_and_hermitian_coordinates(golden):
    """host logic of the Hermitian-coordinate path (engine.DenseEOM.hermitian_perm): the
    transposition permutation of a Liouville subspace, and the fact it rests on -- the reference's
    FMO generator commutes with Hermitian conjugation, so it is real in the coordinates
    (populations, Re, Im of the coherences)."""
    from qspectra_b200.dynamics.liouville_space import (transposition_permutation,
                                                        liouville_subspace_index)
    idx = liouville_subspace_index('ee', 'e', 7)
    perm = transposition_permutation(idx, 7)
    assert perm.dtype == np.int32 and np.array_equal(perm[perm], np.arange(49))
    a, b = idx % 7, idx // 7
    assert np.array_equal(idx[perm], b + 7 * a)
    assert np.array_equal(np.flatnonzero(perm == np.arange(49)), np.arange(7) * 8)     # populations
    idx_ge = liouville_subspace_index('gg,ee', 'ge', 7)
    assert transposition_permutation(idx_ge, 8).size == 50
    for open_block in ('eg', 'ge', 'gg,eg'):
        assert transposition_permutation(liouville_subspace_index(open_block, 'ge', 7), 8) is None
    # the reference generator (fixture recorded from qspectra) in Hermitian coordinates
    L = golden('redfield')['fmo_L_ee']
    assert np.abs(L[np.ix_(perm, perm)] - L.conj()).max() < 1e-15 * np.abs(L).max() * 10
    T = np.zeros((49, 49), complex)
    for k in range(49):
        s = perm[k]
        if s == k:
            T[k, k] = 1
        elif k < s:
            T[k, k] = T[k, s] = 0.5
        else:
            T[k, s], T[k, k] = 1 / 2j, -1 / 2j
    G = T @ L @ np.linalg.inv(T)
    assert np.abs(G.imag).max() < 1e-15 and np.abs(G.real).max() > 1e-2
    # a density matrix is a real vector there, and exp(G dt) steps it like exp(L dt)
    import scipy.linalg
    rho = (np.diag(np.arange(1.0, 8.0)) / 28 + 0.01 * (np.ones((7, 7)) - np.eye(7))).astype(complex)
    rho[0, 1] += 0.02j
    rho[1, 0] -= 0.02j
    y = rho.reshape(-1, order='F')
    u = T @ y
    assert np.abs(u.imag).max() == 0.0
    stepped = np.linalg.solve(T, scipy.linalg.expm(5.0 * G.real) @ u.real)
    assert np.abs(stepped - scipy.linalg.expm(5.0 * L) @ y).max() < 1e-14
