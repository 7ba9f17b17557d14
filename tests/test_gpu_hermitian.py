"""
GPU tests (B200) of the Hermitian-coordinate ("real form") propagation path, csrc/dense_real.cu:
generators that commute with Hermitian conjugation are real matrices in the coordinates
(populations, Re / Im of the coherences), so exp(L dt) and the stepping run in real arithmetic.
Checked against scipy's expm of the complex generator, against the complex device path and
against the fixtures recorded from the reference (tolerance 1e-8 relative L2, BASELINE.json
north_star; the two device paths agree to 1e-12).
"""
import ctypes as C

import numpy as np
import pytest
import scipy.linalg

import qspectra_b200 as qb
from qspectra_b200 import systems, engine, _capi
from qspectra_b200.dynamics.liouville_space import transposition_permutation
from conftest import rel_l2

pytestmark = pytest.mark.gpu
CM_FS = qb.CM_FS


def _compatible_generator(n_states, rng, scale=0.05):
    """random complex generator on the full n^2 operator space that commutes with Hermitian
    conjugation, its transposition permutation and the change of coordinates T"""
    M = n_states * n_states
    perm = transposition_permutation(np.arange(M), n_states)
    T = np.zeros((M, M), complex)
    for k in range(M):
        s = perm[k]
        if s == k:
            T[k, k] = 1
        elif k < s:
            T[k, k] = T[k, s] = 0.5
        else:
            T[k, s], T[k, k] = 1 / 2j, -1 / 2j
    G = scale * rng.randn(M, M) - 0.02 * np.eye(M)
    return np.linalg.solve(T, G @ T), perm, T


def _hermitian_state(n_states, rng):
    rho = rng.randn(n_states, n_states) + 1j * rng.randn(n_states, n_states)
    return (rho + rho.conj().T).reshape(-1, order='F')


@pytest.mark.parametrize('two_kernels', [False, True])
@pytest.mark.parametrize('n_states', [2, 3, 4, 5, 6, 7])
def test_real_form_matches_expm(n_states, two_kernels, monkeypatch):
    """default: change of coordinates fused into the propagator kernel (qsx_dense_hermitian_expm);
    QSX_HERMITIAN_TWO_KERNELS=1: qsx_dense_hermitian_form + qsx_real_expm"""
    if two_kernels:
        monkeypatch.setenv('QSX_HERMITIAN_TWO_KERNELS', '1')
    rng = np.random.RandomState(n_states)
    gens = [_compatible_generator(n_states, rng) for _ in range(3)]
    perm = gens[0][1]
    L = np.array([g[0] for g in gens])
    eom = engine.DenseEOM(L)
    eom.hermitian_perm = perm
    y0 = np.array([_hermitian_state(n_states, rng) for _ in range(3)])
    t = np.arange(0.0, 70.0, 1.0)
    engine.PropagationStats.reset()
    out = eom.propagate(y0, t, method='expm', generators=np.arange(3))
    assert eom.last['hermitian_form'] and engine.PropagationStats.hermitian_builds == 1
    for m in range(3):
        P = scipy.linalg.expm(L[m] * 1.0)
        ref = [y0[m]]
        for _ in t[1:]:
            ref.append(P @ ref[-1])
        assert rel_l2(out[m], np.array(ref)) < 1e-11
    # the complex device path on the same inputs
    eom.hermitian_perm = None
    assert rel_l2(out, eom.propagate(y0, t, method='expm', generators=np.arange(3))) < 1e-12
    engine.PropagationStats.flush()


def test_real_form_pack_unpack_roundtrip():
    torch = _capi.torch_cuda()
    rng = np.random.RandomState(5)
    perm = transposition_permutation(np.arange(49), 7)
    y = np.array([_hermitian_state(7, rng) for _ in range(17)])
    y_dev = _capi.to_device(y)
    u = torch.empty((17, 50), dtype=torch.float64, device='cuda')
    defect = torch.zeros(4, dtype=torch.float64, device='cuda')
    pptr = perm.ctypes.data_as(C.POINTER(C.c_int32))
    _capi.check(_capi.lib().qsx_hermitian_pack(y_dev.data_ptr(), 49, 17, pptr, 50, u.data_ptr(),
                                               defect.data_ptr(), _capi.current_stream_ptr()))
    back = engine.HermitianTrajectory(u, perm, 49).to_complex()
    assert np.array_equal(_capi.to_host(back), y)
    d = defect.cpu().numpy()
    assert d[2] == 0.0 and d[3] > 0.0
    assert np.all(u.cpu().numpy()[:, 49] == 0.0)
    # populations first: the real coordinates of a density matrix are Re / Im of its entries
    rho = y[0].reshape(7, 7, order='F')
    assert np.allclose(u.cpu().numpy()[0, [0, 8, 16]], rho.real[[0, 1, 2], [0, 1, 2]])


def test_incompatible_generator_falls_back():
    """a generator that does NOT commute with Hermitian conjugation is caught by the device-side
    check and propagated on the complex path"""
    rng = np.random.RandomState(11)
    L = 0.05 * (rng.randn(1, 16, 16) + 1j * rng.randn(1, 16, 16))
    eom = engine.DenseEOM(L)
    eom.hermitian_perm = transposition_permutation(np.arange(16), 4)
    y0 = _hermitian_state(4, rng)[None]
    t = np.arange(0.0, 70.0, 1.0)
    out = eom.propagate(y0, t, method='expm')
    P = scipy.linalg.expm(L[0])
    ref = [y0[0]]
    for _ in t[1:]:
        ref.append(P @ ref[-1])
    assert rel_l2(out[0], np.array(ref)) < 1e-11
    assert eom.hermitian_perm is None
    engine.PropagationStats.flush()           # the handled failure does not raise later


def test_non_hermitian_state_uses_complex_path():
    rng = np.random.RandomState(2)
    L, perm, _ = _compatible_generator(3, rng)
    eom = engine.DenseEOM(L[None])
    eom.hermitian_perm = perm
    y0 = (rng.randn(9) + 1j * rng.randn(9))[None]
    t = np.arange(0.0, 70.0, 1.0)
    out = eom.propagate(y0, t, method='expm')
    assert not eom.last.get('hermitian_form')
    P = scipy.linalg.expm(L)
    assert rel_l2(out[0, -1], np.linalg.matrix_power(P, 69) @ y0[0]) < 1e-11


def test_fmo_ensemble_real_form_vs_complex_path(monkeypatch, golden):
    model = qb.RedfieldModel(systems.fmo(), hilbert_subspace='e', unit_convert=CM_FS, secular=False)
    engine.PropagationStats.reset()
    t, rho = qb.simulate_dynamics(model, np.eye(7)[0], 300, ensemble_size=4, method_name='expm')
    assert engine.PropagationStats.hermitian_builds == 1
    assert rel_l2(rho, golden('redfield')['fmo_ens4_rho_300fs']) < 1e-8
    monkeypatch.setenv('QSX_NO_HERMITIAN_FORM', '1')
    engine.PropagationStats.reset()
    _, rho_c = qb.simulate_dynamics(model, np.eye(7)[0], 300, ensemble_size=4, method_name='expm')
    assert engine.PropagationStats.hermitian_builds == 0
    assert rel_l2(rho, rho_c) < 1e-12
    # hermiticity and trace of the averaged density matrices
    assert np.abs(rho - rho.conj().transpose(0, 2, 1)).max() < 1e-14
    assert np.abs(np.einsum('tii->t', rho) - 1).max() < 1e-13
    engine.PropagationStats.flush()


def test_gg_ee_subspace_real_form(golden):
    """M = 50 ('gg,ee' of the 'ge' model): even state dimension, ground-state population row"""
    model = qb.RedfieldModel(systems.fmo(), hilbert_subspace='ge', unit_convert=CM_FS, secular=False)
    rho0 = np.zeros((8, 8), complex)
    rho0[0, 0] = 0.25
    rho0[1, 1] = 0.5
    rho0[2, 2] = 0.25
    rho0[1, 2] = rho0[2, 1] = 0.1
    engine.PropagationStats.reset()
    t, rho = qb.simulate_dynamics(model, rho0, 400, liouville_subspace='gg,ee', save_func=lambda y: y)
    assert engine.PropagationStats.hermitian_builds == 1
    eom = model.equation_of_motion('gg,ee')
    perm, eom.hermitian_perm = eom.hermitian_perm, None
    try:
        _, rho_c = qb.simulate_dynamics(model, rho0, 400, liouville_subspace='gg,ee', save_func=lambda y: y)
    finally:
        eom.hermitian_perm = perm
    assert rho.shape == (len(t), 50)
    assert rel_l2(rho, rho_c) < 1e-12
    engine.PropagationStats.flush()


def test_headline_size_properties(monkeypatch):
    """BASELINE configs[1] at a quarter of its size (2500 members x 197 points, enough for several
    rounds of every persistent CTA): the Hermitian-coordinate path against the complex path on the
    same resident generators, and the size-independent properties of the mean -- unit trace,
    Hermiticity, positivity of the populations, member linearity (mean of two halves)."""
    model = qb.RedfieldModel(systems.fmo(), hilbert_subspace='e', unit_convert=CM_FS, secular=False)
    E = 2500
    t = np.arange(0, 1000.0, model.time_step)
    psi0 = np.eye(7)[0]
    y0 = model.density_matrix_to_state_vector(np.outer(psi0, psi0).astype(complex), 'ee')
    eom = model.ensemble_eom(E, False, 'ee', member0=0)
    y0_dev = _capi.to_device(y0).reshape(1, -1).expand(E, -1).contiguous()
    gens = np.arange(E)
    out = eom.propagate(y0_dev, t, generators=gens, return_device=True, hermitian_state=True, packed=True)
    assert isinstance(out, engine.HermitianTrajectory)
    mean = _capi.to_host(engine.reduce_members(out, 1.0 / E))
    per_member = _capi.to_host(out.to_complex())
    monkeypatch.setenv('QSX_NO_HERMITIAN_FORM', '1')
    eom.__dict__.pop('_propagators', None)
    ref = _capi.to_host(eom.propagate(y0_dev, t, generators=gens, return_device=True))
    assert rel_l2(per_member, ref) < 1e-12
    assert rel_l2(mean, ref.mean(axis=0)) < 1e-12
    rho = mean.reshape(len(t), 7, 7, order='F')
    assert np.abs(np.einsum('tii->t', rho) - 1).max() < 1e-13
    assert np.abs(rho - rho.conj().transpose(0, 2, 1)).max() < 1e-15
    assert np.einsum('tii->ti', rho).real.min() > -1e-12
    halves = 0.5 * (per_member[:E // 2].mean(axis=0) + per_member[E // 2:].mean(axis=0))
    assert rel_l2(mean, halves) < 1e-13
    engine.PropagationStats.flush()


def test_captured_ensemble_step_matches_eager_path():
    """the resident ensemble step as a CUDA graph (engine.CapturedEnsembleStep) against the eager
    calls it captures; replays rebuild the propagators and return the same mean"""
    model = qb.RedfieldModel(systems.fmo(), hilbert_subspace='e', unit_convert=CM_FS, secular=False)
    E = 300
    t = np.arange(0, 1000.0, model.time_step)
    psi0 = np.eye(7)[0]
    y0 = model.density_matrix_to_state_vector(np.outer(psi0, psi0).astype(complex), 'ee')
    eom = model.ensemble_eom(E, False, 'ee', member0=0)
    y0_dev = _capi.to_device(y0).reshape(1, -1).expand(E, -1).contiguous()
    out = eom.propagate(y0_dev, t, generators=np.arange(E), return_device=True, hermitian_state=True, packed=True)
    eager = _capi.to_host(engine.reduce_members(out, 1.0 / E))
    step = engine.CapturedEnsembleStep(eom, y0_dev, t, 1.0 / E)
    first = _capi.to_host(step.run()).copy()
    assert step.verify() == 6 * E or step.verify() >= 5 * E      # 5 products + squarings per member
    assert np.abs(first - eager).max() <= 1e-15
    step.P.zero_()                                                  # a replay rebuilds the propagators
    second = _capi.to_host(step.run())
    assert np.array_equal(first, second)
    engine.PropagationStats.flush()
