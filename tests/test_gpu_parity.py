"""
GPU parity tests (B200): the CUDA path, called through the C ABI, against the
golden fixtures recorded from the real reference and against the CPU oracle on
the same seeded inputs.  Tolerance: 1e-8 relative L2 (BASELINE.json north_star)
for trajectories/spectra; integer maps bit-exact.
"""
import numpy as np
import pytest

import oracle
import qspectra_b200 as qb
from qspectra_b200 import systems, engine
from conftest import rel_l2

pytestmark = pytest.mark.gpu
CM_FS = qb.CM_FS
TOL = 1e-8
TIGHT = oracle.TIGHT


@pytest.fixture(scope='module')
def fmo_model():
    return qb.RedfieldModel(systems.fmo(), hilbert_subspace='e',
                            unit_convert=CM_FS, secular=False)


# ------------------------------------------------------------------ K1 dense
def test_dense_apply_matches_generator(fmo_model, golden):
    L = golden('redfield')['fmo_L_ee']
    rng = np.random.RandomState(0)
    y = rng.randn(5, 49) + 1j * rng.randn(5, 49)
    eom = fmo_model.equation_of_motion('ee')
    assert rel_l2(eom.apply(y), y @ L.T) < 1e-14
    assert rel_l2(eom(0.0, y[0]), L @ y[0]) < 1e-14
    eom_h = fmo_model.equation_of_motion('ee', heisenberg_picture=True)
    assert rel_l2(eom_h.apply(y), y @ L) < 1e-14


@pytest.mark.parametrize('method,kw', [('taylor', {}), ('zvode', {}),
                                       ('rk4', dict(rk4_substeps=24)),
                                       ('dopri5', dict(rtol=1e-11, atol=1e-13))])
def test_fmo_redfield_trajectory(fmo_model, golden, method, kw):
    g = golden('redfield')
    t, rho = qb.simulate_dynamics(fmo_model, np.eye(7)[0], 1000,
                                  method_name=method, **kw)
    assert np.array_equal(t, g['fmo_t'])
    assert rel_l2(rho, g['fmo_rho_1ps']) < TOL


def test_fmo_redfield_ensemble(fmo_model, golden):
    t, rho = qb.simulate_dynamics(fmo_model, np.eye(7)[0], 300, ensemble_size=4)
    assert rel_l2(rho, golden('redfield')['fmo_ens4_rho_300fs']) < TOL


def test_batched_columns_share_generator(fmo_model):
    """columns of one generator are propagated NB at a time: each must equal
    its own single-column run (linearity / batching invariance)."""
    rng = np.random.RandomState(3)
    y0 = rng.randn(11, 49) + 1j * rng.randn(11, 49)
    eom = fmo_model.equation_of_motion('ee')
    t = np.arange(0, 100, fmo_model.time_step)
    batch = eom.propagate(y0, t)
    for j in (0, 7, 10):
        single = eom.propagate(y0[j:j + 1], t)
        assert rel_l2(batch[j], single[0]) < 1e-12
    lin = eom.propagate((2 * y0[0] - 1j * y0[1])[None], t)
    assert rel_l2(lin[0], 2 * batch[0] - 1j * batch[1]) < 1e-11


def test_trace_preservation_100ps(fmo_model):
    """size-independent property at BASELINE length: tr rho(t) == 1 over the
    full 19 601-point, 100 ps grid."""
    t, rho = qb.simulate_dynamics(fmo_model, np.eye(7)[0], 100000)
    assert len(t) == 19601
    tr = np.einsum('tii->t', rho)
    assert np.abs(tr - 1).max() < 1e-9


def test_dimer_absorption(golden):
    g = golden('redfield')
    m = qb.RedfieldModel(systems.dimer(), hilbert_subspace='gef',
                         unit_convert=CM_FS, discard_imag_corr=True)
    f, X = qb.absorption_spectra(m, 10000)
    assert np.array_equal(f, g['dimer_abs_f'])
    assert rel_l2(X, g['dimer_abs_X']) < TOL
    t, x = qb.linear_response(m, 'gg->eg->gg', 2000, polarization='xy',
                              exact_isotropic_average=True)
    assert rel_l2(x, g['dimer_lin_iso_xy']) < TOL


def test_fmo_absorption_variants(golden):
    g = golden('redfield')
    ms = qb.RedfieldModel(systems.fmo(), hilbert_subspace='gef',
                          unit_convert=CM_FS)
    f, X = qb.absorption_spectra(ms, 2000, exact_isotropic_average=True)
    assert rel_l2(X, g['fmo_abs_iso_X']) < TOL
    f, X = qb.absorption_spectra(ms, 2000, ensemble_size=3,
                                 ensemble_random_orientations=True)
    assert rel_l2(X, g['fmo_abs_ens3_ro_X']) < TOL


def test_eigenbasis_model(golden):
    m = qb.RedfieldModel(systems.dimer(), hilbert_subspace='gef',
                         unit_convert=CM_FS, evolve_basis='eigen',
                         sparse_matrix=True)
    rho0 = m.hamiltonian.transform_operator_to_eigenbasis(np.diag([1., 0]), 'e')
    t, rho = qb.simulate_dynamics(m, rho0, 500)
    assert rel_l2(rho, golden('redfield')['dimer_eigen_dyn']) < TOL


# ------------------------------------------------------------------- K2 HEOM
def test_heom_index_maps_bit_exact(golden):
    g = golden('maps')
    for i, (N, K, Lc) in enumerate(g['ado_cases']):
        ham = systems.fmo(n_sites=int(N))
        m = qb.HEOMModel(ham, hilbert_subspace='e', unit_convert=CM_FS,
                         level_cutoff=int(Lc), K=int(K))
        idx, up, down = m.equation_of_motion('ee').index_maps()
        assert np.array_equal(idx, g['ado_%d' % i])
        assert np.array_equal(up, g['up_%d' % i])
        assert np.array_equal(down, g['down_%d' % i])


def test_heom_apply_vs_reference_rhs(golden):
    g = golden('heom')
    for tag, kw in [('k2', dict(level_cutoff=3, K=2)),
                    ('mod', dict(level_cutoff=4, K=1, modified_HEOM=True))]:
        m = qb.HEOMModel(systems.dimer(), hilbert_subspace='ge',
                         unit_convert=CM_FS, **kw)
        for ss in ('ee', 'eg'):
            y = g['dimer_%s_%s_y' % (tag, ss)]
            assert rel_l2(m.equation_of_motion(ss)(0, y),
                          g['dimer_%s_%s_Ly' % (tag, ss)]) < 1e-13
            assert rel_l2(m.equation_of_motion(ss, True)(0, y),
                          g['dimer_%s_%s_LTy' % (tag, ss)]) < 1e-13
    mv = qb.HEOMModel(systems.jonas_dimer(), hilbert_subspace='ge',
                      unit_convert=CM_FS, level_cutoff=3, K=1)
    for ss in ('ee', 'eg'):
        assert rel_l2(mv.equation_of_motion(ss)(0, g['vib_%s_y' % ss]),
                      g['vib_%s_Ly' % ss]) < 1e-13
    for depth in (3, 4):
        mf = qb.HEOMModel(systems.fmo(), hilbert_subspace='e',
                          unit_convert=CM_FS, level_cutoff=depth, K=1)
        D = mf.ado_count * 49
        y = (np.random.RandomState(depth).randn(D)
             + 1j * np.random.RandomState(depth + 10).randn(D))
        assert rel_l2(mf.equation_of_motion('ee')(0, y),
                      g['fmo_d%d_Ly' % depth]) < 1e-13


def test_heom_apply_vs_oracle_csr_subspaces():
    """every Liouville subspace the response functions use, both pictures"""
    ham = systems.fmo(n_sites=3)
    m = qb.HEOMModel(ham, hilbert_subspace='gef', unit_convert=CM_FS,
                     level_cutoff=3, K=1)
    o = oracle.OracleHEOM(ham, hilbert_subspace='gef', unit_convert=CM_FS,
                          level_cutoff=3, K=1)
    rng = np.random.RandomState(5)
    for ss in ('gg', 'eg', 'ge', 'ee', 'fe', 'fg', 'gg,ge,eg,ee'):
        for heis in (False, True):
            A = o.generator(ss, heis)
            y = rng.randn(A.shape[0]) + 1j * rng.randn(A.shape[0])
            got = m.equation_of_motion(ss, heis)(0, y)
            assert rel_l2(got, A @ y) < 1e-13, (ss, heis)


@pytest.mark.parametrize('method,kw', [('taylor', {}), ('rk4', dict(rk4_substeps=40))])
def test_heom_trajectories(golden, method, kw):
    g = golden('heom')
    m = qb.HEOMModel(systems.dimer(), hilbert_subspace='gef', unit_convert=CM_FS,
                     level_cutoff=3, low_temp_corr=False)
    y0 = m.density_matrix_to_state_vector(np.diag([1., 0]).astype(complex), 'ee')
    traj = qb.integrate(m.equation_of_motion('ee'), y0, g['dimer_dyn_t'],
                        method_name=method, **kw)
    assert rel_l2(traj, g['dimer_dyn']) < TOL
    mf = qb.HEOMModel(systems.fmo(), hilbert_subspace='e', unit_convert=CM_FS,
                      level_cutoff=3, K=1)
    t, rho = qb.simulate_dynamics(mf, np.eye(7)[0], 1000, method_name=method, **kw)
    assert np.array_equal(t, g['fmo_d3_t'])
    assert rel_l2(rho, g['fmo_d3_rho'].reshape(-1, 7, 7).transpose(0, 2, 1)) < TOL


def test_heom_depth4_and_trace(golden):
    g = golden('heom')
    mf = qb.HEOMModel(systems.fmo(), hilbert_subspace='e', unit_convert=CM_FS,
                      level_cutoff=4, K=1)
    t, rho = qb.simulate_dynamics(mf, np.eye(7)[0], 200)
    assert rel_l2(rho, g['fmo_d4_rho'].reshape(-1, 7, 7).transpose(0, 2, 1)) < TOL
    assert np.abs(np.einsum('tii->t', rho) - 1).max() < 1e-10


def test_heom_absorption(golden):
    g = golden('heom')
    m = qb.HEOMModel(systems.dimer(), hilbert_subspace='gef', unit_convert=CM_FS,
                     level_cutoff=3, low_temp_corr=False)
    f, X = qb.absorption_spectra(m, 10000)
    assert rel_l2(X, g['dimer_abs_X']) < TOL
    mg = qb.HEOMModel(systems.fmo(), hilbert_subspace='ge', unit_convert=CM_FS,
                      level_cutoff=2, K=1)
    f, X = qb.absorption_spectra(mg, 1000)
    assert rel_l2(X, g['fmo_d2_abs_X']) < TOL


def test_heom_ensemble_members_match_single_runs():
    ham = systems.fmo(n_sites=4)
    m = qb.HEOMModel(ham, hilbert_subspace='e', unit_convert=CM_FS,
                     level_cutoff=3, K=1)
    t, avg = qb.simulate_dynamics(m, np.eye(4)[0], 150, ensemble_size=3)
    singles = [qb.simulate_dynamics(mm, np.eye(4)[0], 150)[1]
               for mm in m.sample_ensemble(3)]
    assert rel_l2(avg, np.mean(singles, axis=0)) < 1e-12


def test_heom_lean_tile_ensemble_matches_single_member_launches():
    """FMO depth 4, 128 disorder members in one launch (2816 staged tiles: the row tile with one
    Hamiltonian staged per tile) against launches of one member each (batch tile)."""
    import torch
    E = 128
    m = qb.HEOMModel(systems.fmo(), hilbert_subspace='e', unit_convert=CM_FS,
                     level_cutoff=4, K=1)
    eom = m.ensemble_eom(E, False, 'ee')
    y0 = m.density_matrix_to_state_vector(np.diag(np.eye(7)[0]).astype(complex), 'ee')
    y0 = torch.from_numpy(y0).cuda().reshape(1, -1).expand(E, -1).contiguous()
    t = m.time_step * np.arange(6)
    for _ in range(2):      # twice: the race was timing dependent
        out = eom.propagate(y0, t, save=('ado0',), generators=np.arange(E),
                            return_device=True).cpu().numpy()
        for e in range(0, E, 16):
            one = eom.propagate(y0[e:e + 1].contiguous(), t, save=('ado0',),
                                generators=np.array([e]), return_device=True).cpu().numpy()[0]
            assert rel_l2(out[e], one) < 1e-13


@pytest.mark.parametrize('modified', [False, True])
@pytest.mark.parametrize('grid', [None, '2', '3', '5'])
def test_heom_row_tile_matches_batch_tile(modified, grid, golden, monkeypatch):
    """The row tile (csrc/heom_row.cuh; default from 64 (column, tile) units on) forced onto the depth-4
    FMO hierarchy: RHS application, adaptive Taylor and product-form trajectories against the
    batch tile that the golden-fixture tests validate, and against the reference trajectory.
    grid = 2: two CTAs walk eleven tiles each, so the buffer ring wraps and the mbarrier
    phases flip several times per stage; grid = 3 and 5 leave a short last round (22 = 7 x 3 + 1
    = 4 x 5 + 2), so the unit -> CTA assignment of the barrier-free stages rotates from stage
    to stage and CTAs own different numbers of tiles."""
    import torch
    m = qb.HEOMModel(systems.fmo(), hilbert_subspace='e', unit_convert=CM_FS, level_cutoff=4,
                     K=1, modified_HEOM=modified, low_temp_corr=True)
    y0 = m.density_matrix_to_state_vector(np.diag(np.eye(7)[0]).astype(complex), 'ee')
    g = golden('heom')
    t = g['fmo_d3_t'][:40]
    rng = np.random.RandomState(3)
    eom = m.equation_of_motion('ee')
    y = rng.randn(2, eom.dim) + 1j * rng.randn(2, eom.dim)
    if grid:
        monkeypatch.setenv('QSX_HEOM_GRID', grid)
    monkeypatch.setenv('QSX_HEOM_REPILOT', '5')
    y0d = torch.from_numpy(y0).cuda().reshape(1, -1)
    res = {}
    for variant, method in (('b', 'taylor'), ('r', 'taylor'), ('r', 'poly')):
        monkeypatch.setenv('QSX_HEOM_VARIANT', variant)
        traj = eom.propagate(y0d, t, save=('ado0',), method=method,
                             return_device=True).cpu().numpy()
        res[variant, method] = (traj, np.asarray(eom.apply(y)), dict(eom.last))
    ref = res['b', 'taylor']
    assert rel_l2(res['r', 'taylor'][1], ref[1]) < 1e-13
    assert rel_l2(res['r', 'taylor'][0], ref[0]) < 1e-12
    assert rel_l2(res['r', 'poly'][0], ref[0]) < 1e-11
    # the product form really ran: fewer accumulator passes do not show in rhs counts, but the
    # pilot schedule does (39 intervals, a Taylor pilot every 6th)
    assert res['r', 'poly'][2]['rhs'] > 0
    if not modified:
        want = g['fmo_d4_rho'].reshape(-1, 49)[:40]       # raw state-vector layout, ADO 0
        assert rel_l2(res['r', 'poly'][0][0], want) < TOL
    # full-state and per-ADO matrix saves go through the sigma -> rho rescaling
    monkeypatch.setenv('QSX_HEOM_VARIANT', 'r')
    full_r = eom.propagate(y0d, t[:4], method='poly', return_device=True).cpu().numpy()
    S = rng.randn(3, 49) + 1j * rng.randn(3, 49)
    mat_r = eom.propagate(y0d, t[:4], method='poly', save=S, return_device=True).cpu().numpy()
    monkeypatch.setenv('QSX_HEOM_VARIANT', 'b')
    full_b = eom.propagate(y0d, t[:4], method='taylor', return_device=True).cpu().numpy()
    mat_b = eom.propagate(y0d, t[:4], method='taylor', save=S, return_device=True).cpu().numpy()
    assert rel_l2(full_r, full_b) < 1e-11
    assert rel_l2(mat_r, mat_b) < 1e-11


@pytest.mark.parametrize('grid', [None, '4', '7'])
def test_heom_row_tile_ensemble(grid, monkeypatch):
    """Members with their own Hamiltonian (H staged with every tile) on the row tile, forced
    onto a small batch, against the batch tile.  With a capped grid the barrier-free stages
    run several rounds per stage over six columns (dependencies stay inside a column)."""
    import torch
    E = 6
    if grid:
        monkeypatch.setenv('QSX_HEOM_GRID', grid)
    m = qb.HEOMModel(systems.fmo(), hilbert_subspace='e', unit_convert=CM_FS,
                     level_cutoff=3, K=1)
    eom = m.ensemble_eom(E, False, 'ee')
    y0 = m.density_matrix_to_state_vector(np.diag(np.eye(7)[0]).astype(complex), 'ee')
    y0 = torch.from_numpy(y0).cuda().reshape(1, -1).expand(E, -1).contiguous()
    t = m.time_step * np.arange(6)
    out = {}
    for variant, method in (('b', 'taylor'), ('r', 'poly')):
        monkeypatch.setenv('QSX_HEOM_VARIANT', variant)
        out[variant] = eom.propagate(y0, t, save=('ado0',), generators=np.arange(E),
                                     method=method, return_device=True).cpu().numpy()
    assert rel_l2(out['r'], out['b']) < 1e-11


@pytest.mark.parametrize('modified', [False, True])
def test_heom_vibronic_row_tile(modified, golden, monkeypatch):
    """BASELINE config 5, second half: vibronic dimer with explicit modes (2 sites x 4 vibrational
    states = 8 states in 'e', bins = 2 sites x 2 exponentials).  The shaped row tile
    Cfg<8, 2, 4> (states grouped per site) against the table-driven generic tile: RHS on a
    seeded vector, product-form and adaptive-Taylor trajectories, a 3-member batch; the RHS is
    also pinned to the reference through test_heom_apply_vs_reference_rhs."""
    import torch
    m = qb.HEOMModel(systems.jonas_dimer(), hilbert_subspace='e', unit_convert=CM_FS,
                     level_cutoff=6, K=1, modified_HEOM=modified)
    eom = m.equation_of_motion('ee')
    rng = np.random.RandomState(11)
    y = rng.randn(2, eom.dim) + 1j * rng.randn(2, eom.dim)
    psi = np.zeros(8, dtype=complex)
    psi[0] = 1.0
    y0 = m.density_matrix_to_state_vector(np.outer(psi, psi.conj()), 'ee')
    y0d = torch.from_numpy(y0).cuda().reshape(1, -1).expand(3, -1).contiguous()
    t = m.time_step * np.arange(12)
    monkeypatch.setenv('QSX_HEOM_REPILOT', '4')
    res = {}
    for variant, method in ((' ', 'poly'), (' ', 'taylor'), ('g', 'taylor')):
        if variant == ' ':
            monkeypatch.delenv('QSX_HEOM_VARIANT', raising=False)
        else:
            monkeypatch.setenv('QSX_HEOM_VARIANT', variant)
        traj = eom.propagate(y0d, t, save=('ado0',), method=method,
                             return_device=True).cpu().numpy()
        full = eom.propagate(y0d[:1], t[:3], method=method, return_device=True).cpu().numpy()
        res[variant, method] = (np.asarray(eom.apply(y)), traj, full)
    ref = res['g', 'taylor']
    for key in ((' ', 'poly'), (' ', 'taylor')):
        assert rel_l2(res[key][0], ref[0]) < 1e-13, key
        assert rel_l2(res[key][1], ref[1]) < 1e-11, key
        assert rel_l2(res[key][2], ref[2]) < 1e-11, key
    rho = res[' ', 'poly'][1][0].reshape(-1, 8, 8)
    assert np.abs(np.einsum('tii->t', rho) - 1).max() < 1e-12
    # the reference itself: RHS and a 300 fs trajectory at level_cutoff 5
    tag = 'mod' if modified else 'plain'
    g = golden('vibronic')
    monkeypatch.delenv('QSX_HEOM_VARIANT', raising=False)
    m5 = qb.HEOMModel(systems.jonas_dimer(), hilbert_subspace='e', unit_convert=CM_FS,
                      level_cutoff=5, K=1, modified_HEOM=modified)
    eom5 = m5.equation_of_motion('ee')
    assert rel_l2(eom5(0, g['vib_%s_y' % tag]), g['vib_%s_Ly' % tag]) < 1e-13
    for method in ('poly', 'taylor'):
        traj = eom5.propagate(torch.from_numpy(y0[:eom5.dim].copy()).cuda().reshape(1, -1),
                              g['vib_%s_t' % tag], save=('ado0',), method=method,
                              return_device=True).cpu().numpy()[0]
        assert rel_l2(traj, g['vib_%s_rho' % tag]) < TOL, method


def test_heom_depth8_default_tile(golden, monkeypatch):
    """BASELINE config 5 (FMO, level_cutoff 8: 116 280 ADOs, 3 634 tiles): the tile and the
    integrator the bench times (row tile, product form) against the batch tile with the adaptive
    Taylor series -- RHS on a seeded random vector, a 3-interval trajectory from site 1, trace
    conservation, and closeness to the converged depth-4 reference trajectory."""
    import torch
    m = qb.HEOMModel(systems.fmo(), hilbert_subspace='e', unit_convert=CM_FS, level_cutoff=8,
                     K=1)
    eom = m.equation_of_motion('ee')
    assert eom.n_ado == 116280
    rng = np.random.RandomState(8)
    y = torch.from_numpy(rng.randn(1, eom.dim) + 1j * rng.randn(1, eom.dim)).cuda()
    y0 = m.density_matrix_to_state_vector(np.diag(np.eye(7)[0]).astype(complex), 'ee')
    y0d = torch.from_numpy(y0).cuda().reshape(1, -1)
    t = m.time_step * np.arange(4)
    res = {}
    for variant, method in ((' ', 'poly'), ('b', 'taylor')):
        if variant == ' ':
            monkeypatch.delenv('QSX_HEOM_VARIANT', raising=False)
        else:
            monkeypatch.setenv('QSX_HEOM_VARIANT', variant)
        dy = torch.empty_like(y)
        eom._apply_dev(y, dy, 1, None)
        traj = eom.propagate(y0d, t, save=('ado0',), method=method,
                             return_device=True).cpu().numpy()[0]
        res[variant] = (dy.cpu().numpy(), traj)
    assert rel_l2(res[' '][0], res['b'][0]) < 1e-13
    assert rel_l2(res[' '][1], res['b'][1]) < 1e-11
    rho = res[' '][1].reshape(-1, 7, 7)
    assert np.abs(np.einsum('tii->t', rho) - 1).max() < 1e-12
    d4 = golden('heom')['fmo_d4_rho'].reshape(-1, 49)[:4]      # raw state-vector layout, ADO 0
    assert rel_l2(res[' '][1], d4) < 2e-3       # hierarchy depth 4 is converged to ~1e-3 on this span


# ------------------------------------------------------------------ response
def test_third_order_response_redfield(golden):
    g = golden('response')
    red = qb.RedfieldModel(systems.dimer(), hilbert_subspace='gef',
                           unit_convert=CM_FS, discard_imag_corr=True)
    t2 = np.linspace(0, 200, 3)
    for geom in ('-++', '+-+', '++-'):
        (t1, _, _), S = qb.third_order_response(red, 300, population_times=t2,
                                                geometry=geom)
        assert np.array_equal(t1, g['t1'])
        assert rel_l2(S, g['red_%s' % geom]) < TOL
    _, S = qb.third_order_response(red, 300, population_times=t2,
                                   polarization='xxyy',
                                   exact_isotropic_average=True)
    assert rel_l2(S, g['red_iso_xxyy']) < TOL
    dred = qb.RedfieldModel(systems.dimer(disorder=80), hilbert_subspace='gef',
                            unit_convert=CM_FS, discard_imag_corr=True)
    _, S = qb.third_order_response(dred, 300, population_times=t2,
                                   ensemble_size=3, include_signal='GSB,ESE')
    assert rel_l2(S, g['red_ens3_gsb_ese']) < TOL
    (f1, _, f3), X = qb.two_dimensional_spectra(red, 300, population_times=t2)
    assert np.allclose(f1, g['red_2d_f1']) and np.allclose(f3, g['red_2d_f3'])
    assert rel_l2(X, g['red_2d']) < TOL


def test_third_order_response_heom(golden):
    g = golden('response')
    hm = qb.HEOMModel(systems.dimer(), hilbert_subspace='gef', unit_convert=CM_FS,
                      level_cutoff=3, low_temp_corr=False)
    _, S = qb.third_order_response(hm, 200, population_times=np.linspace(0, 200, 3)[:2])
    assert rel_l2(S, g['heom_-++']) < TOL


def test_pump_probe(golden):
    g = golden('response')
    red = qb.RedfieldModel(systems.dimer(), hilbert_subspace='gef',
                           unit_convert=CM_FS, discard_imag_corr=True)
    pump = qb.GaussianPulse(12800, 40, scale=1e-3, freq_convert=CM_FS)
    t, st = qb.simulate_pump(red, pump, 'x', time_extra=200, rtol=1e-11, atol=1e-14)
    assert np.array_equal(t, g['pump_t'])
    assert rel_l2(st, g['pump_states']) < TOL
    t, st = qb.simulate_pump(red, pump, 'x', time_extra=100,
                             exact_isotropic_average=True, rtol=1e-11, atol=1e-14)
    assert rel_l2(st, g['pump_iso_states']) < TOL
    f, X = qb.impulsive_probe(red, g['pump_iso_states'], 500,
                              exact_isotropic_average=True)
    assert rel_l2(X, g['probe_X']) < TOL


def test_vibronic_systems(golden):
    g = golden('response')
    jd = qb.RedfieldModel(systems.jonas_dimer(), hilbert_subspace='gef',
                          unit_convert=CM_FS, discard_imag_corr=True)
    t, rho = qb.simulate_dynamics(jd, qb.unit_vec(0, 8), 300)
    assert rel_l2(rho, g['jonas_dyn']) < TOL
    f, X = qb.absorption_spectra(jd, 2000)
    assert rel_l2(X, g['jonas_abs_X']) < TOL
    mono = qb.UnitaryModel(systems.vibronic_monomer(), hilbert_subspace='ge',
                           unit_convert=CM_FS)
    t, rho = qb.simulate_dynamics(mono, qb.unit_vec(0, 5), 300)
    assert rel_l2(rho, g['mono_dyn']) < TOL
    f, X = qb.absorption_spectra(mono, 3000, correlation_decay_time=1000)
    assert rel_l2(X, g['mono_abs_X']) < TOL


# ------------------------------------------------------------------ edge cases
def test_edge_cases(fmo_model):
    eom = fmo_model.equation_of_motion('ee')
    y0 = np.zeros(49, complex)
    y0[0] = 1
    # single output point equal to t0: result is save(y0)
    out = qb.integrate(eom, y0, np.array([0.0]))
    assert out.shape == (1, 49) and np.array_equal(out[0], y0)
    # t0 before the first output time (utils.py:39-44)
    a = qb.integrate(eom, y0, np.array([10.0, 20.0]), t0=0.0)
    b = qb.integrate(eom, y0, np.array([0.0, 10.0, 20.0]))
    assert rel_l2(a, b[1:]) < 1e-12
    # ragged (non-uniform) grid and repeated times
    c = qb.integrate(eom, y0, np.array([0.0, 3.0, 3.0, 47.5]))
    assert np.array_equal(c[1], c[2])
    # zero state stays zero
    z = qb.integrate(eom, np.zeros(49, complex), np.array([0.0, 5.0]))
    assert np.all(z == 0)
    with pytest.raises(TypeError):
        qb.integrate(lambda t, y: y, y0, np.array([0.0, 1.0]))
    with pytest.raises(ValueError):
        qb.integrate(eom, y0, np.array([1.0, 0.5]))
    with pytest.raises(NotImplementedError):
        qb.HEOMModel(systems.dimer(), aki_temp_corr=True)
    with pytest.raises(qb.operator_tools.SubspaceError):
        fmo_model.equation_of_motion('eg')


# ------------------------------------------------------------------- K3 ZOFE
def test_zofe_rhs_all_flag_branches(golden):
    g = golden('zofe')
    h3 = systems.fmo(bath='pseudomode', n_sites=3)
    for hh in (0, 1):
        for rh in (0, 1):
            m = qb.ZOFEModel(h3, hilbert_subspace='ge', unit_convert=CM_FS,
                             ham_hermit=bool(hh), rho_hermit=bool(rh))
            dy = m.equation_of_motion('ee')(0, g['fmo3_y_%d%d' % (hh, rh)])
            assert rel_l2(dy, g['fmo3_dy_%d%d' % (hh, rh)]) < 1e-13, (hh, rh)
    with pytest.raises(NotImplementedError):
        m.equation_of_motion('ee', heisenberg_picture=True)


def test_zofe_trajectory_and_absorption(golden):
    g = golden('zofe')
    m7 = qb.ZOFEModel(systems.fmo(bath='pseudomode'), hilbert_subspace='e',
                      unit_convert=CM_FS)
    t, rho = qb.simulate_dynamics(m7, np.eye(7)[0], 150, rtol=1e-11, atol=1e-13)
    assert np.array_equal(t, g['fmo7_t'])
    assert rel_l2(rho, g['fmo7_rho']) < TOL
    t, rho4 = qb.simulate_dynamics(m7, np.eye(7)[0], 150, method_name='rk4',
                                   rk4_substeps=40)
    assert rel_l2(rho4, g['fmo7_rho']) < TOL
    md = qb.ZOFEModel(systems.dimer(bath='pseudomode'), hilbert_subspace='ge',
                      unit_convert=CM_FS)
    f, X = qb.absorption_spectra(md, 1500, rtol=1e-11, atol=1e-13)
    assert np.array_equal(f, g['dimer_abs_f'])
    assert rel_l2(X, g['dimer_abs_X']) < TOL


def test_zofe_ensemble_batch_equals_single_runs():
    ham = systems.fmo(bath='pseudomode', n_sites=3)
    m = qb.ZOFEModel(ham, hilbert_subspace='e', unit_convert=CM_FS)
    t, avg = qb.simulate_dynamics(m, np.eye(3)[0], 60, ensemble_size=3)
    singles = [qb.simulate_dynamics(mm, np.eye(3)[0], 60)[1]
               for mm in m.sample_ensemble(3)]
    assert rel_l2(avg, np.mean(singles, axis=0)) < 1e-9


# ------------------------------------------------------- K5 device generator build
def test_device_redfield_build_matches_reference(golden):
    g = golden('redfield')
    f = qb.RedfieldModel(systems.fmo(), hilbert_subspace='e', unit_convert=CM_FS,
                         secular=False)
    # (a) eigensystems from the host (LAPACK), tensors on the device
    E, U = f.ensemble_eigensystems(3)
    bath = f.hamiltonian.bath
    number = np.einsum('jaa->ja', f.hamiltonian.system_bath_couplings('e'))
    L = engine.redfield_build(E, U, number, 0, bath.temperature, bath.reorg_energy,
                              bath.cutoff_freq, False, False, CM_FS,
                              f.liouville_subspace_index('ee')).cpu().numpy()
    for n in range(3):
        assert rel_l2(L[n], g['fmo_member%d_L' % n]) < 1e-12
    # (b) everything on the device (Jacobi eigensystems): probe the staged
    # generators through the apply entry point
    eom = f.ensemble_eom(3, False, 'ee')
    eye = np.eye(49, dtype=complex)
    for n in range(3):
        Ln = eom.apply(eye, generators=np.full(49, n)).T
        assert rel_l2(Ln, g['fmo_member%d_L' % n]) < 1e-11


@pytest.mark.parametrize('secular,dic,basis', [(True, True, 'site'), (True, False, 'site'),
                                               (False, True, 'eigen')])
def test_device_redfield_build_variants(secular, dic, basis):
    """'gef' (block-diagonal H: Jacobi must keep the manifolds in place) with
    secular / real-correlation / eigenbasis options, against the host builder."""
    ham = systems.dimer(disorder=60)
    m = qb.RedfieldModel(ham, hilbert_subspace='gef', unit_convert=CM_FS,
                         secular=secular, discard_imag_corr=dic, evolve_basis=basis)
    members = list(m.sample_ensemble(4))
    for ss in ('ee', 'eg', 'fe'):
        ref = m.ensemble_generators(members, ss)
        eom = m.ensemble_eom(4, False, ss)
        M = ref.shape[-1]
        for n in range(4):
            Ln = eom.apply(np.eye(M, dtype=complex), generators=np.full(M, n)).T
            if basis == 'site':
                assert rel_l2(Ln, ref[n]) < 1e-10, (ss, n)
            else:
                # eigenvector signs are a gauge in the eigenbasis (numpy vs scipy
                # LAPACK drivers): compare magnitudes
                assert rel_l2(np.abs(Ln), np.abs(ref[n])) < 1e-10, (ss, n)


@pytest.mark.parametrize('basis', ['site', 'eigen'])
def test_device_redfield_build_vibronic(basis):
    """Vibronic dimer (2 sites x one explicit two-level mode each: 1 + 8 states in 'ge') with
    static disorder: the members' eigensystems come from the host (Hamiltonian.eig), Redfield
    tensors and basis transform from K5 -- against the host builder that the golden
    fixtures pin, and an ensemble-averaged trajectory against per-member runs."""
    ham = systems.jonas_dimer(disorder=60)
    m = qb.RedfieldModel(ham, hilbert_subspace='ge', unit_convert=CM_FS, secular=False,
                         evolve_basis=basis)
    assert not m._device_buildable() and m._tensor_device_buildable()
    members = list(m.sample_ensemble(3))
    for ss in ('ee', 'eg'):
        ref = m.ensemble_generators(members, ss)
        eom = m.ensemble_eom(3, False, ss)
        M = ref.shape[-1]
        for n in range(3):
            Ln = eom.apply(np.eye(M, dtype=complex), generators=np.full(M, n)).T
            assert rel_l2(Ln, ref[n]) < 1e-10, (ss, n)
    psi0 = np.zeros(8)
    psi0[0] = 1
    me = qb.RedfieldModel(ham, hilbert_subspace='e', unit_convert=CM_FS, secular=False,
                          evolve_basis=basis)
    t, avg = qb.simulate_dynamics(me, psi0, 200, ensemble_size=3)
    singles = [qb.simulate_dynamics(mm, psi0, 200)[1] for mm in me.sample_ensemble(3)]
    assert rel_l2(avg, np.mean(singles, axis=0)) < 1e-10


def test_device_redfield_build_restricted_blocks():
    """4-site 'gef' (1 + 4 + 6 states): the builder forms only the ket x bra
    block each Liouville subspace touches; mixed subspaces and the non-secular
    tensor against the host builder (full 11^4 tensor, then sliced)."""
    ham = systems.synthetic_aggregate(4, disorder=70)
    for secular in (False, True):
        m = qb.RedfieldModel(ham, hilbert_subspace='gef', unit_convert=CM_FS,
                             secular=secular)
        members = list(m.sample_ensemble(3))
        for ss in ('fe', 'ef', 'ge', 'gg,ee', 'ge,ef', 'gg,ee,ff', 'eg,fe'):
            ref = m.ensemble_generators(members, ss)
            eom = m.ensemble_eom(3, False, ss)
            M = ref.shape[-1]
            for n in range(3):
                Ln = eom.apply(np.eye(M, dtype=complex), generators=np.full(M, n)).T
                assert rel_l2(Ln, ref[n]) < 1e-10, (ss, n, secular)
        heis = m.ensemble_eom(3, False, 'fe', heisenberg_picture=True)
        M = heis.dim
        Ln = heis.apply(np.eye(M, dtype=complex), generators=np.full(M, 1)).T
        assert rel_l2(Ln, m.ensemble_generators(members, 'fe')[1].T) < 1e-10


@pytest.mark.parametrize('M,n_gen', [(8, 3), (24, 3), (41, 2), (48, 2), (50, 600), (52, 3), (56, 3),
                                     (57, 2), (64, 5), (147, 3), (200, 2)])
def test_expm_kernels_all_tile_counts(M, n_gen):
    """exp(L dt) from the DMMA kernels for every padded size class (one 8-row block per warp,
    1..7 warps; 49..56 take the two-CTAs-per-SM kernel whose CTAs loop over the generators:
    600 generators exercise that loop and its per-CTA scratch tile; above 56 the tiled
    tensor-core GEMM of csrc/dense_wide.cu, one launch per product) against scipy's expm."""
    import scipy.linalg
    rng = np.random.RandomState(M)
    L = (rng.randn(n_gen, M, M) + 1j * rng.randn(n_gen, M, M)) / np.sqrt(M)
    L -= 0.5 * np.eye(M)
    eom = engine.DenseEOM(L)
    dt = 0.7
    prop = eom.propagator(dt)
    for n in sorted(set([0, n_gen // 2, n_gen - 1])):
        P = prop.apply(np.eye(M, dtype=complex), generators=np.full(M, n)).T
        assert rel_l2(P, scipy.linalg.expm(L[n] * dt)) < 1e-13, (M, n)


# ------------------------------------------------ tensor-core propagator (expm)
def test_expm_propagator_matches_reference(fmo_model, golden):
    g = golden('redfield')
    t, rho = qb.simulate_dynamics(fmo_model, np.eye(7)[0], 1000, method_name='expm')
    assert rel_l2(rho, g['fmo_rho_1ps']) < TOL
    eom = fmo_model.equation_of_motion('ee')
    assert eom.last['method'] == 'expm'
    # the propagator itself against scipy's expm of the reference generator
    import scipy.linalg
    dt = fmo_model.time_step
    P = eom.propagator(dt).apply(np.eye(49, dtype=complex)).T
    assert rel_l2(P, scipy.linalg.expm(g['fmo_L_ee'] * dt)) < 1e-13
    # default method picks it for long uniform grids and agrees with Taylor
    t, a = qb.simulate_dynamics(fmo_model, np.eye(7)[0], 3000)
    assert eom.last['method'] == 'expm'
    t, b = qb.simulate_dynamics(fmo_model, np.eye(7)[0], 3000, method_name='taylor')
    assert rel_l2(a, b) < 1e-10
    # small systems (dimer, M = 2..4) and the ensemble path
    t, rho = qb.simulate_dynamics(fmo_model, np.eye(7)[0], 300, ensemble_size=4,
                                  method_name='expm')
    assert rel_l2(rho, g['fmo_ens4_rho_300fs']) < TOL
    m = qb.RedfieldModel(systems.dimer(), hilbert_subspace='gef', unit_convert=CM_FS,
                         discard_imag_corr=True)
    f, X = qb.absorption_spectra(m, 10000, method_name='expm')
    assert rel_l2(X, g['dimer_abs_X']) < TOL
    with pytest.raises(ValueError):
        qb.integrate(eom, np.eye(49, dtype=complex)[0], np.array([0., 1., 5.]),
                     method_name='expm')


# ------------------------------------------------ pulse-driven HEOM (grid-resident DOPRI5)
def test_heom_pump_matches_oracle():
    ham = systems.dimer()
    kw = dict(hilbert_subspace='gef', unit_convert=CM_FS, level_cutoff=3, K=1)
    hm = qb.HEOMModel(ham, **kw)
    ho = oracle.OracleHEOM(ham, **kw)
    pump = qb.GaussianPulse(12800, 40, scale=1e-3, freq_convert=CM_FS)
    t, st = qb.simulate_pump(hm, pump, 'x', time_extra=100, rtol=1e-11, atol=1e-14)
    to, so = oracle.simulate_with_fields(ho, [pump, pump], '-+', 'xx', time_extra=100,
                                         **TIGHT)
    assert np.array_equal(t, to)
    assert st.shape == so.shape == (len(t), 15 * 9)
    assert rel_l2(st, so) < TOL
    # free HEOM evolution with the adaptive integrator agrees with the Taylor default
    y0 = hm.density_matrix_to_state_vector(np.diag([1., 0]).astype(complex), 'ee')
    tt = np.arange(0, 200, hm.time_step)
    a = qb.integrate(hm.equation_of_motion('ee'), y0, tt, method_name='dopri5',
                     rtol=1e-11, atol=1e-14)
    b = qb.integrate(hm.equation_of_motion('ee'), y0, tt)
    assert rel_l2(a, b) < 1e-9


def test_zofe_pump_matches_oracle():
    """pulse-driven ZOFE: the dipole operator acts on rho and on every auxiliary
    operator; checked against a host restatement of the reference closure
    (eom.py:87-94 with ZOFESpaceOperator, zofe.py:8-41)."""
    ham = systems.dimer(bath='pseudomode')
    zm = qb.ZOFEModel(ham, hilbert_subspace='ge', unit_convert=CM_FS)
    zo = oracle.OracleZOFE(ham, hilbert_subspace='ge', unit_convert=CM_FS)
    pump = qb.GaussianPulse(12800, 40, scale=1e-3, freq_convert=CM_FS)
    t, st = qb.simulate_pump(zm, pump, 'x', time_extra=60, rtol=1e-11, atol=1e-14)
    eom = zo.equation_of_motion('gg,ge,eg,ee')
    Vm = zm.hamiltonian.dipole_operator('ge', 'x', '-')
    Vp = zm.hamiltonian.dipole_operator('ge', 'x', '+')
    n, shape = 3, zm.oop_shape

    def comm(V, y):
        rho = y[:n * n].reshape((n, n), order='F')
        O = y[n * n:].reshape(shape, order='F')
        drho = V @ rho - rho @ V
        dO = np.einsum('cd,psde->psce', V, O) - O @ V
        return np.append(drho.reshape(-1, order='F'), dO.reshape(-1, order='F'))

    def rhs(tt, y):
        E = pump(tt, zm.rw_freq)
        return eom(tt, y) + (-1j * E) * comm(Vm, y) + (-1j * np.conj(E)) * comm(Vp, y)

    ref = oracle.integrate(rhs, zm.thermal_state('gg,ge,eg,ee'), t, t0=pump.t_init, **TIGHT)
    assert rel_l2(st, ref) < TOL


def test_batched_third_order_ensemble(golden):
    """the device-batched ensemble path of third_order_response against the
    reference's serial ensemble loop (golden) and against per-member runs"""
    g = golden('response')
    t2 = np.linspace(0, 200, 3)
    dred = qb.RedfieldModel(systems.dimer(disorder=80), hilbert_subspace='gef',
                            unit_convert=CM_FS, discard_imag_corr=True)
    _, S = qb.third_order_response(dred, 300, population_times=t2, ensemble_size=3,
                                   include_signal='GSB,ESE')
    assert rel_l2(S, g['red_ens3_gsb_ese']) < TOL
    # random orientations (per-member dipole operators) + isotropic average
    _, A = qb.third_order_response(dred, 200, population_times=t2[:2], ensemble_size=2,
                                   ensemble_random_orientations=True,
                                   polarization='xxyy', exact_isotropic_average=True)
    singles = [qb.third_order_response(m, 200, population_times=t2[:2],
                                       polarization='xxyy',
                                       exact_isotropic_average=True)[1]
               for m in dred.sample_ensemble(2, True)]
    assert rel_l2(A, np.mean(singles, axis=0)) < 1e-9
    # non-uniform population times exercise the Taylor branch of the t2 stage
    tn = np.array([0., 15., 70., 200.])
    _, B = qb.third_order_response(dred, 200, population_times=tn, ensemble_size=2)
    singles = [qb.third_order_response(m, 200, population_times=tn)[1]
               for m in dred.sample_ensemble(2)]
    assert rel_l2(B, np.mean(singles, axis=0)) < 1e-9


def test_t2_stage_gemm_path_matches_kernel_path(golden, monkeypatch):
    """The t2 stage of a batched response function (n_t1 columns per unit under one propagator
    and one save operator) takes the tensor-core GEMM form of propagator stepping from 4096
    columns on (csrc/dense_wide.cu: qsx_dense_map_gemm; also the split K6 contraction): against
    the per-group stepping kernel (QSX_MAP_NO_GEMM=1), for a small state (dimer) and the wide FMO
    'fe' stage, with shared and per-configuration dipole operators -- and the FMO single-model
    fixture of the reference stays the anchor of both."""
    t2 = np.linspace(0, 200, 3)
    dred = qb.RedfieldModel(systems.dimer(disorder=80), hilbert_subspace='gef',
                            unit_convert=CM_FS, discard_imag_corr=True)
    fred = qb.RedfieldModel(systems.fmo(), hilbert_subspace='gef', unit_convert=CM_FS)
    cases = [(dred, dict(ensemble_size=96)),                                   # 96 x 1 x 52 columns
             (dred, dict(ensemble_size=8, polarization='xxyy', exact_isotropic_average=True)),
             (fred, dict(ensemble_size=6, exact_isotropic_average=True,
                         include_signal='ESA'))]                               # 6 x 21 x 79: 'ee' -> 'fe' save
    for model, kw in cases:
        res = {}
        for flag in ('0', '1'):
            if flag == '1':
                monkeypatch.setenv('QSX_MAP_NO_GEMM', '1')
            else:
                monkeypatch.delenv('QSX_MAP_NO_GEMM', raising=False)
            _, res[flag] = qb.third_order_response(model, 400 if model is fred else 500,
                                                   population_times=t2, **kw)
        assert rel_l2(res['0'], res['1']) < 1e-12, kw
    monkeypatch.delenv('QSX_MAP_NO_GEMM', raising=False)


# ------------------------------------------------- device Fourier transform (K7)
@pytest.mark.parametrize('shape,axis,sign', [((103,), 0, 1), ((37, 5, 41), 0, -1),
                                             ((37, 5, 41), 2, 1), ((6, 130, 3), 1, -1),
                                             ((1028,), -1, 1)])
def test_device_fourier_transform_matches_host(shape, axis, sign):
    """the device transform against the host restatement of the reference's
    symmetrise-pad + shifted FFT (simulate/utils.py:128-219)"""
    import torch
    rng = np.random.RandomState(5)
    x = rng.randn(*shape) + 1j * rng.randn(*shape)
    t = 4.0 * np.arange(shape[axis])
    f_host, X_host = qb.fourier_transform(t, x, axis, rw_freq=12345.0, unit_convert=CM_FS,
                                          sign=sign)
    f_dev, X_dev = qb.fourier_transform(t, torch.from_numpy(x).cuda(), axis, rw_freq=12345.0,
                                        unit_convert=CM_FS, sign=sign)
    assert X_dev.is_cuda and tuple(X_dev.shape) == X_host.shape
    np.testing.assert_array_equal(f_dev, f_host)
    assert rel_l2(X_dev.cpu().numpy(), X_host) < 1e-13
    # grids that do not start at zero fall back to the host restatement
    f2, X2 = qb.fourier_transform(t + 8.0, torch.from_numpy(x).cuda(), axis, sign=sign)
    f3, X3 = qb.fourier_transform(t + 8.0, x, axis, sign=sign)
    assert isinstance(X2, np.ndarray) and rel_l2(X2, X3) < 1e-15


def test_two_dimensional_spectra_ensemble_on_device(golden):
    """disorder-ensemble 2D spectrum: batched propagation + device transforms against
    the reference's Fourier transform of the (golden) ensemble response"""
    g = golden('response')
    t2 = np.linspace(0, 200, 3)
    dred = qb.RedfieldModel(systems.dimer(disorder=80), hilbert_subspace='gef',
                            unit_convert=CM_FS, discard_imag_corr=True)
    (f1, _, f3), X = qb.two_dimensional_spectra(dred, 300, population_times=t2,
                                                ensemble_size=3, include_signal='GSB,ESE')
    t1 = np.arange(0, 300, dred.time_step)
    r1, R = qb.fourier_transform(t1, g['red_ens3_gsb_ese'], 0, rw_freq=dred.rw_freq,
                                 sign=-1, unit_convert=dred.unit_convert)
    r3, R = qb.fourier_transform(t1, R, 2, rw_freq=dred.rw_freq,
                                 unit_convert=dred.unit_convert)
    np.testing.assert_allclose(f1, r1)
    np.testing.assert_allclose(f3, r3)
    assert rel_l2(X, R) < TOL


def test_device_disorder_streams_match_numpy():
    """static-disorder shifts generated on the device against numpy's seeded legacy
    streams (hamiltonian.py:458-461, 566-573): same integer stream, Box-Muller within
    a few ulp of the host libm"""
    ham = systems.fmo()
    dev = ham.sampled_site_shifts_device(300, member0=5).cpu().numpy()
    host = ham.sampled_site_shifts(300, member0=5)
    ref = np.array([ham.sample(n).H_1exc.diagonal() - ham.H_1exc.diagonal()
                    for n in range(5, 12)])
    assert np.abs(host[:7] - ref).max() < 1e-9          # (H + shift) - H rounding only
    assert dev.shape == host.shape
    assert np.abs(dev - host).max() <= 8 * np.finfo(float).eps * np.abs(host).max()
    assert (dev == host).mean() > 0.5


@pytest.mark.parametrize('prefix,member0,n_gauss', [([7, 11], 0, 7), ([7, 11], 0, 130), ([3], 5, 400)])
def test_device_gauss_streams_long(prefix, member0, n_gauss):
    """the device replay of RandomState(prefix + [n]).randn(n_gauss): outputs taken straight from
    the seeded state (first 227 words), after the first full twist (130 Gaussians need ~330
    words) and after the second (400 need ~1000)"""
    from qspectra_b200 import _capi
    n = 257
    dev = _capi.sample_gauss_device(prefix, member0, n, n_gauss).cpu().numpy()
    ref = np.array([np.random.RandomState(list(prefix) + [member0 + k]).randn(n_gauss) for k in range(n)])
    assert np.abs(dev - ref).max() <= 16 * np.finfo(float).eps * np.abs(ref).max()
    assert (dev == ref).mean() > 0.5


def test_wide_state_propagator_stepping():
    """state dimensions above the single-CTA propagator kernel's 56 (FMO 'fe', M = 147): tiled
    tensor-core propagator build + streamed propagator stepping against Taylor"""
    mf = qb.RedfieldModel(systems.fmo(), hilbert_subspace='gef', unit_convert=CM_FS)
    rng = np.random.RandomState(2)
    for heis in (False, True):
        eom = mf.equation_of_motion('fe', heisenberg_picture=heis)
        assert eom.dim == 147
        y0 = rng.randn(2, 147) + 1j * rng.randn(2, 147)
        t = mf.time_step * np.arange(80)
        a = eom.propagate(y0, t)
        assert eom.last['method'] == 'expm'
        b = eom.propagate(y0, t, method='taylor')
        assert rel_l2(a, b) < 1e-10
    ens = mf.ensemble_eom(3, False, 'fe', heisenberg_picture=True)
    y0 = rng.randn(3, 147) + 1j * rng.randn(3, 147)
    a = ens.propagate(y0, t, generators=np.arange(3))
    b = ens.propagate(y0, t, generators=np.arange(3), method='taylor')
    assert ens.last['method'] == 'taylor' and rel_l2(a, b) < 1e-10


# ------------------------------------------------------------ round-2 fixtures
def test_fmo_gef_third_order_pathways(golden):
    """FMO 'gef' third-order response, one pathway class at a time (GSB: gg/eg stages, ESE:
    49-dimensional 'ee' stage, ESA: the 147-dimensional 'fe' stage -- the wide-state
    propagator path) against the reference."""
    g = golden('round2')
    m = qb.RedfieldModel(systems.fmo(), hilbert_subspace='gef', unit_convert=CM_FS,
                         secular=False)
    for sig in ('GSB', 'ESE', 'ESA'):
        (t1, t2, t3), S = qb.third_order_response(m, 400, population_times=g['fmo_gef_t2'],
                                                  include_signal=sig)
        assert np.array_equal(t1, g['fmo_gef_t1'])
        assert rel_l2(S, g['fmo_gef_%s' % sig]) < TOL, sig


def test_fmo_gef_device_generator_blocks(golden):
    """K5 on the 'gef' manifold of FMO (36 states): the blocks the third-order pathways use,
    disorder member 3, against the reference's evolution_super_operator."""
    g = golden('round2')
    m = qb.RedfieldModel(systems.fmo(), hilbert_subspace='gef', unit_convert=CM_FS,
                         secular=False)
    for ss in ('fe', 'eg', 'ee'):
        eom = m.ensemble_eom(1, False, ss, member0=3)
        M = eom.dim
        L = eom.apply(np.eye(M, dtype=complex), generators=np.zeros(M, dtype=int)).T
        assert rel_l2(L, g['fmo_gef_member3_L_%s' % ss]) < 1e-10, ss


def test_eigen_basis_ensemble_third_order(golden):
    """evolve_basis='eigen' with static disorder: every member has its own eigenbasis, so dipole
    operators and the thermal state must be the member's (reference decorators.py:55-60)."""
    g = golden('round2')
    de = qb.RedfieldModel(systems.dimer(disorder=80), hilbert_subspace='gef',
                          unit_convert=CM_FS, discard_imag_corr=True, evolve_basis='eigen')
    (_, _, _), S = qb.third_order_response(de, 300, population_times=np.linspace(0, 200, 3),
                                           ensemble_size=3)
    assert rel_l2(S, g['dimer_eigen_ens3']) < TOL


def test_response_contraction_kernel():
    """K6 (csrc/dense_wide.cu): S[ab, c] += sum_u w_u sum_i X[u, ab, i] Y[u, c, i] against numpy,
    ragged sizes (tile edges, K not a multiple of the chunk), accumulation into a non-zero S."""
    import torch
    rng = np.random.RandomState(6)
    for U, n_ab, n_c, K in [(1, 5, 3, 2), (7, 70, 45, 7), (3, 197 * 2, 197, 49), (2, 33, 65, 147)]:
        X = rng.randn(U, n_ab, K) + 1j * rng.randn(U, n_ab, K)
        Y = rng.randn(U, n_c, K) + 1j * rng.randn(U, n_c, K)
        w = rng.randn(U) + 1j * rng.randn(U)
        S0 = rng.randn(n_ab, n_c) + 1j * rng.randn(n_ab, n_c)
        S = torch.from_numpy(S0.copy()).cuda()
        engine.response_contract(torch.from_numpy(X).cuda(), torch.from_numpy(Y).cuda(),
                                 torch.from_numpy(w).cuda(), S)
        want = S0 + np.einsum('u,uai,uci->ac', w, X, Y)
        assert rel_l2(S.cpu().numpy(), want) < 1e-13, (U, n_ab, n_c, K)


def test_heom_third_order_ensemble_isotropic_one_batch():
    """HEOM third-order response with static disorder and the exact isotropic average
    (members x 21 polarisation configurations as columns of one device batch per stage, per-ADO
    dipole blocks selected per column) against member-by-member, configuration-by-
    configuration runs combined on the host (reference decorators.py:40-96)."""
    from qspectra_b200.simulate.response import _polarization_variants
    ham = systems.dimer(disorder=60)
    m = qb.HEOMModel(ham, hilbert_subspace='gef', unit_convert=CM_FS, level_cutoff=3,
                     low_temp_corr=False)
    t2 = np.array([0., 60.])
    E = 3
    (_, _, _), S = qb.third_order_response(m, 120, population_times=t2, polarization='xxyy',
                                           ensemble_size=E, exact_isotropic_average=True)
    want = 0
    for member in m.sample_ensemble(E):
        for w, pol in _polarization_variants('xxyy', True):
            (_, _, _), Sp = qb.third_order_response(member, 120, population_times=t2,
                                                    polarization=pol)
            want = want + w * Sp
    assert rel_l2(S, want / E) < 1e-9


def test_linear_response_batched_matches_member_loop():
    """absorption with disorder + random orientations + exact isotropic average: one batch
    (members x xx/yy/zz) against the member / configuration loop."""
    ham = systems.fmo(n_sites=4)
    m = qb.RedfieldModel(ham, hilbert_subspace='ge', unit_convert=CM_FS)
    t, x = qb.linear_response(m, 'gg->eg->gg', 400, polarization='xx', ensemble_size=3,
                              ensemble_random_orientations=True, exact_isotropic_average=True)
    want = 0
    for member in m.sample_ensemble(3, True):
        for p in ('xx', 'yy', 'zz'):
            want = want + qb.linear_response(member, 'gg->eg->gg', 400, polarization=p)[1] / 3.0
    assert rel_l2(x, want / 3) < 1e-10
