"""
Named benchmark systems (BASELINE.json configs).  Every builder takes the
namespace that provides the Hamiltonian/bath classes (``qspectra_b200`` by
default; the golden-vector generator passes the real ``qspectra`` so both
sides are built from literally the same numbers).

Parameter sources: FMO 7-site Hamiltonian, dipoles and Debye bath from the
reference's ``examples/FMO dynamics with Redfield theory.ipynb`` (cell 3);
dimer from ``tests/test_spectra.py:31-35`` / ``examples/HEOM vs Redfield vs
ZOFE.ipynb``; vibronic (Jonas) dimer and monomer from the corresponding
notebooks; the 16-pseudomode table from ``examples/FMO dynamics with ZOFE
master equation.ipynb``.
"""
import numpy as np

from .constants import CM_K

FMO_H = np.array([
    [12400, -87.7, 5.5, -5.9, 6.7, -13.7, -9.9],
    [-87.7, 12520, 30.8, 8.2, 0.7, 11.8, 4.3],
    [5.5, 30.8, 12200, -53.5, -2.2, -9.6, 6.],
    [-5.9, 8.2, -53.5, 12310, -70.7, -17., -63.3],
    [6.7, 0.7, -2.2, -70.7, 12470, 81.1, -1.3],
    [-13.7, 11.8, -9.6, -17., 81.1, 12620, 39.7],
    [-9.9, 4.3, 6., -63.3, -1.3, 39.7, 12430]])

_FMO_DIPOLE_RAW = np.array(
    [[3.019, 3.442, 0.797, 3.213, 2.969, 0.547, 1.983],
     [2.284, -2.023, -3.871, 2.145, -2.642, 3.562, 2.837],
     [1.506, 0.431, 0.853, 1.112, -0.661, -1.851, 2.015]]).T
FMO_DIPOLES = np.array([d / np.linalg.norm(d) for d in _FMO_DIPOLE_RAW])

DIMER_H = np.array([[12881., 120.], [120., 12719.]])
DIMER_DIPOLES = np.array([[1., 0., 0.],
                          [2. * np.cos(.3), 2. * np.sin(.3), 0.]])

PM_OMEGA = [-500., -200., -90., 1., 21., 60., 80., 130., 200., 300., 400.,
            500., 600., 800., 1100., 1500.]
PM_GAMMA = [500., 100., 50., 50., 50., 50., 80., 40., 80., 150., 200., 200.,
            80., 250., 200., 300.]
PM_HUANG = [-2.5133e-03, -7.5398e-03, -2.5133e-02, 5.0265e+01, 2.2619e+00,
            4.5239e-02, 2.7646e-01, 9.2991e-03, 2.2619e-02, 1.5080e-02,
            3.0159e-03, 3.5186e-03, 2.8274e-04, 1.7593e-03, 4.3982e-04,
            4.3982e-04]


def _ns(ns):
    if ns is None:
        import qspectra_b200 as ns
    return ns


def debye_bath(ns=None, reorg=35., cutoff=106., kelvin=77.):
    return _ns(ns).DebyeBath(CM_K * kelvin, reorg, cutoff)


def pseudomode_bath(n_sites, ns=None, n_pm=16):
    on = np.ones(n_sites, complex)
    cols = [np.array([v * on for v in tab[:n_pm]])
            for tab in (PM_OMEGA, PM_GAMMA, PM_HUANG)]
    return _ns(ns).PseudomodeBath(n_pm, cols[0], cols[1], cols[2])


def fmo(ns=None, disorder=100, bath='debye', n_sites=7, **kw):
    """FMO monomer (first ``n_sites`` sites), Debye 77 K bath, FWHM disorder."""
    ns = _ns(ns)
    H = FMO_H[:n_sites, :n_sites]
    b = (debye_bath(ns) if bath == 'debye'
         else pseudomode_bath(n_sites, ns) if bath == 'pseudomode' else bath)
    return ns.ElectronicHamiltonian(H, bath=b, dipoles=FMO_DIPOLES[:n_sites],
                                    disorder=disorder, **kw)


def dimer(ns=None, disorder=None, bath='debye', **kw):
    ns = _ns(ns)
    b = (debye_bath(ns) if bath == 'debye'
         else pseudomode_bath(2, ns) if bath == 'pseudomode' else bath)
    return ns.ElectronicHamiltonian(DIMER_H, bath=b, dipoles=DIMER_DIPOLES,
                                    disorder=disorder, **kw)


def jonas_dimer(ns=None, levels=(2, 2), disorder=None):
    """Vibronic dimer with one explicit 200 cm^-1 mode per site."""
    ns = _ns(ns)
    el = ns.ElectronicHamiltonian(
        np.array([[11500, 66], [66, 11650]]), dipoles=[[1, 0, 0], [0, 1, 0]],
        bath=debye_bath(ns, reorg=1.5 * 35), disorder=disorder)
    return ns.VibronicHamiltonian(el, np.array(levels), np.array([200, 200]),
                                  -32 * np.eye(2))


def vibronic_monomer(ns=None, levels=5):
    ns = _ns(ns)
    el = ns.ElectronicHamiltonian([[11500]], dipoles=[[1, 0, 0]], bath=None)
    return ns.VibronicHamiltonian(el, n_vibrational_levels=[levels],
                                  vib_energies=[200],
                                  elec_vib_couplings=[[-100]])


def synthetic_aggregate(n_sites, ns=None, seed=0, disorder=100, bath='debye'):
    """Synthetic Frenkel-exciton aggregate of a named size: site energies
    ~ N(12400, 100 cm^-1 FWHM-scaled), nearest-neighbour couplings
    O(10-100 cm^-1) decaying with distance, unit dipoles."""
    ns = _ns(ns)
    rng = np.random.RandomState(seed)
    H = np.diag(12400 + 100 / 2.3548 * rng.randn(n_sites))
    for i in range(n_sites):
        for j in range(i + 1, n_sites):
            J = rng.uniform(-100, 100) / (j - i) ** 3
            H[i, j] = H[j, i] = np.round(J, 1)
    d = rng.randn(n_sites, 3)
    d /= np.linalg.norm(d, axis=1)[:, None]
    b = (debye_bath(ns) if bath == 'debye'
         else pseudomode_bath(n_sites, ns) if bath == 'pseudomode' else bath)
    return ns.ElectronicHamiltonian(H, bath=b, dipoles=d, disorder=disorder)
