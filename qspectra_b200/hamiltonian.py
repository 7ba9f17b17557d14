"""
System description: electronic / vibronic Frenkel-exciton Hamiltonians.

Contract: reference ``qspectra/hamiltonian.py`` -- thermal/ground state
:26-78, rotating frame :177-209, deterministic disorder sampling :233-274 and
:552-578 (member n uses ``RandomState(list(seed) + [n])``: diagonal ``randn``
draw first, then ``rand(3)`` for the orientation), eigensystem :310-328,
Nyquist ``freq_step``/``time_step`` of the *un-sampled* Hamiltonian :382-412,
dipole / number / system-bath operators :580-608, vibronic extension :633-782.

The rotating-frame / sampled variants keep the reference's bookkeeping
attributes (``_not_rotating``, ``_not_sampled``, ``rw_freq``) because the
dynamics layer reads them (time grid of the un-sampled parent, thermal state of
the lab-frame parent).  ``disorder_stream(n)`` exposes member n's seeded
generator so the device ensemble builder can replay exactly the same draws.
"""
from abc import ABCMeta, abstractmethod
from numbers import Number
import warnings

import numpy as np
import scipy.linalg

from .constants import GAUSSIAN_SD_FWHM
from .operator_tools import (transition_operator, operator_extend, unit_vec,
                             tensor, extend_vib_operator, vib_create,
                             vib_annihilate, hilbert_subspace_index,
                             basis_transform_vector, basis_transform_operator)
from .polarization import polarization_vector, random_rotation_matrix
from .utils import imemoize, memoized_property, check_random_state, simple_repr


class HamiltonianError(Exception):
    """Raised when a Hamiltonian lacks what an operation needs."""


def check_hermitian(matrix):
    matrix = np.asarray(matrix)
    if not np.allclose(matrix.conj().T, matrix):
        raise ValueError('matrix input must to be Hermitian')
    return matrix


def ground_state(hamiltonian_matrix):
    """Equal mixture of the (exactly) degenerate lowest eigenvectors."""
    E, U = scipy.linalg.eigh(hamiltonian_matrix)
    lowest = [np.outer(U[:, i], U[:, i]) for i in range(len(E)) if E[i] == E[0]]
    return np.mean(lowest, axis=0).astype(complex)


def thermal_state(hamiltonian_matrix, temperature):
    """exp(-H/T)/Z, or the ground state for T <= 0."""
    if temperature > 0:
        rho = scipy.linalg.expm(-np.asarray(hamiltonian_matrix)
                                / float(temperature))
        Z = np.trace(rho)
        if Z == 0 or np.isnan(rho).any():
            raise OverflowError(
                'temperature=%s too low to reliably calculate thermal_state; '
                'raise it or set it to zero (in which case ground_state is '
                'substituted' % temperature)
        rho = rho / Z
    else:
        rho = ground_state(hamiltonian_matrix)
    return rho.astype(complex)


def diagonal_gaussian_disorder(fwhm, n_sites):
    def disorder(random_state):
        return np.diag((fwhm * GAUSSIAN_SD_FWHM) * random_state.randn(n_sites))
    return disorder


class Hamiltonian(metaclass=ABCMeta):
    """Base class; subclasses provide ``H(subspace)`` plus the two hooks
    ``_in_rotating_frame`` and ``_sample``."""

    def __init__(self, energy_spread_extra=None, site_labels=None):
        self.energy_spread_extra = energy_spread_extra
        self.site_labels = site_labels
        self._not_sampled = self
        self._not_rotating = self
        self.rw_freq = 0

    @property
    def _original(self):
        return self._not_rotating._not_sampled

    # -- matrices ----------------------------------------------------------
    @abstractmethod
    def H(self, subspace):
        """system Hamiltonian matrix in the given Hilbert subspace"""

    def n_states(self, subspace):
        return len(self.H(subspace))

    @imemoize
    def ground_state(self, subspace):
        return ground_state(self._not_rotating.H(subspace))

    @imemoize
    def thermal_state(self, subspace):
        bath = getattr(self._not_rotating, 'bath', None)
        temperature = getattr(bath, 'temperature', 0)
        return thermal_state(self._not_rotating.H(subspace), temperature)

    # -- frames and ensembles -----------------------------------------------
    @imemoize
    def in_rotating_frame(self, rw_freq=None):
        if rw_freq is None:
            rw_freq = self._original.transition_energy
        lab = self._not_rotating
        ham = lab._in_rotating_frame(rw_freq)
        ham._not_rotating = lab
        if self._not_sampled is not self:
            ham._not_sampled = self._not_sampled.in_rotating_frame(rw_freq)
        ham.rw_freq = rw_freq
        return ham

    def _in_rotating_frame(self, rw_freq):
        raise NotImplementedError('%s does not implement rotating frame '
                                  'transformations' % type(self).__name__)

    def sample_ensemble(self, ensemble_size=1, random_orientations=False):
        for n in range(ensemble_size):
            yield self.sample(n, random_orientations)

    def sample(self, n=None, random_orientations=False):
        if n is None:
            n = np.random.randint(2 ** 30)
        ham = self._not_sampled._sample(n, random_orientations)
        if self._not_rotating is not self:
            ham._not_rotating = self._not_rotating.sample(n, random_orientations)
        ham._not_sampled = self._not_sampled
        ham.rw_freq = self.rw_freq
        return ham

    def _sample(self, n, random_orientations):
        raise NotImplementedError('%s does not implement ensemble sampling'
                                  % type(self).__name__)

    # -- operators -----------------------------------------------------------
    def dipole_operator(self, subspace='gef', polarization='x',
                        transitions='-+'):
        raise NotImplementedError('%s does not implement dipole operators'
                                  % type(self).__name__)

    def system_bath_couplings(self, subspace='gef'):
        raise NotImplementedError('%s does not implement system-bath couplings'
                                  % type(self).__name__)

    # -- spectrum ------------------------------------------------------------
    @imemoize
    def eig(self, subspace):
        """(E, U) from the lab-frame matrix (keeps g < e < f ordering), with
        the rotating-frame shift applied to E afterwards."""
        E, U = scipy.linalg.eigh(self._not_rotating.H(subspace))
        for letter, quanta in (('e', 1), ('f', 2)):
            if letter in subspace:
                E[self.hilbert_subspace_index(letter, subspace)] -= \
                    quanta * self.rw_freq
        return (E, U)

    def E(self, subspace):
        return self.eig(subspace)[0]

    def U(self, subspace):
        return self.eig(subspace)[1]

    def transform_vector_to_eigenbasis(self, rho, subspace):
        return basis_transform_vector(rho, self.U(subspace))

    def transform_vector_from_eigenbasis(self, rho, subspace):
        return basis_transform_vector(rho, self.U(subspace).T.conj())

    def transform_operator_to_eigenbasis(self, rho, subspace):
        return basis_transform_operator(rho, self.U(subspace))

    def transform_operator_from_eigenbasis(self, rho, subspace):
        return basis_transform_operator(rho, self.U(subspace).T.conj())

    @property
    def transition_energy(self):
        return np.mean(self.E('e'))

    @property
    def freq_step(self):
        energies = self._not_sampled.E('gef')
        extra = (0.01 * self._original.transition_energy
                 if self.energy_spread_extra is None
                 else self.energy_spread_extra)
        return 2 * (max(energies.max(), -energies.min()) + extra)

    @property
    def time_step(self):
        return 1.0 / self.freq_step

    def hilbert_subspace_index(self, subspace, all_subspaces):
        return hilbert_subspace_index(subspace, all_subspaces, self.n_sites,
                                      self.n_vibrational_states)

    def basis_labels(self, subspace, braket=False):
        return self.site_labels


class ElectronicHamiltonian(Hamiltonian):
    """Frenkel-exciton Hamiltonian with identical independent baths per site."""

    def __init__(self, H_1exc, bath=None, dipoles=None, disorder=None,
                 random_seed=0, energy_spread_extra=None, site_labels=None):
        self.H_1exc = check_hermitian(H_1exc)
        self.bath = bath
        self.dipoles = np.asarray(dipoles) if dipoles is not None else None
        self.disorder = disorder
        self.random_seed = random_seed
        self.n_vibrational_states = 1
        super(ElectronicHamiltonian, self).__init__(energy_spread_extra,
                                                    site_labels)

    def __repr__(self):
        return simple_repr(self, ['H_1exc', 'bath', 'dipoles', 'disorder',
                                  'random_seed', 'energy_spread_extra'])

    def __eq__(self, other):
        return self._eq(other, 1)

    def __ne__(self, other):
        return not self == other

    __hash__ = object.__hash__

    def _eq(self, other, depth):
        if not isinstance(other, ElectronicHamiltonian):
            return False
        same = (np.all(self.H_1exc == other.H_1exc)
                and self.bath == other.bath
                and np.all(self.dipoles == other.dipoles)
                and self.disorder == other.disorder
                and np.all(np.atleast_1d(self.random_seed)
                           == np.atleast_1d(other.random_seed))
                and self.energy_spread_extra == other.energy_spread_extra
                and self.rw_freq == other.rw_freq
                and self.site_labels == other.site_labels)
        if same and depth:
            same = (self._not_sampled._eq(other._not_sampled, depth - 1)
                    and self._not_rotating._eq(other._not_rotating, depth - 1))
        return bool(same)

    @property
    def n_sites(self):
        return len(self.H_1exc)

    @imemoize
    def H(self, subspace):
        return operator_extend(self.H_1exc, subspace)

    def _spawn(self, H_1exc, dipoles, seed):
        return type(self)(H_1exc, self.bath, dipoles, self.disorder, seed,
                          self.energy_spread_extra, self.site_labels)

    def _in_rotating_frame(self, rw_freq):
        return self._spawn(self.H_1exc - rw_freq * np.identity(self.n_sites),
                           self.dipoles, self.random_seed)

    def disorder_stream(self, n):
        """The seeded generator of ensemble member ``n``."""
        return check_random_state(list(np.atleast_1d(self.random_seed)) + [n])

    def sampled_site_shifts(self, ensemble_size, member0=0):
        """(ensemble_size, n_sites) diagonal static-disorder shifts of members
        member0 .. member0+ensemble_size-1 -- bit-identical to what
        ``sample(n)`` adds -- or None when ``disorder`` is a user callable.
        Uses the C replay of numpy's seeded streams (qsx_sample_streams), so no
        Hamiltonian object is created per member."""
        base = self._not_sampled
        if base.disorder is None:
            return np.zeros((ensemble_size, base.n_sites))
        if not isinstance(base.disorder, Number):
            return None
        from ._capi import sample_streams
        gauss, _ = sample_streams(base.random_seed, member0, ensemble_size,
                                  base.n_sites)
        return (base.disorder * GAUSSIAN_SD_FWHM) * gauss

    def sampled_site_shifts_device(self, ensemble_size, member0=0):
        """The same shifts as a CUDA tensor generated on the device (kernel
        `sample_streams_kernel`: integer stream identical, Box-Muller through the
        device log/sqrt), or None when `disorder` is a user callable."""
        base = self._not_sampled
        if base.disorder is not None and not isinstance(base.disorder, Number):
            return None
        from . import _capi
        if base.disorder is None:
            torch = _capi.torch_cuda()
            return torch.zeros((ensemble_size, base.n_sites), dtype=torch.float64, device='cuda')
        return _capi.sample_gauss_device(base.random_seed, member0, ensemble_size,
                                         base.n_sites, base.disorder * GAUSSIAN_SD_FWHM)

    def _sample(self, n, random_orientations):
        rng = self.disorder_stream(n)
        if self.disorder is None:
            if not random_orientations:
                warnings.warn('called sample with `disorder=None` and '
                              '`random_orientations=False`: sampled '
                              'Hamiltonian is identical to original',
                              RuntimeWarning, stacklevel=2)
            shift = 0
        elif isinstance(self.disorder, Number):
            shift = diagonal_gaussian_disorder(self.disorder, self.n_sites)(rng)
        else:
            shift = self.disorder(rng)
        dipoles = self.dipoles
        if random_orientations:
            dipoles = np.einsum('mn,in->im', random_rotation_matrix(rng),
                                self.dipoles)
        return self._spawn(self.H_1exc + shift, dipoles,
                           list(np.atleast_1d(self.random_seed)) + [n])

    def dipole_operator(self, subspace='gef', polarization='x',
                        transitions='-+'):
        if self.dipoles is None:
            raise HamiltonianError('transition dipole moments undefined')
        mu = self.dipoles @ polarization_vector(polarization)
        ops = [transition_operator(n, self.n_sites, subspace, transitions)
               for n in range(self.n_sites)]
        return np.einsum('nij,n->ij', ops, mu)

    def number_operator(self, site, subspace='gef'):
        return operator_extend(
            np.diag(unit_vec(site, self.n_sites, dtype=float)), subspace)

    def system_bath_couplings(self, subspace='gef'):
        if self.bath is None:
            raise HamiltonianError('bath undefined')
        return np.array([self.number_operator(n, subspace)
                         for n in range(self.n_sites)])

    def basis_labels(self, subspace='gef', braket=False):
        place = np.array([10 ** (self.n_sites - i - 1)
                          for i in range(self.n_sites)], dtype='O')
        fock = np.diag(operator_extend(np.diag(place), subspace))
        labels = [str(i).zfill(self.n_sites) for i in fock]
        if self.site_labels is not None:
            labels = [','.join(lab for i, lab in enumerate(self.site_labels)
                               if st[i] == '1') for st in labels]
            if 'g' in subspace:
                labels[0] = 'g'
        return ['|{}>'.format(x) for x in labels] if braket else labels


class VibronicHamiltonian(Hamiltonian):
    """Electronic Hamiltonian (x) explicit harmonic modes with linear
    coupling c_nm |n><n| (b_m + b_m^+)."""

    def __init__(self, electronic, n_vibrational_levels, vib_energies,
                 elec_vib_couplings, energy_spread_extra=None,
                 site_labels=None):
        self.electronic = electronic
        self.bath = electronic.bath
        self.n_sites = electronic.n_sites
        self.n_vibrational_levels = np.asarray(n_vibrational_levels)
        self.vib_energies = np.asarray(vib_energies)
        self.elec_vib_couplings = np.asarray(elec_vib_couplings)
        super(VibronicHamiltonian, self).__init__(energy_spread_extra,
                                                  site_labels)
        # reference quirk (hamiltonian.py:667 then :673): the constructor
        # argument wins over the electronic part's value
        self.energy_spread_extra = energy_spread_extra

    def __repr__(self):
        return simple_repr(self, ['electronic', 'n_vibrational_levels',
                                  'vib_energies', 'elec_vib_couplings',
                                  'energy_spread_extra'])

    def __eq__(self, other):
        return (isinstance(other, VibronicHamiltonian)
                and self.electronic == other.electronic
                and np.all(self.n_vibrational_levels
                           == other.n_vibrational_levels)
                and np.all(self.vib_energies == other.vib_energies)
                and np.all(self.elec_vib_couplings == other.elec_vib_couplings)
                and self.rw_freq == other.rw_freq)

    def __ne__(self, other):
        return not self == other

    __hash__ = object.__hash__

    @memoized_property
    def n_vibrational_states(self):
        return np.prod(self.n_vibrational_levels)

    @memoized_property
    def H_vibrational(self):
        H_vib = np.zeros((self.n_vibrational_states,) * 2)
        for m, (levels, energy) in enumerate(zip(self.n_vibrational_levels,
                                                 self.vib_energies)):
            H_vib += energy * extend_vib_operator(
                self.n_vibrational_levels, m, np.diag(np.arange(levels)))
        return H_vib

    def H_electronic_vibrational(self, subspace='gef'):
        dim = self.electronic.n_states(subspace) * self.n_vibrational_states
        out = np.zeros((dim, dim))
        for i in range(self.n_sites):
            number = self.electronic.number_operator(i, subspace)
            for m, levels in enumerate(self.n_vibrational_levels):
                q = vib_annihilate(levels) + vib_create(levels)
                out += self.elec_vib_couplings[i, m] * tensor(
                    number, extend_vib_operator(self.n_vibrational_levels, m, q))
        return out

    @imemoize
    def H(self, subspace='gef'):
        return (self.el_to_sys_operator(self.electronic.H(subspace))
                + self.vib_to_sys_operator(self.H_vibrational, subspace)
                + self.H_electronic_vibrational(subspace))

    def _respawn(self, electronic):
        return type(self)(electronic, self.n_vibrational_levels,
                          self.vib_energies, self.elec_vib_couplings,
                          self.energy_spread_extra, self.site_labels)

    def _in_rotating_frame(self, rw_freq):
        return self._respawn(self.electronic.in_rotating_frame(rw_freq))

    def _sample(self, n, random_orientations):
        return self._respawn(self.electronic.sample(n, random_orientations))

    def el_to_sys_operator(self, el_operator):
        return tensor(el_operator, np.eye(self.n_vibrational_states))

    def vib_to_sys_operator(self, vib_operator, subspace='gef'):
        return tensor(np.eye(self.electronic.n_states(subspace)), vib_operator)

    def dipole_operator(self, *args, **kwargs):
        return self.el_to_sys_operator(
            self.electronic.dipole_operator(*args, **kwargs))

    def system_bath_couplings(self, *args, **kwargs):
        return self.el_to_sys_operator(
            self.electronic.system_bath_couplings(*args, **kwargs))
