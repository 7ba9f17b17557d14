"""
System description: electronic / vibronic Frenkel-exciton Hamiltonians.

Contract: reference ``qspectra/hamiltonian.py`` -- thermal/ground state
:26-78, rotating frame :177-209, deterministic disorder sampling :233-274 and
:552-578 (member n uses ``RandomState(list(seed) + [n])``: diagonal ``randn``
draw first, then ``rand(3)`` for the orientation), eigensystem :310-328,
Nyquist ``freq_step``/``time_step`` of the *un-sampled* Hamiltonian :382-412,
dipole / number / system-bath operators :580-608, vibronic extension :633-782.

Sampled and rotating-frame Hamiltonians are *variants* of a root object, addressed
by (member, rw_freq) coordinates (see ``Hamiltonian``); the relatives the reference
reaches through ``_not_rotating`` / ``_not_sampled`` links are the properties
``lab_frame`` / ``unsampled`` here.  ``disorder_stream(n)`` exposes member n's seeded
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
    """A system Hamiltonian, or rather one *variant* of one.

    The object a user constructs is the **root**: lab frame, no disorder applied.
    Every other Hamiltonian the engine meets is the root seen through two coordinates,

        member   None, or (n, random_orientations): ensemble member n
        rw_freq  rotating-wave frequency subtracted per excitation quantum

    and is made by the root's ``_build(member, rw_freq)`` hook (subclasses) through
    ``_variant``.  ``in_rotating_frame`` and ``sample`` only move along one coordinate, so
    they commute by construction.  Quantities the reference defines on a particular
    relative (reference hamiltonian.py:96-455) name that relative explicitly here:

        thermal / ground state, eigenvectors   -> ``lab_frame``  (same member, rw_freq 0)
        frequency / time step                  -> ``unsampled``  (same frame, no member)
        default rotating frequency             -> the root's mean 'e' eigenenergy
    """

    def __init__(self, energy_spread_extra=None, site_labels=None):
        self.energy_spread_extra = energy_spread_extra
        self.site_labels = site_labels
        self.rw_freq = 0
        self._root = self
        self._member = None
        self._frames = {}           # un-sampled variants of a root, by rw_freq

    # -- variants ------------------------------------------------------------
    def _variant(self, member, rw_freq):
        root = self._root
        if member is None and rw_freq == 0:
            return root
        if member is None and rw_freq in root._frames:
            return root._frames[rw_freq]
        ham = root._build(member, rw_freq)
        ham._root, ham._member, ham.rw_freq = root, member, rw_freq
        if member is None:
            root._frames[rw_freq] = ham      # members are not cached: ensembles are large
        return ham

    @abstractmethod
    def _build(self, member, rw_freq):
        """Concrete Hamiltonian for the given coordinates (called on the root)."""

    @property
    def lab_frame(self):
        return self if self.rw_freq == 0 else self._variant(self._member, 0)

    @property
    def unsampled(self):
        return self if self._member is None else self._variant(None, self.rw_freq)

    def in_rotating_frame(self, rw_freq=None):
        if rw_freq is None:
            rw_freq = self._root.transition_energy
        return self if rw_freq == self.rw_freq else self._variant(self._member, rw_freq)

    def sample(self, n=None, random_orientations=False):
        """Ensemble member ``n`` of the root (sampling a member re-samples the root, so the
        distribution does not depend on what ``sample`` is called on)."""
        if n is None:
            n = np.random.randint(2 ** 30)
        return self._variant((int(n), bool(random_orientations)), self.rw_freq)

    def sample_ensemble(self, ensemble_size=1, random_orientations=False):
        return (self.sample(n, random_orientations) for n in range(ensemble_size))

    # -- matrices --------------------------------------------------------------
    @abstractmethod
    def H(self, subspace):
        """system Hamiltonian matrix in the given Hilbert subspace"""

    def n_states(self, subspace):
        return len(self.H(subspace))

    def dipole_operator(self, subspace='gef', polarization='x', transitions='-+'):
        raise NotImplementedError('%s has no dipole operators' % type(self).__name__)

    def system_bath_couplings(self, subspace='gef'):
        raise NotImplementedError('%s has no system-bath couplings' % type(self).__name__)

    # -- states ----------------------------------------------------------------
    @imemoize
    def ground_state(self, subspace):
        return ground_state(self.lab_frame.H(subspace))

    @imemoize
    def thermal_state(self, subspace):
        lab = self.lab_frame
        temperature = getattr(getattr(lab, 'bath', None), 'temperature', 0)
        return thermal_state(lab.H(subspace), temperature)

    # -- spectrum --------------------------------------------------------------
    @imemoize
    def eig(self, subspace):
        """(E, U): eigenvectors of the lab-frame matrix (which keeps the manifolds ordered
        g < e < f), energies moved to this variant's frame."""
        E, U = scipy.linalg.eigh(self.lab_frame.H(subspace))
        for letter, quanta in (('e', 1), ('f', 2)):
            if letter in subspace and self.rw_freq:
                E[self.hilbert_subspace_index(letter, subspace)] -= quanta * self.rw_freq
        return E, U

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
        """Nyquist bound of the un-sampled spectrum in this frame: all ensemble members share
        one time / frequency grid."""
        energies = self.unsampled.E('gef')
        extra = (0.01 * self._root.transition_energy if self.energy_spread_extra is None
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

    def _same_data(self, other):
        return (isinstance(other, ElectronicHamiltonian)
                and np.all(self.H_1exc == other.H_1exc)
                and self.bath == other.bath
                and np.all(self.dipoles == other.dipoles)
                and self.disorder == other.disorder
                and np.all(np.atleast_1d(self.random_seed) == np.atleast_1d(other.random_seed))
                and self.energy_spread_extra == other.energy_spread_extra
                and self.rw_freq == other.rw_freq
                and self.site_labels == other.site_labels)

    def __eq__(self, other):
        """Same matrices, bath, dipoles and frame, and descended from equal roots."""
        return bool(self._same_data(other) and self._root._same_data(other._root))

    def __ne__(self, other):
        return not self == other

    __hash__ = object.__hash__

    @property
    def n_sites(self):
        return len(self.H_1exc)

    @imemoize
    def H(self, subspace):
        return operator_extend(self.H_1exc, subspace)

    def disorder_stream(self, n):
        """The seeded generator of ensemble member ``n``."""
        return check_random_state(list(np.atleast_1d(self.random_seed)) + [n])

    def sampled_site_shifts(self, ensemble_size, member0=0):
        """(ensemble_size, n_sites) diagonal static-disorder shifts of members
        member0 .. member0+ensemble_size-1 -- bit-identical to what
        ``sample(n)`` adds -- or None when ``disorder`` is a user callable.
        Uses the C replay of numpy's seeded streams (qsx_sample_streams), so no
        Hamiltonian object is created per member."""
        base = self._root
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
        base = self._root
        if base.disorder is not None and not isinstance(base.disorder, Number):
            return None
        from . import _capi
        if base.disorder is None:
            torch = _capi.torch_cuda()
            return torch.zeros((ensemble_size, base.n_sites), dtype=torch.float64, device='cuda')
        return _capi.sample_gauss_device(base.random_seed, member0, ensemble_size,
                                         base.n_sites, base.disorder * GAUSSIAN_SD_FWHM)

    def _build(self, member, rw_freq):
        """Root -> variant: static disorder and orientation of member (n, random_orientations)
        from its seeded stream (the diagonal ``randn`` draw first, then ``rand(3)`` for the
        rotation, reference hamiltonian.py:552-578) added to the matrix in the requested frame
        (frame first: bit-identical with sampling a rotating-frame Hamiltonian, the order the
        dynamical models use)."""
        H, dipoles, seed = self.H_1exc, self.dipoles, self.random_seed
        if rw_freq:
            H = H - rw_freq * np.identity(self.n_sites)
        if member is not None:
            n, random_orientations = member
            rng = self.disorder_stream(n)
            if self.disorder is None:
                if not random_orientations:
                    warnings.warn('called sample with `disorder=None` and '
                                  '`random_orientations=False`: sampled Hamiltonian is '
                                  'identical to original', RuntimeWarning, stacklevel=4)
            elif isinstance(self.disorder, Number):
                H = H + diagonal_gaussian_disorder(self.disorder, self.n_sites)(rng)
            else:
                H = H + self.disorder(rng)
            if random_orientations:
                dipoles = np.einsum('mn,in->im', random_rotation_matrix(rng), dipoles)
            seed = list(np.atleast_1d(seed)) + [n]
        return type(self)(H, self.bath, dipoles, self.disorder, seed,
                          self.energy_spread_extra, self.site_labels)

    def dipole_operator(self, subspace='gef', polarization='x',
                        transitions='-+'):
        if self.dipoles is None:
            raise HamiltonianError('transition dipole moments undefined')
        mu = self.dipoles @ polarization_vector(polarization)
        ops = [transition_operator(n, self.n_sites, subspace, transitions)
               for n in range(self.n_sites)]
        return np.einsum('nij,n->ij', ops, mu)

    @imemoize
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

    def _build(self, member, rw_freq):
        electronic = self.electronic
        if member is not None:
            electronic = electronic.sample(*member)
        electronic = electronic.in_rotating_frame(rw_freq) if rw_freq else electronic.lab_frame
        return type(self)(electronic, self.n_vibrational_levels, self.vib_energies,
                          self.elec_vib_couplings, self.energy_spread_extra, self.site_labels)

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
