"""
Liouville-space conventions, dipole super-operators and the dense-generator
model base class.

Contract: reference ``qspectra/dynamics/liouville_space.py`` -- subspace index
:9-29 (integer maps, bit-exact), vectorisation :46-65, tensor->super index law
:68-90, super-operators :93-131, ``LiouvilleSpaceOperator`` :151-209,
``LiouvilleSpaceModel`` :212-349.  The equation of motion itself runs on the
GPU (``engine.DenseEOM`` -> csrc/dense.cu).
"""
import itertools

import numpy as np

from .base import DynamicalModel, SystemOperator
from ..engine import DenseEOM, LinearMap
from ..operator_tools import (SubspaceError, n_excitations,
                              full_liouville_subspace)
from ..utils import _CACHE_ATTR, imemoize, memoized_property


# ------------------------------------------------------------- conventions
def matrix_to_ket_vec(matrix):
    """stacked columns"""
    return np.asarray(matrix).reshape(-1, order='F')


def ket_vec_to_matrix(ket_vec):
    ket_vec = np.asarray(ket_vec)
    N = int(np.sqrt(np.prod(ket_vec.shape)))
    return ket_vec.reshape((N, N), order='F')


def matrix_to_bra_vec(matrix):
    """stacked rows"""
    return np.asarray(matrix).reshape(-1, order='C')


def all_liouville_subspaces(hilbert_subspace):
    return ','.join(a + b for a, b in itertools.product(hilbert_subspace,
                                                        repeat=2))


def liouville_subspace_index(liouville_subspace, full_subspace, n_sites,
                             n_vibrational_states=1):
    """Sorted flat (column-major) positions, inside the vectorised operator on
    ``full_subspace``, of the blocks named by e.g. 'eg,fe'."""
    counts = n_excitations(n_sites, n_vibrational_states)
    manifolds, total = {}, 0
    for letter, size in zip('gef', counts):
        if letter in full_subspace:
            manifolds[letter] = (total, total + int(size))
            total += int(size)
    keep = np.zeros((total, total), dtype=bool)
    for block in liouville_subspace.split(','):
        try:
            row, col = block
            (r0, r1), (c0, c1) = manifolds[row], manifolds[col]
        except (KeyError, ValueError):
            raise SubspaceError("{} not in subspace '{}'".format(
                block, full_subspace))
        keep[r0:r1, c0:c1] = True
    return np.flatnonzero(keep.reshape(-1, order='F'))


def transposition_permutation(index, n_states):
    """perm[k] = position, inside the sorted flat (column-major) subspace ``index``, of the
    transposed ket-bra pair of element k; None if the subspace is not closed under
    transposition ('eg', 'fe', ...).  Physical generators commute with Hermitian conjugation,
    L[perm r][perm c] = conj L[r][c], which the engine uses to propagate Hermitian states in
    real coordinates (engine.DenseEOM.hermitian_perm, csrc/dense_real.cu)."""
    index = np.asarray(index)
    a, b = index % n_states, index // n_states
    mirrored = b + n_states * a
    perm = np.searchsorted(index, mirrored)
    if np.any(perm >= index.size) or np.any(index[np.minimum(perm, index.size - 1)] != mirrored):
        return None
    return perm.astype(np.int32)


def tensor_to_super(tensor_operator):
    """R[i, j, k, l] -> S[i + N j, k + N l]"""
    R = np.asarray(tensor_operator)
    N = R.shape[0]
    return np.ascontiguousarray(R.transpose(1, 0, 3, 2)).reshape(N * N, N * N)


def super_left_matrix(operator):
    """vec(A rho) = (I (x) A) vec(rho)"""
    operator = np.asarray(operator)
    return np.kron(np.identity(len(operator)), operator)


def super_right_matrix(operator):
    """vec(rho A) = (A^T (x) I) vec(rho)"""
    operator = np.asarray(operator)
    return np.kron(operator.T, np.identity(len(operator)))


def super_commutator_matrix(operator):
    return super_left_matrix(operator) - super_right_matrix(operator)


def _block(kind, operator, to_idx, from_idx):
    """Sub-block [to, from] of a left/right super-operator without forming the
    full N^2 x N^2 Kronecker product."""
    op = np.asarray(operator)
    N = len(op)
    ta, tb = to_idx % N, to_idx // N
    fa, fb = from_idx % N, from_idx // N
    if kind == 'left':      # (I (x) A)[(a,b),(c,d)] = A[a,c] delta(b,d)
        return op[np.ix_(ta, fa)] * (tb[:, None] == fb[None, :])
    # (A^T (x) I)[(a,b),(c,d)] = A[d,b] delta(a,c)
    return op[np.ix_(fb, tb)].T * (ta[:, None] == fa[None, :])


class LiouvilleSpaceOperator(SystemOperator):
    """A Hilbert-space operator acting between Liouville subspaces
    ('eg,fe->gg,ee' or a single subspace)."""

    def __init__(self, operator, liouv_subspace_map, dynamical_model):
        self.operator = np.asarray(operator)
        parts = (liouv_subspace_map.split('->') if '->' in liouv_subspace_map
                 else [liouv_subspace_map, liouv_subspace_map])
        self.from_indices, self.to_indices = [
            dynamical_model.liouville_subspace_index(p) for p in parts]

    @property
    def bra_vector(self):
        op = np.asarray(self.operator, dtype=complex)
        return matrix_to_bra_vec(op)[self.from_indices]

    @memoized_property
    def _left(self):
        return _block('left', self.operator, self.to_indices, self.from_indices)

    @memoized_property
    def _right(self):
        return _block('right', self.operator, self.to_indices, self.from_indices)

    @memoized_property
    def left_multiply(self):
        return LinearMap(self._left)

    @memoized_property
    def right_multiply(self):
        return LinearMap(self._right)

    @memoized_property
    def commutator(self):
        return LinearMap(self._left - self._right)

    @memoized_property
    def expectation_value(self):
        # tr(M rho) = sum_i vec(I)_i (I (x) M)_ij vec(rho)_j
        N = len(self.operator)
        tr = np.identity(N).reshape(-1)[self.to_indices]
        return LinearMap(tr.dot(self._left))


class LiouvilleSpaceModel(DynamicalModel):
    """Dense-generator models.  Subclasses provide
    ``evolution_super_operator`` (full N^2 x N^2 array on the host)."""
    system_operator = LiouvilleSpaceOperator

    def __init__(self, hamiltonian, rw_freq=None, hilbert_subspace='gef',
                 unit_convert=1, evolve_basis='site', sparse_matrix=False):
        super(LiouvilleSpaceModel, self).__init__(hamiltonian, rw_freq,
                                                  hilbert_subspace,
                                                  unit_convert)
        self.evolve_basis = evolve_basis
        # accepted for signature compatibility; generators are dense on the GPU
        self.sparse_matrix = sparse_matrix

    @property
    def evolve_basis(self):
        return self._evolve_basis

    @evolve_basis.setter
    def evolve_basis(self, val):
        if val not in ('site', 'eigen'):
            raise ValueError('invalid basis')
        self._evolve_basis = val

    # -- states --------------------------------------------------------------
    def liouville_subspace_index(self, subspace):
        return liouville_subspace_index(
            subspace, self.hilbert_subspace, self.hamiltonian.n_sites,
            self.hamiltonian.n_vibrational_states)

    def map_between_subspaces(self, state, from_subspace, to_subspace):
        f_idx, t_idx = [self.liouville_subspace_index(s)
                        for s in (from_subspace, to_subspace)]
        N = self.hamiltonian.n_states(self.hilbert_subspace)
        full = np.zeros(N * N, dtype=complex)
        full[f_idx] = state
        return full[t_idx]

    def density_matrix_to_state_vector(self, rho0, liouville_subspace):
        return self.map_between_subspaces(
            matrix_to_ket_vec(rho0),
            full_liouville_subspace(liouville_subspace), liouville_subspace)

    def state_vector_to_density_matrix(self, rho):
        rho = np.asarray(rho)
        N = int(np.sqrt(rho.shape[-1]))
        return rho.reshape(-1, N, N, order='F')

    def thermal_state(self, liouville_subspace):
        # NB reference quirk 11: the Liouville string is handed to the
        # Hamiltonian, which only looks for the letters g/e/f in it.
        rho0 = self.hamiltonian.thermal_state(liouville_subspace)
        rho = self.map_between_subspaces(
            matrix_to_ket_vec(rho0),
            full_liouville_subspace(liouville_subspace), liouville_subspace)
        if self.evolve_basis == 'eigen':
            rho = self.hamiltonian.transform_vector_to_eigenbasis(
                rho, liouville_subspace)
        return rho

    def dipole_operator(self, liouv_subspace_map, polarization,
                        transitions='-+'):
        # the response functions ask for the same few operators once per polarisation
        # configuration and pathway: keep them per model (hashable polarisations only)
        try:
            key = ('dipole_operator', liouv_subspace_map, polarization, transitions)
            # the instance cache of utils.imemoize: dropped by copy_with_new_cache (sampled members)
            memo = self.__dict__.setdefault(_CACHE_ATTR, {})
            return memo[key]
        except KeyError:
            memo[key] = op = self._dipole_operator(liouv_subspace_map, polarization, transitions)
            return op
        except TypeError:           # array-valued polarisation
            return self._dipole_operator(liouv_subspace_map, polarization, transitions)

    def _dipole_operator(self, liouv_subspace_map, polarization, transitions):
        operator = self.hamiltonian.dipole_operator(self.hilbert_subspace,
                                                    polarization, transitions)
        if self.evolve_basis == 'eigen':
            operator = self.hamiltonian.transform_operator_to_eigenbasis(
                operator, self.hilbert_subspace)
        return self.system_operator(operator, liouv_subspace_map, self)

    # -- generators ----------------------------------------------------------
    @property
    def evolution_super_operator(self):
        raise NotImplementedError('subclass must implement the property '
                                  '`evolution_super_operator`')

    def generator(self, liouville_subspace):
        """L[np.ix_(index, index)] on the host (complex128, M x M)."""
        index = self.liouville_subspace_index(liouville_subspace)
        return np.asarray(self.evolution_super_operator)[np.ix_(index, index)]

    @imemoize
    def equation_of_motion(self, liouville_subspace, heisenberg_picture=False):
        """DeviceEOM for dy/dt = L y; the Heisenberg picture uses the plain
        transpose L^T (reference liouville_space.py:325-330)."""
        return self.with_transposition(
            DenseEOM(self.generator(liouville_subspace), heisenberg_picture),
            liouville_subspace, heisenberg_picture)

    def with_transposition(self, eom, liouville_subspace, heisenberg_picture=False):
        """Tell a dense device generator how its subspace transposes (Schroedinger picture only)."""
        if not heisenberg_picture and isinstance(eom, DenseEOM):
            eom.hermitian_perm = transposition_permutation(
                self.liouville_subspace_index(liouville_subspace),
                self.hamiltonian.n_states(self.hilbert_subspace))
        return eom

    def ensemble_generators(self, members, liouville_subspace):
        """(n_members, M, M) stack of the members' generators."""
        return np.array([m.generator(liouville_subspace) for m in members])

    def ensemble_equation_of_motion(self, members, liouville_subspace,
                                    heisenberg_picture=False):
        if len(members) == 1:
            return members[0].equation_of_motion(liouville_subspace,
                                                 heisenberg_picture)
        return self.with_transposition(
            DenseEOM(self.ensemble_generators(members, liouville_subspace), heisenberg_picture),
            liouville_subspace, heisenberg_picture)
