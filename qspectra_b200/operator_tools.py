"""
Hilbert-space bookkeeping for hard-core-boson exciton manifolds.

Behavioural contract = reference ``qspectra/operator_tools.py`` (state order
:122-152, 1->2 excitation lift :155-191, block extension :194-235, transition
operators :238-270, subspace helpers :279-335, basis transforms :59-119).  The
implementation here is bit-mask based and vectorised; integer outputs are
compared bit-exactly against the reference's golden vectors in
``tests/test_host_foundation.py``.
"""
from functools import reduce
from itertools import combinations

import functools

import numpy as np

_EXC_ORDER = 'gef'


class SubspaceError(Exception):
    """Raised for an invalid Hilbert/Liouville subspace request."""


# --------------------------------------------------------------------------
# vibrational helpers
# --------------------------------------------------------------------------
def tensor(*ops):
    return reduce(np.kron, ops)


def extend_vib_operator(n_vibrational_levels, m, vib_operator):
    """Embed an operator of vibrational mode ``m`` in the full vibrational space."""
    levels = np.asarray(n_vibrational_levels)
    left = int(np.prod(levels[:m]))
    right = int(np.prod(levels[m + 1:]))
    return tensor(np.eye(left), vib_operator, np.eye(right))


def vib_annihilate(N):
    return np.diag(np.sqrt(np.arange(1, N)), k=1)


def vib_create(N):
    return np.diag(np.sqrt(np.arange(1, N)), k=-1)


def unit_vec(n, N, dtype=complex):
    v = np.zeros(N, dtype=dtype)
    v[n] = 1
    return v


# --------------------------------------------------------------------------
# basis transforms
# --------------------------------------------------------------------------
def _lift_transform(last_dim, U):
    U = np.asarray(U)
    if U.ndim != 2 or U.shape[0] != U.shape[1]:
        raise ValueError('basis transformation must be a square matrix')
    n = U.shape[0]
    if last_dim == n:
        return U
    if last_dim == n * n:
        return np.kron(U, U)
    raise ValueError('basis transformation incompatible with operator '
                     'dimensions')


def basis_transform_operator(X, U):
    """``U^dagger X U`` for Hilbert operators; ``kron(U,U)`` is used for
    Liouville-space super-operators."""
    X = np.asarray(X)
    if X.ndim != 2:
        raise ValueError('operator must have ndim=2')
    W = _lift_transform(X.shape[-1], U)
    return W.conj().T @ X @ W


def basis_transform_vector(rho, U):
    """Transform a (batch of) state vector(s) on the last axis."""
    rho = np.asarray(rho)
    W = _lift_transform(rho.shape[-1], U)
    return np.tensordot(rho, W.conj().T, axes=(-1, -1))


# --------------------------------------------------------------------------
# state enumeration
# --------------------------------------------------------------------------
def all_states(N, subspace='gef'):
    """Occupied-site lists: g = [], e = [i], f = [i, j] with i < j, in that
    order (lexicographic inside each manifold)."""
    states = []
    for n_exc, letter in enumerate(_EXC_ORDER):
        if letter in subspace:
            states.extend(list(c) for c in combinations(range(N), n_exc))
    return states


@functools.lru_cache(maxsize=None)
def _state_masks(N, subspace):
    masks = np.array([sum(1 << s for s in st) for st in all_states(N, subspace)],
                     dtype=np.int64)
    masks.setflags(write=False)
    return masks


def operator_1_to_2(operator1):
    """Lift ``sum_nm A_nm a+_n a_m`` from the 1- to the 2-excitation manifold."""
    A = np.asarray(operator1)
    pairs = all_states(len(A), 'f')
    out = np.zeros((len(pairs), len(pairs)), dtype=A.dtype)
    for m, (i, j) in enumerate(pairs):
        for n, (k, l) in enumerate(pairs):
            out[m, n] = (A[j, l] * (i == k) + A[j, k] * (i == l)
                         + A[i, l] * (j == k) + A[i, k] * (j == l))
    return out


def operator_extend(operator1, subspace='gef'):
    """Block-diagonal extension of a 1-excitation operator to g/e/f blocks."""
    A = np.asarray(operator1)
    blocks = []
    if 'g' in subspace:
        blocks.append(np.zeros((1, 1), dtype=A.dtype))
    if 'e' in subspace:
        blocks.append(A)
    if 'f' in subspace:
        blocks.append(operator_1_to_2(A))
    size = sum(len(b) for b in blocks)
    out = np.zeros((size, size), dtype=A.dtype)
    pos = 0
    for b in blocks:
        out[pos:pos + len(b), pos:pos + len(b)] = b
        pos += len(b)
    return out


def transition_operator(n, n_sites, subspace='gef', include_transitions='-+'):
    """0/1 matrix of a+_n ('+') and/or a_n ('-') between the listed states (a fresh array;
    the construction is cached: the response functions ask for the same few operators once
    per polarisation configuration and pathway)."""
    return _transition_operator(int(n), int(n_sites), str(subspace), str(include_transitions)).copy()


@functools.lru_cache(maxsize=4096)
def _transition_operator(n, n_sites, subspace, include_transitions):
    masks = _state_masks(n_sites, subspace)
    bit = 1 << n
    has = (masks & bit) != 0
    # row state == column state with site n added
    raises = (masks[:, None] == (masks[None, :] | bit)) & ~has[None, :]
    out = np.zeros((len(masks), len(masks)))
    if '+' in include_transitions:
        out[raises] = 1
    if '-' in include_transitions:
        out[raises.T] = 1
    return out


# --------------------------------------------------------------------------
# subspace helpers
# --------------------------------------------------------------------------
def n_excitations(n_sites=1, n_vibrational_states=1):
    counts = np.array([1, n_sites, int(n_sites * (n_sites - 1) / 2)])
    return counts * n_vibrational_states


def excitation_to_number(excitation):
    return _EXC_ORDER.index(excitation)


def extract_subspace(subspaces_string):
    letters = set(subspaces_string) - set(',->')
    return sorted(letters, key=excitation_to_number)


def full_liouville_subspace(subspaces_string):
    letters = extract_subspace(subspaces_string)
    return ','.join(a + b for a in letters for b in letters)


def hilbert_subspace_index(subspace, all_subspaces, n_sites,
                           n_vibrational_states=1):
    """slice selecting manifold ``subspace`` inside the ordered ``all_subspaces``."""
    if subspace not in all_subspaces:
        raise SubspaceError("{} not in set of all subspaces '{}'".format(
            subspace, all_subspaces))
    counts = n_excitations(n_sites, n_vibrational_states)
    sizes = [int(counts[_EXC_ORDER.index(s)]) for s in all_subspaces]
    k = all_subspaces.index(subspace)
    start = sum(sizes[:k])
    return slice(start, start + sizes[k])
