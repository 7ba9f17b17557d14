"""
ORACLE (test infrastructure) -- Liouville-space conventions.

Restates reference ``qspectra/dynamics/liouville_space.py``:
column-major vec :46-58, row-major bra-vec :61-65, subspace index :9-29,
tensor->super index law :68-90, left/right/commutator super-operators :93-131,
operator sub-blocks :151-209.
"""
import numpy as np

__all__ = ['ket_vec', 'unket_vec', 'bra_vec', 'subspace_index', 'super_left',
           'super_right', 'super_commutator', 'tensor_to_super_matrix',
           'operator_blocks', 'full_subspace', 'embed_state']

_ORDER = 'gef'


def ket_vec(matrix):
    """stacked columns (liouville_space.py:46-50)"""
    return np.asarray(matrix).reshape(-1, order='F')


def unket_vec(vec):
    """inverse of ket_vec for a square operator (liouville_space.py:53-58)"""
    n = int(round(np.sqrt(np.size(vec))))
    return np.asarray(vec).reshape((n, n), order='F')


def bra_vec(matrix):
    """stacked rows (liouville_space.py:61-65)"""
    return np.asarray(matrix).reshape(-1, order='C')


def _manifold_sizes(n_sites, n_vib):
    # operator_tools.py:279-285
    return {'g': n_vib, 'e': n_sites * n_vib,
            'f': (n_sites * (n_sites - 1) // 2) * n_vib}


def subspace_index(liouville_subspace, hilbert_subspace, n_sites, n_vib=1):
    """Sorted flat (column-major) positions of the kept blocks
    (liouville_space.py:9-29)."""
    sizes = _manifold_sizes(n_sites, n_vib)
    ranges, total = {}, 0
    for letter in _ORDER:
        if letter in hilbert_subspace:
            ranges[letter] = np.arange(total, total + sizes[letter])
            total += sizes[letter]
    keep = np.zeros((total, total), dtype=bool)
    for block in liouville_subspace.split(','):
        row, col = block
        if row not in ranges or col not in ranges:
            raise KeyError("%s not in subspace '%s'" % (block, hilbert_subspace))
        keep[np.ix_(ranges[row], ranges[col])] = True
    return np.flatnonzero(ket_vec(keep))


def full_subspace(subspace_string):
    """All blocks spanned by the letters present (operator_tools.py:288-308)."""
    letters = sorted(set(subspace_string) - set(',->'), key=_ORDER.index)
    return ','.join(a + b for a in letters for b in letters)


def super_left(op):
    """vec(A rho) = (I (x) A) vec(rho)  (liouville_space.py:103-116)"""
    return np.kron(np.identity(len(op)), op)


def super_right(op):
    """vec(rho A) = (A^T (x) I) vec(rho)  (liouville_space.py:119-131)"""
    return np.kron(np.asarray(op).T, np.identity(len(op)))


def super_commutator(op):
    return super_left(op) - super_right(op)


def tensor_to_super_matrix(R):
    """S[i + N j, k + N l] = R[i, j, k, l]  (liouville_space.py:68-90)"""
    N = R.shape[0]
    return R.transpose(1, 0, 3, 2).reshape(N * N, N * N)


def operator_blocks(op, from_idx, to_idx):
    """(left, right, commutator, bra_vector, expectation row) of a Hilbert
    operator restricted to Liouville index sets (liouville_space.py:162-209)."""
    mesh = np.ix_(to_idx, from_idx)
    left = super_left(op)[mesh]
    right = super_right(op)[mesh]
    bra = bra_vec(np.asarray(op, dtype=complex))[from_idx]
    tr = np.identity(len(op)).reshape(-1)[to_idx]
    return left, right, left - right, bra, tr @ left


def embed_state(state, from_idx, to_idx, n_states):
    """Scatter into the full N^2 vector then gather (liouville_space.py:343-349)."""
    full = np.zeros(n_states * n_states, dtype=complex)
    full[from_idx] = state
    return full[to_idx]
