"""
ORACLE (test infrastructure) -- equations of motion.

Restates, in array form, the reference generators:
* Redfield tensor / evolve: ``qspectra/dynamics/redfield.py:9-104``
* unitary: ``qspectra/dynamics/unitary.py:5-9``
* HEOM hierarchy: ``qspectra/dynamics/heom.py:61-174`` (Matsubara data, ADO
  enumeration) and ``:298-443`` (generator), assembled here as one scipy COO ->
  CSR matrix from Kronecker blocks instead of per-block ``lil_matrix``
  assignment (same matrix; compared entry-by-entry with the reference in
  tests/golden).
* ZOFE right-hand side: ``qspectra/dynamics/zofe.py:121-234``.
"""
import numpy as np
import scipy.sparse as sp

from .liouville import (super_commutator, super_left, super_right,
                        tensor_to_super_matrix)

__all__ = ['redfield_tensor', 'redfield_generator', 'unitary_generator',
           'matsubara_frequencies', 'corr_func_coeffs', 'multichoose',
           'ado_table', 'ado_neighbours', 'heom_generator', 'zofe_rhs',
           'debye_corr_complex', 'bath_corr_real']


# ------------------------------------------------------------------ baths
def debye_corr_complex(x, T, lam, gamma, cutoff=1000):
    """bath.py:84-102"""
    if x == 0:
        return lam * (2 * T / gamma - 1j)
    nu = 2 * np.pi * np.arange(cutoff) * T
    return (lam * gamma
            * ((1 / np.tan(gamma / (2 * T)) - 1j) / (gamma - 1j * x)
               + 4 * T * np.sum(nu / ((nu ** 2 - gamma ** 2) * (nu - 1j * x)))))


def bath_corr_real(x, T, J, J0):
    """bath.py:17-31"""
    if x == 0:
        return T * J0
    J_anti = J(x) if x >= 0 else -J(-x)
    return (1 / np.expm1(x / T) + 1) * J_anti


# --------------------------------------------------------------- Redfield
def redfield_tensor(E, U, couplings, corr_func, secular=True):
    """4-index Redfield tensor in the eigenbasis (redfield.py:9-74).

    E, U      : eigen-energies / eigenvectors of the subspace Hamiltonian
    couplings : (n_baths, N, N) system-bath operators in the site basis
    corr_func : scalar callable C(omega)
    """
    N = len(E)
    K = np.array([U.conj().T @ V @ U for V in couplings])        # :44-45
    C = np.array([[corr_func(Ei - Ej) for Ej in E] for Ei in E])  # :53
    Gamma = np.einsum('iab,icd,dc->abcd', K, K, C)                # :61
    I = np.identity(N)
    Gs = np.einsum('abbc->ac', Gamma)
    R = (np.einsum('ac,bd->abcd', I, Gs).conj()
         + np.einsum('bd,ac->abcd', I, Gs)
         - np.einsum('cabd->abcd', Gamma).conj()
         - np.einsum('dbac->abcd', Gamma))                        # :64-69
    if secular:
        Ib = np.identity(N, dtype=bool)
        R = R * (np.einsum('ab,cd->abcd', Ib, Ib)
                 | np.einsum('ac,bd->abcd', Ib, Ib))              # :77-83
    return R


def redfield_generator(E, U, couplings, corr_func, secular=True,
                       basis='site'):
    """L = -i [diag(E), .] - R, optionally rotated to the site basis with
    W = kron(U^+, U^+) as W^+ L W (redfield.py:95-104, operator_tools.py:43-87)."""
    R = tensor_to_super_matrix(redfield_tensor(E, U, couplings, corr_func,
                                               secular))
    L = -1j * super_commutator(np.diag(E)) - R
    if basis == 'eigen':
        return L
    if basis != 'site':
        raise ValueError('invalid basis')
    Ud = U.T.conj()
    W = np.kron(Ud, Ud)
    return W.T.conj() @ L @ W


def unitary_generator(H):
    """unitary.py:5-9 (unit_convert already folded into H by the caller)"""
    return -1j * super_commutator(H)


# ------------------------------------------------------------------- HEOM
def matsubara_frequencies(K, gamma, T):
    """heom.py:61-67"""
    v = 2 * np.pi * T * np.arange(K + 1)
    v[0] = gamma
    return v


def corr_func_coeffs(K, gamma, T, lam, nu):
    """heom.py:69-89 (aki_temp_corr=False branch)"""
    c = [lam * gamma * (1 / np.tan(gamma / (2 * T)) - 1j)]
    for k in range(1, K + 1):
        c.append(4 * lam * gamma * T * nu[k] / (nu[k] ** 2 - gamma ** 2))
    return c


def multichoose(n, c):
    """All ways to put c balls in n bins, in the reference's recursion order
    (heom.py:155-174): lexicographically ascending."""
    if not c:
        return [[0] * n]
    if not n:
        return []
    if n == 1:
        return [[c]]
    return ([[0] + v for v in multichoose(n - 1, c)]
            + [[v[0] + 1] + v[1:] for v in multichoose(n, c - 1)])


def ado_table(n_sites, K, level_cutoff):
    """(n_ado, n_sites*(K+1)) int64 table, levels concatenated (heom.py:92-152)."""
    bins = n_sites * (K + 1)
    rows = []
    for c in range(level_cutoff):
        rows.extend(multichoose(bins, c))
    return np.array(rows, dtype=np.int64).reshape(len(rows), bins)


def ado_neighbours(table):
    """up[n, b] / down[n, b] = index of table[n] +- e_b, or -1
    (heom.py:411-418: ``mat_to_ind`` returns None when absent)."""
    lookup = {tuple(v): i for i, v in enumerate(table.tolist())}
    n_ado, bins = table.shape
    up = -np.ones((n_ado, bins), dtype=np.int64)
    down = -np.ones((n_ado, bins), dtype=np.int64)
    for n, v in enumerate(table.tolist()):
        for b in range(bins):
            v[b] += 1
            up[n, b] = lookup.get(tuple(v), -1)
            v[b] -= 2
            down[n, b] = lookup.get(tuple(v), -1)
            v[b] += 1
    return up, down


def heom_generator(H, couplings, idx, gamma, T, lam, K, level_cutoff,
                   low_temp_corr=True, modified=False):
    """Sparse hierarchy generator on the concatenated ADO state
    (heom.py:298-443; ``aki_temp_corr`` unsupported -- reference quirk, it
    multiplies by a matrix instead of the temperature, heom.py:362 vs :405).

    H, couplings are given in the Hilbert subspace; ``idx`` selects the
    Liouville subspace (ADO n occupies y[n*M:(n+1)*M], heom.py:371-372).
    """
    n_sites = len(couplings)
    mesh = np.ix_(idx, idx)
    M = len(idx)
    nu = matsubara_frequencies(K, gamma, T)
    c = corr_func_coeffs(K, gamma, T, lam, nu)
    table = ado_table(n_sites, K, level_cutoff)
    up, down = ado_neighbours(table)
    n_ado = len(table)

    comm_H = super_commutator(H)[mesh]                                  # :375-376
    PL = [super_left(V)[mesh] for V in couplings]                       # :382-385
    PR = [super_right(V)[mesh] for V in couplings]

    nu_inf = matsubara_frequencies(K + 5000, gamma, T)                  # :387-393
    c_inf = np.array(corr_func_coeffs(K + 5000, gamma, T, lam, nu_inf))
    tc = np.sum((c_inf / nu_inf)[K + 1:])

    diag_block = -1j * comm_H
    if low_temp_corr:                                                   # :403-408
        dbl = sum(pl + pr - 2 * pl @ pr for pl, pr in zip(PL, PR))
        diag_block = diag_block - tc * dbl

    shifts = table.reshape(n_ado, n_sites, K + 1) @ nu                  # :399
    shifts = shifts.sum(axis=1)
    blocks = [sp.kron(sp.identity(n_ado), sp.csr_matrix(diag_block)),
              sp.kron(sp.diags(-shifts), sp.identity(M))]
    for j in range(n_sites):
        for k in range(K + 1):
            b = j * (K + 1) + k
            n_jk = table[:, b]
            # coupling to the ADO one level up                            :419-427
            rows = np.flatnonzero(up[:, b] >= 0)
            coef = (np.sqrt((n_jk[rows] + 1) * abs(c[k])) if modified
                    else np.ones(len(rows)))
            S = sp.csr_matrix((coef, (rows, up[rows, b])), shape=(n_ado, n_ado))
            blocks.append(sp.kron(S, sp.csr_matrix(-1j * (PL[j] - PR[j]))))
            # coupling to the ADO one level down                          :429-439
            rows = np.flatnonzero(down[:, b] >= 0)
            coef = (np.sqrt(n_jk[rows] / abs(c[k])) if modified
                    else n_jk[rows].astype(float))
            S = sp.csr_matrix((coef, (rows, down[rows, b])),
                              shape=(n_ado, n_ado))
            blocks.append(sp.kron(S, sp.csr_matrix(
                -1j * (c[k] * PL[j] - np.conj(c[k]) * PR[j]))))
    return sum(blocks).tocsr()


# ------------------------------------------------------------------- ZOFE
def zofe_rhs(vec, H, Ln, Gamma, w, ham_hermit=False, rho_hermit=False):
    """d/dt [vec_F(rho); vec_F(O)] (zofe.py:121-202).  ``Ln`` already carries
    the reference's sign flip L_n = -V_n (zofe.py:216-217); Gamma = Omega^2
    huang, w = i Omega + gamma, both (n_pm, n_sites)."""
    P, S = Gamma.shape
    n = len(H)
    rho = vec[:n * n].reshape((n, n), order='F')
    O = vec[n * n:].reshape((P, S, n, n), order='F')
    Os = O.sum(axis=0)
    Ld = Ln.swapaxes(1, 2).conj()
    a = np.einsum('sab,sbc->ac', Ld, Os)
    b = -1j * H - a
    c = np.einsum('sab,bc,sdc->ad', Ln, rho, Os.conj())
    d = b @ rho + c
    if not rho_hermit:
        big = np.einsum('sab,bc,scd->ad', Os, rho, Ld)
        f = (rho @ b.T.conj() if ham_hermit
             else rho @ (1j * H - a.T.conj())) + big
    else:
        f = (d.T.conj() if ham_hermit
             else rho @ (1j * H - a.T.conj()) + c.T.conj())
    rhodot = d + f
    Odot = (Gamma[:, :, None, None] * Ln[None]
            - w[:, :, None, None] * O
            + np.einsum('ab,psbc->psac', b, O)
            - np.einsum('psab,bc->psac', O, b))
    return np.append(rhodot.reshape(-1, order='F'), Odot.reshape(-1, order='F'))
