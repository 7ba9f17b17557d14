"""
Redfield theory generators (secular / non-secular).

Contract: reference ``qspectra/dynamics/redfield.py`` -- ``redfield_tensor``
:9-74 (May & Kuhn eqs. 3.319/3.322 with the one-sided correlation spectrum),
``secular_terms`` :77-83, ``redfield_evolve`` :95-104, ``RedfieldModel``
:107-152.  The tensor contraction is written here for a *batch* of ensemble
members at once (``redfield_generators``); propagation of the resulting
Liouvillians runs on the GPU (csrc/dense.cu).
"""
import numpy as np

from .liouville_space import LiouvilleSpaceModel, tensor_to_super
from .. import engine, _capi
from ..bath import DebyeBath
from ..hamiltonian import ElectronicHamiltonian
from ..utils import memoized_property


def secular_terms(n_states):
    """mask of R[a,b,c,d] kept by the secular/Bloch approximation:
    population transfer (a==b, c==d) and coherence decay (a==c, b==d)"""
    eye = np.identity(n_states, dtype=bool)
    return (eye[:, :, None, None] & eye[None, None, :, :]) | \
           (eye[:, None, :, None] & eye[None, :, None, :])


def _correlation_matrix(corr_func, E):
    """C[..., i, j] = corr_func(E_i - E_j) for a batch of spectra."""
    gaps = E[..., :, None] - E[..., None, :]
    flat = gaps.reshape(-1)
    return np.array([corr_func(x) for x in flat],
                    dtype=complex).reshape(gaps.shape)


def redfield_tensors(E, U, couplings, corr_func, secular=True):
    """Redfield tensors R[m, a, b, c, d] (eigenbasis) for a batch of members.

    E (m, N), U (m, N, N): eigen-systems; couplings (n_baths, N, N) site-basis
    system-bath operators shared by all members."""
    E = np.asarray(E)
    U = np.asarray(U)
    V = np.asarray(couplings)
    N = E.shape[-1]
    K = np.einsum('mxa,ixy,myb->miab', U.conj(), V, U)
    Cw = _correlation_matrix(corr_func, E)
    # Gamma[a,b,c,d] = sum_i K_i[a,b] K_i[c,d] C[d,c]
    KC = K * Cw.transpose(0, 2, 1)[:, None]           # K_i[c,d] * C[d,c]
    Gamma = np.einsum('miab,micd->mabcd', K, KC)
    Gs = np.einsum('mabbc->mac', Gamma)
    eye = np.identity(N)
    R = (np.einsum('ac,mbd->mabcd', eye, Gs.conj())
         + np.einsum('bd,mac->mabcd', eye, Gs)
         - np.einsum('mcabd->mabcd', Gamma).conj()
         - np.einsum('mdbac->mabcd', Gamma))
    if secular:
        R = R * secular_terms(N)
    return R


def redfield_generators(E, U, couplings, corr_func, secular=True,
                        evolve_basis='site'):
    """L[m] = -i [diag(E_m), .] - R_m as (m, N^2, N^2) super-operators,
    optionally rotated to the site basis (redfield.py:95-104: W^+ L W with
    W = kron(U^+, U^+))."""
    if evolve_basis not in ('site', 'eigen'):
        raise ValueError('invalid basis')
    R = redfield_tensors(E, U, couplings, corr_func, secular)
    m, N = E.shape
    S = np.ascontiguousarray(R.transpose(0, 2, 1, 4, 3)).reshape(m, N * N, N * N)
    # -i (I (x) diag(E) - diag(E) (x) I): diagonal with entries E_a - E_b at a + N b
    gaps = (E[:, :, None] - E[:, None, :]).transpose(0, 2, 1).reshape(m, N * N)
    L = -S
    idx = np.arange(N * N)
    L[:, idx, idx] += -1j * gaps
    if evolve_basis == 'eigen':
        return L
    Ud = U.conj().transpose(0, 2, 1)
    W = np.einsum('mac,mbd->mabcd', Ud, Ud).reshape(m, N * N, N * N)
    return np.matmul(np.matmul(W.conj().transpose(0, 2, 1), L), W)


def redfield_tensor(hamiltonian, subspace='ge', secular=True,
                    discard_imag_corr=False):
    """Single-system 4-index tensor, same signature as the reference's."""
    corr = (hamiltonian.bath.corr_func_real if discard_imag_corr
            else hamiltonian.bath.corr_func_complex)
    return redfield_tensors(hamiltonian.E(subspace)[None],
                            hamiltonian.U(subspace)[None],
                            hamiltonian.system_bath_couplings(subspace),
                            corr, secular)[0]


def redfield_dissipator(*args, **kwargs):
    return tensor_to_super(redfield_tensor(*args, **kwargs))


def redfield_evolve(hamiltonian, subspace='ge', evolve_basis='site',
                    secular=True, discard_imag_corr=False):
    corr = (hamiltonian.bath.corr_func_real if discard_imag_corr
            else hamiltonian.bath.corr_func_complex)
    return redfield_generators(hamiltonian.E(subspace)[None],
                               hamiltonian.U(subspace)[None],
                               hamiltonian.system_bath_couplings(subspace),
                               corr, secular, evolve_basis)[0]


class RedfieldModel(LiouvilleSpaceModel):
    """DynamicalModel for Redfield theory; identical independent baths."""

    def __init__(self, hamiltonian, rw_freq=None, hilbert_subspace='gef',
                 unit_convert=1, secular=True, discard_imag_corr=False,
                 evolve_basis='site', sparse_matrix=False):
        super(RedfieldModel, self).__init__(hamiltonian, rw_freq,
                                            hilbert_subspace, unit_convert,
                                            evolve_basis, sparse_matrix)
        self.secular = secular
        self.discard_imag_corr = discard_imag_corr

    @memoized_property
    def evolution_super_operator(self):
        return self.unit_convert * redfield_evolve(
            self.hamiltonian, self.hilbert_subspace,
            evolve_basis=self.evolve_basis, secular=self.secular,
            discard_imag_corr=self.discard_imag_corr)

    def ensemble_generators(self, members, liouville_subspace, chunk=None):
        """Batched host construction of the members' generators restricted to
        the subspace (the reference rebuilds them one by one, base.py:120-128)."""
        ss = self.hilbert_subspace
        index = self.liouville_subspace_index(liouville_subspace)
        N = self.hamiltonian.n_states(ss)
        if chunk is None:
            chunk = max(1, int(2 ** 24 // max(1, N ** 4)))
        corr = (self.hamiltonian.bath.corr_func_real if self.discard_imag_corr
                else self.hamiltonian.bath.corr_func_complex)
        V = self.hamiltonian.system_bath_couplings(ss)
        out = np.empty((len(members), index.size, index.size), dtype=complex)
        for lo in range(0, len(members), chunk):
            part = members[lo:lo + chunk]
            E = np.array([m.hamiltonian.E(ss) for m in part])
            U = np.array([m.hamiltonian.U(ss) for m in part])
            L = redfield_generators(E, U, V, corr, self.secular,
                                    self.evolve_basis)
            out[lo:lo + chunk] = self.unit_convert * L[:, index[:, None],
                                                       index[None, :]]
        return out

    # -- device-side ensemble construction (kernel K5) ---------------------------
    def _device_buildable(self):
        ham = self.hamiltonian
        return (isinstance(ham, ElectronicHamiltonian)
                and type(ham.bath) is DebyeBath
                and not np.iscomplexobj(ham.H_1exc))

    def _tensor_device_buildable(self):
        """K5 with host-supplied eigensystems: any Hamiltonian class (vibronic, complex)
        with a Drude-Lorentz bath and diagonal system-bath operators, up to 64 states."""
        ham, ss = self.hamiltonian, self.hilbert_subspace
        if type(ham.bath) is not DebyeBath or ham.n_states(ss) > 64:
            return False
        V = np.asarray(ham.system_bath_couplings(ss))
        diag = np.einsum('jaa->ja', V)
        return bool(np.all(V == diag[:, :, None] * np.eye(V.shape[1])[None]))

    def ensemble_eigensystems(self, ensemble_size, member0=0):
        """(E, U) of the members in the rotating frame, computed from the
        lab-frame matrices like Hamiltonian.eig (hamiltonian.py:310-328), with
        one batched LAPACK call instead of one eigh per member."""
        ham, ss = self.hamiltonian, self.hilbert_subspace
        shifts = ham.sampled_site_shifts(ensemble_size, member0)
        if shifts is None:
            return None
        lab = ham._root
        H0 = np.asarray(lab.H(ss), dtype=float)
        number = np.einsum('jaa->ja', ham.system_bath_couplings(ss))
        diag = np.arange(H0.shape[0])
        H = np.broadcast_to(H0, (ensemble_size,) + H0.shape).copy()
        H[:, diag, diag] += shifts @ number
        E, U = np.linalg.eigh(H)
        for letter, quanta in (('e', 1), ('f', 2)):
            if letter in ss:
                E[:, ham.hilbert_subspace_index(letter, ss)] -= quanta * ham.rw_freq
        return E, U

    def ensemble_eom(self, ensemble_size, random_orientations,
                     liouville_subspace, heisenberg_picture=False, member0=0):
        """Device-side construction of all members' generators (K5): the host
        only replays the members' seeded disorder draws; eigensystems (cyclic
        Jacobi), bath correlation matrices, Redfield tensors and the site-basis
        transform are computed on the GPU."""
        ham, ss = self.hamiltonian, self.hilbert_subspace
        # site-basis evolution: the members' disorder shifts are generated on the device
        # as well (nothing but the seed crosses PCIe); the eigenbasis path below needs
        # them on the host for LAPACK's eigenvector gauge
        on_device = self._device_buildable() and self.evolve_basis != 'eigen'
        shifts = ham.sampled_site_shifts_device(ensemble_size, member0) if on_device else None
        bath = ham.bath
        if shifts is None:
            if not self._tensor_device_buildable():
                return super(RedfieldModel, self).ensemble_eom(
                    ensemble_size, random_orientations, liouville_subspace,
                    heisenberg_picture, member0)
            # Eigensystems from the member Hamiltonians themselves (Hamiltonian.eig, scipy's
            # driver: for eigen-basis evolution the eigenvector sign / degenerate-subspace gauge
            # must be the one the members' own dipole operators and states are expressed in;
            # vibronic and complex Hamiltonians have no device eigensolver), Redfield tensors
            # and the basis transform on the device.
            members = [ham.sample(member0 + n) for n in range(ensemble_size)]
            E = np.array([m.E(ss) for m in members])
            U = np.array([m.U(ss) for m in members])
            number = np.einsum('jaa->ja', ham.system_bath_couplings(ss)).real
            kind = (_capi.BATH_DEBYE_REAL if self.discard_imag_corr
                    else _capi.BATH_DEBYE_COMPLEX)
            L = engine.redfield_build(
                E, U, number, kind, bath.temperature, bath.reorg_energy,
                bath.cutoff_freq, self.secular, self.evolve_basis == 'eigen',
                self.unit_convert,
                self.liouville_subspace_index(liouville_subspace),
                transposed=not heisenberg_picture)
            return self.with_transposition(engine.DenseEOM.from_transposed(L),
                                       liouville_subspace, heisenberg_picture)
        number = np.einsum('jaa->ja', ham.system_bath_couplings(ss))
        kind = (_capi.BATH_DEBYE_REAL if self.discard_imag_corr
                else _capi.BATH_DEBYE_COMPLEX)
        lab = ham._root
        quanta = np.zeros(ham.n_states(ss))
        for letter, n_exc in (('e', 1), ('f', 2)):
            if letter in ss:
                quanta[ham.hilbert_subspace_index(letter, ss)] = n_exc
        L = engine.redfield_build_sampled(
            np.asarray(lab.H(ss), dtype=float), shifts, quanta, ham.rw_freq,
            number, kind, bath.temperature, bath.reorg_energy, bath.cutoff_freq,
            self.secular, self.evolve_basis == 'eigen', self.unit_convert,
            self.liouville_subspace_index(liouville_subspace),
            transposed=not heisenberg_picture)
        # the builder wrote the engine's storage layout directly (the Heisenberg
        # picture L^T in that layout is plain row-major L): no copy, no cudaMalloc
        return self.with_transposition(engine.DenseEOM.from_transposed(L),
                                       liouville_subspace, heisenberg_picture)
