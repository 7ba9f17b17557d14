"""
ZOFE master equation model (pseudomode baths).

Contract: reference ``qspectra/dynamics/zofe.py`` -- ``ZOFESpaceOperator``
:8-41, ``ZOFEModel`` :44-119, the right-hand side :121-202 and
``equation_of_motion`` :204-234 (no Heisenberg picture: raises
``NotImplementedError``, which the simulate layer uses as control flow).
The right-hand side and its integration run on the GPU (csrc/zofe.cu).
"""
import numpy as np

from .base import DynamicalModel, SystemOperator
from ..bath import PseudomodeBath
from ..engine import ZofeEOM, LinearMap
from ..utils import imemoize


class _ZofeMap(object):
    """left / right multiplication of rho and of every auxiliary operator."""

    def __init__(self, operator, model, side):
        self.operator, self.model, self.side = np.asarray(operator), model, side

    def __call__(self, state):
        rho, oop = self.model.state_vec_to_operators(np.asarray(state))
        A = self.operator
        if self.side == 'left':
            return self.model.operators_to_state_vec(
                A @ rho, np.einsum('cd,psde->psce', A, oop))
        return self.model.operators_to_state_vec(rho @ A, oop @ A)


class ZOFESpaceOperator(SystemOperator):
    def __init__(self, operator, liouv_subspace_map, dynamical_model):
        self.operator = np.asarray(operator)
        self.dynamical_model = dynamical_model

    @property
    def left_multiply(self):
        return _ZofeMap(self.operator, self.dynamical_model, 'left')

    @property
    def right_multiply(self):
        return _ZofeMap(self.operator, self.dynamical_model, 'right')

    def commutator(self, state):
        return self.left_multiply(state) - self.right_multiply(state)

    @property
    def expectation_value(self):
        # tr(M rho) = sum_ij M_ij rho_ji on the column-major vec(rho):
        # rho_ji sits at j + n i  ->  row vector M^T flattened row-major ... = M.T.reshape(order='F')
        n = len(self.operator)
        row = self.operator.T.reshape(-1, order='F')
        assert row.size == n * n
        return LinearMap(row, ado0_only=True)


class ZOFEModel(DynamicalModel):
    system_operator = ZOFESpaceOperator

    def __init__(self, hamiltonian, rw_freq=None, hilbert_subspace='gef',
                 unit_convert=1, ham_hermit=False, rho_hermit=False):
        super(ZOFEModel, self).__init__(hamiltonian, rw_freq, hilbert_subspace,
                                        unit_convert)
        if not isinstance(self.hamiltonian.bath, PseudomodeBath):
            raise NotImplementedError('ZOFE only implemented for baths of type '
                                      'PseudomodeBath')
        n = self.hamiltonian.n_states(self.hilbert_subspace)
        self.oop_shape = (self.hamiltonian.bath.numb_pm,
                          self.hamiltonian.n_sites, n, n)
        self.ham_hermit = ham_hermit
        self.rho_hermit = rho_hermit

    # -- states ---------------------------------------------------------------
    def density_matrix_to_state_vector(self, rho0, liouville_subspace):
        return np.append(np.asarray(rho0, dtype=complex).reshape(-1, order='F'),
                         np.zeros(int(np.prod(self.oop_shape)), dtype=complex))

    def state_vector_to_density_matrix(self, rhos):
        n = self.oop_shape[-1]
        rhos = np.asarray(rhos)
        return np.array([r[:n * n].reshape((n, n), order='F') for r in rhos])

    #: simulate_dynamics only needs rho: let the device save just that
    dynamics_save = ('ado0',)

    def thermal_state(self, _):
        rho0 = self.hamiltonian.thermal_state(self.hilbert_subspace)
        return self.density_matrix_to_state_vector(rho0, None)

    def map_between_subspaces(self, state, from_subspace, to_subspace):
        return state

    def state_vec_to_operators(self, rho_oop_vec):
        n = self.oop_shape[-1]
        rho = rho_oop_vec[:n * n].reshape((n, n), order='F')
        oop = rho_oop_vec[n * n:].reshape(self.oop_shape, order='F')
        return rho, oop

    def operators_to_state_vec(self, rho, oop):
        return np.append(rho.reshape(-1, order='F'), oop.reshape(-1, order='F'))

    # -- dynamics -------------------------------------------------------------
    def _device_eom(self, hamiltonians):
        ham, ss, bath = self.hamiltonian, self.hilbert_subspace, self.hamiltonian.bath
        V = np.asarray(ham.system_bath_couplings(ss))
        diag = np.einsum('jaa->ja', V)
        if np.abs(V - np.einsum('ja,ab->jab', diag, np.eye(V.shape[-1]))).max() > 0 \
                or np.abs(np.imag(diag)).max() > 0:
            raise NotImplementedError('ZOFE kernel needs real diagonal '
                                      'system-bath coupling operators')
        Omega = np.asarray(bath.Omega, dtype=complex)
        gamma = np.asarray(bath.gamma, dtype=complex)
        huang = np.asarray(bath.huang, dtype=complex)
        H = np.array([h.H(ss) for h in hamiltonians], dtype=complex)
        return ZofeEOM(H, np.real(diag), Omega ** 2 * huang, 1j * Omega + gamma,
                       self.unit_convert, self.ham_hermit, self.rho_hermit)

    @imemoize
    def equation_of_motion(self, liouville_subspace, heisenberg_picture=False):
        if heisenberg_picture:
            raise NotImplementedError('ZOFE not implemented in the Heisenberg '
                                      'picture')
        return self._device_eom([self.hamiltonian])

    def ensemble_equation_of_motion(self, members, liouville_subspace,
                                    heisenberg_picture=False):
        if heisenberg_picture:
            raise NotImplementedError('ZOFE not implemented in the Heisenberg '
                                      'picture')
        if len(members) == 1:
            return members[0].equation_of_motion(liouville_subspace)
        return self._device_eom([m.hamiltonian for m in members])
