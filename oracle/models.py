"""
ORACLE (test infrastructure) -- plugin-protocol models on top of the array
generators, driven by any object with the reference's Hamiltonian interface
(``H``, ``E``, ``U``, ``system_bath_couplings``, ``dipole_operator``,
``thermal_state``, ``bath``, ``sample_ensemble``, ``time_step`` ...).

Restates ``qspectra/dynamics/base.py:41-140``, ``liouville_space.py:151-349``,
``redfield.py:107-152``, ``unitary.py``, ``heom.py:22-58, 177-296`` and
``zofe.py:8-119, 204-234``.
"""
import copy

import numpy as np
import scipy.sparse as sp

from . import generators as gen
from .liouville import (subspace_index, full_subspace, ket_vec, embed_state,
                        operator_blocks)

__all__ = ['OracleOperator', 'OracleRedfield', 'OracleUnitary', 'OracleHEOM',
           'OracleZOFE']


class OracleOperator(object):
    """SystemOperator protocol (base.py:143-181) for Liouville-space models;
    ``n_ado`` > 1 applies the blocks to every ADO (heom.py:22-58)."""

    def __init__(self, op, subspace_map, index_of, n_ado=1):
        parts = (subspace_map.split('->') if '->' in subspace_map
                 else [subspace_map, subspace_map])
        f_idx, t_idx = index_of(parts[0]), index_of(parts[1])
        (self._left, self._right, self._comm, bra,
         self._expect) = operator_blocks(op, f_idx, t_idx)
        self.n_ado = n_ado
        self.bra_vector = np.concatenate(
            [bra, np.zeros(len(bra) * (n_ado - 1), dtype=complex)])
        self._mf = len(f_idx)

    def _per_ado(self, block, state):
        ados = np.asarray(state).reshape(self.n_ado, self._mf)
        return (ados @ block.T).reshape(-1)

    def left_multiply(self, state):
        return self._per_ado(self._left, state)

    def right_multiply(self, state):
        return self._per_ado(self._right, state)

    def commutator(self, state):
        return self._per_ado(self._comm, state)

    def expectation_value(self, state):
        return self._expect @ np.asarray(state)[:self._mf]


class _Base(object):
    def __init__(self, hamiltonian, rw_freq=None, hilbert_subspace='gef',
                 unit_convert=1):
        self.hamiltonian = hamiltonian.in_rotating_frame(rw_freq)   # base.py:43
        self.rw_freq = self.hamiltonian.rw_freq
        self.hilbert_subspace = hilbert_subspace
        self.unit_convert = unit_convert
        self.n_ado = 1

    @property
    def time_step(self):                                            # base.py:130-136
        return self.hamiltonian.time_step / self.unit_convert

    def sample_ensemble(self, *args, **kwargs):                     # base.py:120-128
        for ham in self.hamiltonian.sample_ensemble(*args, **kwargs):
            member = copy.copy(self)
            member.hamiltonian = ham
            yield member

    def index(self, subspace):
        h = self.hamiltonian
        return subspace_index(subspace, self.hilbert_subspace, h.n_sites,
                              int(h.n_vibrational_states))

    def dipole_operator(self, subspace_map, polarization, transitions='-+'):
        op = self.hamiltonian.dipole_operator(self.hilbert_subspace,
                                              polarization, transitions)
        return OracleOperator(op, subspace_map, self.index, self.n_ado)

    def _lspace_map(self, state, from_subspace, to_subspace):
        N = self.hamiltonian.n_states(self.hilbert_subspace)
        return embed_state(state, self.index(from_subspace),
                           self.index(to_subspace), N)

    def _pad(self, state):
        return np.concatenate(
            [state, np.zeros(len(state) * (self.n_ado - 1), dtype=complex)])

    def map_between_subspaces(self, state, from_subspace, to_subspace):
        ados = np.asarray(state).reshape(self.n_ado, -1)
        return np.concatenate([self._lspace_map(a, from_subspace, to_subspace)
                               for a in ados])

    def thermal_state(self, subspace):                   # liouville_space.py:300-309
        rho0 = self.hamiltonian.thermal_state(subspace)
        return self._pad(self._lspace_map(ket_vec(rho0),
                                          full_subspace(subspace), subspace))

    def density_matrix_to_state_vector(self, rho0, subspace):
        return self._pad(self._lspace_map(ket_vec(rho0),
                                          full_subspace(subspace), subspace))

    def state_vector_to_density_matrix(self, states):
        states = np.asarray(states)
        M = states.shape[-1] // self.n_ado
        states = states[..., :M]
        N = int(np.sqrt(M))
        return states.reshape(-1, N, N, order='F')


class _Linear(_Base):
    def generator(self, subspace, heisenberg_picture=False):
        idx = self.index(subspace)
        L = self.full_generator()[np.ix_(idx, idx)]
        return L.T if heisenberg_picture else L      # liouville_space.py:325-330

    def equation_of_motion(self, subspace, heisenberg_picture=False):
        L = self.generator(subspace, heisenberg_picture)
        return lambda t, y: L.dot(y)


class OracleRedfield(_Linear):
    def __init__(self, hamiltonian, rw_freq=None, hilbert_subspace='gef',
                 unit_convert=1, secular=True, discard_imag_corr=False,
                 evolve_basis='site'):
        super(OracleRedfield, self).__init__(hamiltonian, rw_freq,
                                             hilbert_subspace, unit_convert)
        self.secular = secular
        self.discard_imag_corr = discard_imag_corr
        self.evolve_basis = evolve_basis

    def full_generator(self):                                   # redfield.py:146-152
        h, ss = self.hamiltonian, self.hilbert_subspace
        corr = (h.bath.corr_func_real if self.discard_imag_corr
                else h.bath.corr_func_complex)
        return self.unit_convert * gen.redfield_generator(
            h.E(ss), h.U(ss), h.system_bath_couplings(ss), corr,
            self.secular, self.evolve_basis)


class OracleUnitary(_Linear):
    def full_generator(self):                                   # unitary.py:5-9
        return gen.unitary_generator(
            self.unit_convert * self.hamiltonian.H(self.hilbert_subspace))


class OracleHEOM(_Linear):
    def __init__(self, hamiltonian, rw_freq=None, hilbert_subspace='gef',
                 unit_convert=1, level_cutoff=3, K=1, low_temp_corr=True,
                 modified_HEOM=False):
        super(OracleHEOM, self).__init__(hamiltonian, rw_freq,
                                         hilbert_subspace, unit_convert)
        self.level_cutoff, self.K = level_cutoff, K
        self.low_temp_corr, self.modified_HEOM = low_temp_corr, modified_HEOM
        self.table = gen.ado_table(self.hamiltonian.n_sites, K, level_cutoff)
        self.n_ado = len(self.table)
        # reference quirk 3: the thermal state comes from a model built on the
        # hamiltonian handed to the constructor, not the sampled one
        self._thermal_ham = self.hamiltonian

    def thermal_state(self, subspace):
        rho0 = self._thermal_ham.thermal_state(subspace)
        return self._pad(self._lspace_map(ket_vec(rho0),
                                          full_subspace(subspace), subspace))

    def generator(self, subspace, heisenberg_picture=False):     # heom.py:228-244
        h, ss, b = self.hamiltonian, self.hilbert_subspace, self.hamiltonian.bath
        L = self.unit_convert * gen.heom_generator(
            h.H(ss), h.system_bath_couplings(ss), self.index(subspace),
            b.cutoff_freq, b.temperature, b.reorg_energy, self.K,
            self.level_cutoff, self.low_temp_corr, self.modified_HEOM)
        return sp.csr_matrix(L.T) if heisenberg_picture else L


class OracleZOFE(_Base):
    def __init__(self, hamiltonian, rw_freq=None, hilbert_subspace='gef',
                 unit_convert=1, ham_hermit=False, rho_hermit=False):
        super(OracleZOFE, self).__init__(hamiltonian, rw_freq,
                                         hilbert_subspace, unit_convert)
        self.ham_hermit, self.rho_hermit = ham_hermit, rho_hermit
        b = self.hamiltonian.bath
        n = self.hamiltonian.n_states(hilbert_subspace)
        self.oop_shape = (b.numb_pm, self.hamiltonian.n_sites, n, n)

    def density_matrix_to_state_vector(self, rho0, subspace):    # zofe.py:89-93
        return np.append(ket_vec(rho0),
                         np.zeros(int(np.prod(self.oop_shape)), dtype=complex))

    def thermal_state(self, _):
        return self.density_matrix_to_state_vector(
            self.hamiltonian.thermal_state(self.hilbert_subspace), None)

    def map_between_subspaces(self, state, from_subspace, to_subspace):
        return state                                              # zofe.py:106-107

    def state_vector_to_density_matrix(self, states):
        n = self.oop_shape[-1]
        return np.array([s[:n * n].reshape((n, n), order='F') for s in states])

    def dipole_operator(self, subspace_map, polarization, transitions='-+'):
        raise NotImplementedError('oracle covers ZOFE free evolution only')

    def equation_of_motion(self, subspace, heisenberg_picture=False):
        if heisenberg_picture:                                    # zofe.py:211-213
            raise NotImplementedError('ZOFE not implemented in the Heisenberg '
                                      'picture')
        h, b = self.hamiltonian, self.hamiltonian.bath
        Ln = -np.asanyarray(h.system_bath_couplings(self.hilbert_subspace))
        Gamma = b.Omega ** 2 * b.huang
        w = 1j * b.Omega + b.gamma
        H = h.H(self.hilbert_subspace)
        u = self.unit_convert
        return lambda t, y: u * gen.zofe_rhs(y, H, Ln, Gamma, w,
                                             self.ham_hermit, self.rho_hermit)
