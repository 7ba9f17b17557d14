"""
Plugin contract of the propagation engine: the ``DynamicalModel`` and
``SystemOperator`` protocols of the reference (dynamics/base.py:7-181), kept
name-for-name so GPU models are drop-in replacements.

What differs: ``equation_of_motion`` returns a ``DeviceEOM`` (callable like the
reference's closure, but resident on the GPU), and models expose
``ensemble_equation_of_motion`` so the simulate layer can propagate a whole
disorder ensemble in one kernel instead of looping over ``sample_ensemble``.
"""
from abc import ABCMeta, abstractmethod

from ..utils import copy_with_new_cache


class DynamicalModel(metaclass=ABCMeta):
    """See reference dynamics/base.py:7-46 for the parameter semantics."""

    def __init__(self, hamiltonian, rw_freq=None, hilbert_subspace='gef',
                 unit_convert=1):
        self.hamiltonian = hamiltonian.in_rotating_frame(rw_freq)
        self.rw_freq = self.hamiltonian.rw_freq
        self.hilbert_subspace = hilbert_subspace
        self.unit_convert = unit_convert

    def __repr__(self):
        return '%s(hamiltonian=%r, rw_freq=%r, hilbert_subspace=%r, ' \
               'unit_convert=%r)' % (type(self).__name__, self.hamiltonian,
                                     self.rw_freq, self.hilbert_subspace,
                                     self.unit_convert)

    @abstractmethod
    def thermal_state(self, liouville_subspace):
        """thermal state as a state vector on the subspace"""

    @abstractmethod
    def equation_of_motion(self, liouville_subspace, heisenberg_picture=False):
        """-> callable f(t, y) (a DeviceEOM).  Raises NotImplementedError when
        the model has no Heisenberg-picture form (control flow in the simulate
        layer, reference response.py:34, 328)."""

    @abstractmethod
    def map_between_subspaces(self, state, from_subspace, to_subspace):
        """re-express a state vector on another Liouville subspace"""

    @abstractmethod
    def density_matrix_to_state_vector(self, rho0, liouville_subspace):
        """density matrix -> initial state vector"""

    @abstractmethod
    def state_vector_to_density_matrix(self, states):
        """trajectory of state vectors -> density matrices"""

    def dipole_operator(self, liouv_subspace_map, polarization,
                        transitions='-+'):
        operator = self.hamiltonian.dipole_operator(self.hilbert_subspace,
                                                    polarization, transitions)
        return self.system_operator(operator, liouv_subspace_map, self)

    def dipole_destroy(self, liouville_subspace_map, polarization):
        return self.dipole_operator(liouville_subspace_map, polarization, '-')

    def dipole_create(self, liouville_subspace_map, polarization):
        return self.dipole_operator(liouville_subspace_map, polarization, '+')

    def sample_ensemble(self, *args, **kwargs):
        """Yields re-sampled shallow copies (fresh memo cache), reference
        base.py:120-128."""
        for ham in self.hamiltonian.sample_ensemble(*args, **kwargs):
            member = copy_with_new_cache(self)
            member.hamiltonian = ham
            yield member

    @property
    def time_step(self):
        return self.hamiltonian.time_step / self.unit_convert

    def hilbert_subspace_index(self, subspace):
        return self.hamiltonian.hilbert_subspace_index(subspace,
                                                       self.hilbert_subspace)

    # -- device extensions (not in the reference) ----------------------------
    #: save spec used by simulate_dynamics when the caller gives no save_func
    dynamics_save = None

    def saved_states_to_density_matrix(self, states):
        return self.state_vector_to_density_matrix(states)

    def ensemble_equation_of_motion(self, members, liouville_subspace,
                                    heisenberg_picture=False):
        """One DeviceEOM holding the generators of all ``members`` (models
        produced by ``sample_ensemble``).  Default: only single-member lists."""
        if len(members) == 1:
            return members[0].equation_of_motion(liouville_subspace,
                                                 heisenberg_picture)
        raise NotImplementedError('%s does not batch ensembles'
                                  % type(self).__name__)

    def ensemble_eom(self, ensemble_size, random_orientations,
                     liouville_subspace, heisenberg_picture=False, member0=0):
        """DeviceEOM whose generator g belongs to ensemble member member0 + g.
        Generic form: materialise the member models; subclasses override this
        with device-side construction."""
        members = [self.sample(member0 + n, random_orientations)
                   for n in range(ensemble_size)]
        return self.ensemble_equation_of_motion(members, liouville_subspace,
                                                heisenberg_picture)

    def sample(self, n, random_orientations=False):
        """n-th ensemble member of this model (reference base.py:120-128)."""
        member = copy_with_new_cache(self)
        member.hamiltonian = self.hamiltonian.sample(n, random_orientations)
        return member


class SystemOperator(metaclass=ABCMeta):
    """Reference dynamics/base.py:143-181.  Implementations return LinearMap
    objects so that ``integrate`` can fuse them as save_func epilogues."""

    def commutator(self, state):
        return self.left_multiply(state) - self.right_multiply(state)

    @abstractmethod
    def left_multiply(self, state):
        """operator . rho"""

    @abstractmethod
    def right_multiply(self, state):
        """rho . operator"""

    @abstractmethod
    def expectation_value(self, state):
        """tr(operator . rho)"""
