"""
qspectra_b200 -- B200-native propagation engine behind the qspectra
``DynamicalModel`` plugin API.  Public names mirror the reference facade
(``qspectra/__init__.py:1-39``).
"""
from .bath import DebyeBath, ArbitraryBath, UncoupledBath, PseudomodeBath
from .constants import CM_FS, CM_K, GAUSSIAN_SD_FWHM
from .hamiltonian import (Hamiltonian, ElectronicHamiltonian,
                          VibronicHamiltonian)
from .operator_tools import (unit_vec, basis_transform_operator,
                             basis_transform_vector, all_states,
                             n_excitations)
from .polarization import (polarization_vector, check_polarizations,
                           invariant_weights_4th_order,
                           invariant_polarizations, FOURTH_ORDER_INVARIANTS,
                           MAGIC_ANGLE)
from .pulse import CustomPulse, GaussianPulse
