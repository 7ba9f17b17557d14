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
from .dynamics.liouville_space import (matrix_to_ket_vec, ket_vec_to_matrix,
                                       matrix_to_bra_vec)
from .dynamics.redfield import RedfieldModel
from .dynamics.unitary import UnitaryModel
from .dynamics.heom import HEOMModel
from .dynamics.zofe import ZOFEModel
from .simulate.eom import (simulate_dynamics, simulate_with_fields,
                           simulate_pump)
from .simulate.response import (linear_response, absorption_spectra,
                                impulsive_probe, third_order_response,
                                two_dimensional_spectra, PUMP_PROBE_PATHWAYS,
                                THIRD_ORDER_PATHWAYS)
from .simulate.utils import (fourier_transform, integrate, bound_signal,
                             IntegratorError)
