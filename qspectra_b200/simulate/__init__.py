from .utils import integrate, fourier_transform, bound_signal, IntegratorError
from .eom import simulate_dynamics, simulate_with_fields, simulate_pump
from .response import (linear_response, absorption_spectra, impulsive_probe,
                       third_order_response, two_dimensional_spectra,
                       PUMP_PROBE_PATHWAYS, THIRD_ORDER_PATHWAYS)
