"""
Pulse envelopes in the rotating frame (contract: reference
``qspectra/pulse.py:8-114``).  ``GaussianPulse`` additionally exposes
``device_params`` so the fused integrator can evaluate the field on the GPU
without a host callback.
"""
from abc import ABCMeta, abstractmethod

import numpy as np

from .constants import GAUSSIAN_SD_FWHM
from .utils import simple_repr


class Pulse(metaclass=ABCMeta):
    """Callable ``E(t, rw_freq)`` with ``t_init`` / ``t_final`` attributes."""

    @abstractmethod
    def __call__(self, t, rw_freq):
        """complex field at time(s) ``t`` in the frame rotating at ``rw_freq``"""


class CustomPulse(Pulse):
    def __init__(self, t_init, t_final, call):
        self.t_init = t_init
        self.t_final = t_final
        self.call = call

    def __call__(self, t, rw_freq):
        return self.call(t, rw_freq)

    def __repr__(self):
        return simple_repr(self, ['t_init', 't_final', 'call'])


class GaussianPulse(Pulse):
    def __init__(self, carrier_freq, fwhm, t_peak=0, scale=1, freq_convert=1,
                 t_limits_multiple=3):
        sigma = GAUSSIAN_SD_FWHM * fwhm
        self.fwhm = fwhm
        self.two_sigma_squared = 2 * sigma ** 2
        self.t_init = t_peak - t_limits_multiple * sigma
        self.t_final = t_peak + t_limits_multiple * sigma
        self.t_peak = t_peak
        self.carrier_freq = carrier_freq
        self.scale = scale
        self.freq_convert = freq_convert
        self.t_limits_multiple = t_limits_multiple

    def __call__(self, t, rw_freq):
        dt = t - self.t_peak
        phase = 1j * self.freq_convert * (self.carrier_freq - rw_freq) * dt
        return self.scale * np.exp(phase - dt ** 2 / self.two_sigma_squared)

    def device_params(self, rw_freq):
        """(scale, detuning [rad / time], t_peak, 1 / two_sigma_squared)."""
        return (float(self.scale),
                float(self.freq_convert * (self.carrier_freq - rw_freq)),
                float(self.t_peak), float(1.0 / self.two_sigma_squared))

    def __repr__(self):
        return simple_repr(self, ['carrier_freq', 'fwhm', 't_peak', 'scale',
                                  'freq_convert', 't_limits_multiple'])
