"""
Bath models (contract: reference ``qspectra/bath.py:17-144``).

``DebyeBath.corr_func_complex`` is the one-sided correlation spectrum with a
1000-term Matsubara sum; the batched device version used for disorder
ensembles lives in ``csrc/redfield_build.cu`` and is tested against this one.
"""
import numpy as np

from .utils import simple_repr


class Bath(object):
    def corr_func_real(self, x):
        """(n(x) + 1) * J_antisymmetrised(x); T * J'(0) at x == 0."""
        T = self.temperature
        if x == 0:
            return T * self.spectral_density_limit_at_zero
        J = self.spectral_density_func
        J_anti = J(x) if x >= 0 else -J(-x)
        return (1 / np.expm1(x / T) + 1) * J_anti

    def spectral_density_func(self, x):
        raise NotImplementedError

    @property
    def spectral_density_limit_at_zero(self):
        raise NotImplementedError


class ArbitraryBath(Bath):
    def __init__(self, temperature, spectral_density_func,
                 spectral_density_limit_at_zero):
        self.temperature = temperature
        self.spectral_density_func = spectral_density_func
        self.spectral_density_limit_at_zero = spectral_density_limit_at_zero


class UncoupledBath(Bath):
    def corr_func_complex(self, _):
        return 0 + 0j

    def spectral_density_func(self, _):
        return 0

    @property
    def spectral_density_limit_at_zero(self):
        return 0


class DebyeBath(Bath):
    """Drude-Lorentz bath J(x) = 2 lambda gamma x / (gamma^2 + x^2)."""

    def __init__(self, temperature, reorg_energy, cutoff_freq):
        self.temperature = float(temperature)
        self.reorg_energy = float(reorg_energy)
        self.cutoff_freq = float(cutoff_freq)

    def __repr__(self):
        return simple_repr(self, ['temperature', 'reorg_energy', 'cutoff_freq'])

    def __eq__(self, other):
        return (type(other) is type(self)
                and self.temperature == other.temperature
                and self.reorg_energy == other.reorg_energy
                and self.cutoff_freq == other.cutoff_freq)

    __hash__ = None

    def spectral_density_func(self, x):
        g = self.cutoff_freq
        return 2 * self.reorg_energy * g * x / (g ** 2 + x ** 2)

    @property
    def spectral_density_limit_at_zero(self):
        return 2 * self.reorg_energy / self.cutoff_freq

    def corr_func_complex(self, x, matsubara_cutoff=1000):
        T, lam, g = self.temperature, self.reorg_energy, self.cutoff_freq
        if x == 0:
            return lam * (2 * T / g - 1j)
        nu = 2 * np.pi * np.arange(matsubara_cutoff) * T
        matsubara = np.sum(nu / ((nu ** 2 - g ** 2) * (nu - 1j * x)))
        drude = (1 / np.tan(g / (2 * T)) - 1j) / (g - 1j * x)
        return lam * g * (drude + 4 * T * matsubara)


class PseudomodeBath(Bath):
    """Lorentzian pseudomode decomposition of the bath correlation spectrum
    (arrays of shape (numb_pm, n_sites))."""

    def __init__(self, numb_pm, Omega, gamma, huang):
        self.numb_pm = numb_pm
        self.Omega = Omega
        self.gamma = gamma
        self.huang = huang
