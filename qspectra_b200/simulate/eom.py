"""
Equation-of-motion based simulation entry points.

Contract: reference ``qspectra/simulate/eom.py`` -- ``simulate_dynamics``
:11-74, ``simulate_with_fields`` :79-164, ``simulate_pump`` :167-227.  The
disorder-ensemble loop of ``optional_ensemble_average`` is executed as ONE
batched device propagation (one generator per member) followed by a mean over
members, instead of a serial Python loop.
"""
import numpy as np

from .decorators import (optional_ensemble_average,
                         optional_2nd_order_isotropic_average)
from .utils import integrate, DrivenEOM
from .. import _capi


def ensemble_members(dynamical_model, ensemble_size, random_orientations):
    """[model] or the list of sampled member models (reference
    decorators.py:55-57 order: member n = sample n)."""
    if ensemble_size is None:
        return [dynamical_model], False
    return list(dynamical_model.sample_ensemble(ensemble_size,
                                                random_orientations)), True


def simulate_dynamics(dynamical_model, initial_state, duration=None, times=None,
                      liouville_subspace='ee', save_func=None,
                      ensemble_size=None, ensemble_random_orientations=False,
                      **integrate_kwargs):
    """Free evolution; returns (t, density matrices) -- ensemble-averaged when
    ``ensemble_size`` is given.  Same arguments as the reference."""
    members, averaged = ensemble_members(dynamical_model, ensemble_size,
                                         ensemble_random_orientations)
    initial_state = np.asarray(initial_state)
    if initial_state.ndim == 1:
        # wavefunction -> density matrix; np.outer(psi*, psi) as in the reference
        # (eom.py:17-19; that is the transpose of |psi><psi| for complex psi)
        initial_state = np.outer(initial_state.conj(), initial_state)
    t = (np.arange(0, duration, dynamical_model.time_step)
         if times is None else np.asarray(times, dtype=float))
    eom = dynamical_model.ensemble_equation_of_motion(members,
                                                      liouville_subspace)
    y0 = members[0].density_matrix_to_state_vector(initial_state,
                                                   liouville_subspace)
    save = save_func if save_func is not None else dynamical_model.dynamics_save
    if not averaged:
        states = integrate(eom, y0, t, save_func=save, **integrate_kwargs)
    else:
        E = len(members)
        states = integrate(eom, np.broadcast_to(y0, (E,) + y0.shape), t,
                           save_func=save, generators=np.arange(E),
                           **integrate_kwargs)
        states = states.sum(axis=0) / E
    if save_func is None:
        states = dynamical_model.saved_states_to_density_matrix(states)
    return (t, states)


def _field_terms(dynamical_model, pulses, geometry, polarization,
                 liouville_subspace):
    """Pulse descriptors + commutator matrices of the dipole operators."""
    V = [dynamical_model.dipole_operator(liouville_subspace, polar, trans)
         for polar, trans in zip(polarization, geometry)]
    descr = []
    for pulse, trans in zip(pulses, geometry):
        if not hasattr(pulse, 'device_params'):
            raise NotImplementedError(
                'only GaussianPulse fields can be evaluated inside the fused '
                'GPU integrator (got %r)' % (pulse,))
        descr.append(pulse.device_params(dynamical_model.rw_freq)
                     + (trans == '+',))
    n_ado = getattr(V[0].commutator, 'n_ado', 1)
    if n_ado != 1:
        raise NotImplementedError('pulse-driven HEOM propagation is not '
                                  'available yet')
    ops = np.array([Vi.commutator.matrix for Vi in V], dtype=complex)
    return descr, ops


def _simulate_with_fields(dynamical_model, pulses, geometry, polarization,
                          time_extra, times, liouville_subspace, save_func,
                          **integrate_kwargs):
    eom = dynamical_model.equation_of_motion(liouville_subspace)
    descr, ops = _field_terms(dynamical_model, pulses, geometry, polarization,
                              liouville_subspace)
    f = DrivenEOM(eom, descr, ops[None])
    initial_state = dynamical_model.thermal_state(liouville_subspace)
    t0 = min(p.t_init for p in pulses)
    tf = max(p.t_final for p in pulses)
    t = (np.arange(t0, tf + time_extra, dynamical_model.time_step)
         if times is None else tf + np.asarray(times, dtype=float))
    states = integrate(f, initial_state, t, t0=t0, save_func=save_func,
                       **integrate_kwargs)
    return (t, states)


def simulate_with_fields(dynamical_model, pulses, geometry='-+',
                         polarization='xx', time_extra=0, times=None,
                         liouville_subspace='gg,ge,eg,ee', save_func=None,
                         ensemble_size=None, ensemble_random_orientations=False,
                         **integrate_kwargs):
    """Evolution under pulses in the rotating-wave approximation."""
    return optional_ensemble_average(_simulate_with_fields)(
        dynamical_model, pulses, geometry, polarization, time_extra, times,
        liouville_subspace, save_func, ensemble_size=ensemble_size,
        ensemble_random_orientations=ensemble_random_orientations,
        **integrate_kwargs)


def simulate_pump(dynamical_model, pump, polarization='x', time_extra=0,
                  times=None, liouville_subspace='gg,ge,eg,ee', save_func=None,
                  ensemble_size=None, ensemble_random_orientations=False,
                  exact_isotropic_average=False, **integrate_kwargs):
    """Evolution under a pump field (second order: the pump acts as '-' and
    '+')."""
    return optional_ensemble_average(
        optional_2nd_order_isotropic_average(_simulate_with_fields))(
            dynamical_model, [pump, pump], '-+', [polarization, polarization],
            time_extra, times, liouville_subspace, save_func,
            ensemble_size=ensemble_size,
            ensemble_random_orientations=ensemble_random_orientations,
            exact_isotropic_average=exact_isotropic_average,
            **integrate_kwargs)
