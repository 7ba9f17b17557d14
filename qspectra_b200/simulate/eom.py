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
from .. import _capi, engine
from ..engine import LinearMap


def simulate_dynamics(dynamical_model, initial_state, duration=None, times=None,
                      liouville_subspace='ee', save_func=None,
                      ensemble_size=None, ensemble_random_orientations=False,
                      **integrate_kwargs):
    """Free evolution; returns (t, density matrices) -- ensemble-averaged when
    ``ensemble_size`` is given.  Same arguments as the reference.  The whole
    ensemble is one batched device propagation followed by a device-side mean
    (kernels K5 -> K1/K4 -> K6); ``member_offset`` (extension) shifts the
    member numbers for multi-GPU sharding."""
    member0 = integrate_kwargs.pop('member_offset', 0)
    initial_state = np.asarray(initial_state)
    if initial_state.ndim == 1:
        # wavefunction -> density matrix; np.outer(psi*, psi) as in the reference
        # (eom.py:17-19; that is the transpose of |psi><psi| for complex psi)
        initial_state = np.outer(initial_state.conj(), initial_state)
    t = (np.arange(0, duration, dynamical_model.time_step)
         if times is None else np.asarray(times, dtype=float))
    y0 = dynamical_model.density_matrix_to_state_vector(initial_state,
                                                        liouville_subspace)
    save = save_func if save_func is not None else dynamical_model.dynamics_save
    if ensemble_size is None:
        eom = dynamical_model.equation_of_motion(liouville_subspace)
        states = integrate(eom, y0, t, save_func=save, **integrate_kwargs)
    else:
        eom = dynamical_model.ensemble_eom(
            ensemble_size, ensemble_random_orientations, liouville_subspace,
            member0=member0)
        states = ensemble_mean(eom, y0, t, ensemble_size, save,
                               **integrate_kwargs)
    if save_func is None:
        states = dynamical_model.saved_states_to_density_matrix(states)
    return (t, states)


def ensemble_mean(eom, y0, t, ensemble_size, save, return_device=False,
                  scale=None, **integrate_kwargs):
    """mean over members of save(y_m(t)): batched propagation + device sum."""
    opts = {k: integrate_kwargs[k] for k in ('rtol', 'atol', 'rk4_substeps')
            if k in integrate_kwargs}
    if save is not None and not isinstance(save, (LinearMap, tuple)):
        # arbitrary host callable: save full states, post-process on the host
        states = integrate(eom, np.broadcast_to(y0, (ensemble_size,) + y0.shape),
                           t, save_func=save, generators=np.arange(ensemble_size),
                           **integrate_kwargs)
        return states.sum(axis=0) / ensemble_size
    torch = _capi.torch_cuda()
    y0_dev = _capi.to_device(y0).reshape(1, -1).expand(ensemble_size, -1).contiguous()
    # Hermitian initial state on a subspace closed under transposition: dense generators step it
    # in real coordinates, and the member sum runs over the real rows (engine.HermitianTrajectory)
    hermitian = isinstance(eom, engine.DenseEOM) and eom._hermitian_state(np.asarray(y0))
    if hermitian:
        opts.update(hermitian_state=True, packed=True)
    out = eom.propagate(y0_dev, t, method=integrate_kwargs.get('method_name', 'zvode'),
                        save=save, generators=np.arange(ensemble_size),
                        return_device=True, **opts)
    mean = engine.reduce_members(out, (1.0 / ensemble_size) if scale is None else scale)
    if return_device:
        return mean
    res = _capi.to_host(mean)
    if isinstance(out, engine.HermitianTrajectory) and not out.source.ok():
        # the device-side check found a generator that does not commute with Hermitian
        # conjugation: repeat on the complex path
        out.source.__dict__['_fallback'] = True
        eom.hermitian_perm = None
        return ensemble_mean(eom, y0, t, ensemble_size, save, return_device, scale,
                             **integrate_kwargs)
    if isinstance(save, LinearMap) and save.matrix.ndim == 1:
        res = res[..., 0]
    return res


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
    comm = V[0].commutator
    if not hasattr(comm, 'matrix'):
        # ZOFE: the Hilbert-space dipole operator multiplies rho and every
        # auxiliary operator from the left / right (zofe.py:24-35)
        return descr, np.array([Vi.operator for Vi in V], dtype=complex)
    # dense models: (M, M) commutator blocks; HEOM: the same block acts on every
    # ADO (heom.py:22-58), the device applies it per ADO
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
