"""
ORACLE (test infrastructure) -- the simulate drivers.

Restates ``qspectra/simulate/eom.py:11-105`` (free and field-driven
evolution), ``simulate/response.py:13-43, 103-154, 252-337`` (linear /
third-order response) and the three averaging loops of
``simulate/decorators.py:40-125`` as plain functions.
"""
import numpy as np

from .propagate import integrate

__all__ = ['THIRD_ORDER_PATHWAYS', 'ensemble_average', 'simulate_dynamics',
           'simulate_with_fields', 'linear_response', 'absorption_spectra',
           'third_order_response', 'fourier_transform', 'iso2', 'iso4']

THIRD_ORDER_PATHWAYS = {                                    # response.py:252-264
    '-++': {'ESE': 'gg->ge->ee->eg->gg', 'GSB': 'gg->ge->gg->eg->gg',
            'ESA': 'gg->ge->ee->fe->ee'},
    '+-+': {'ESE': 'gg->eg->ee->eg->gg', 'GSB': 'gg->eg->gg->eg->gg',
            'ESA': 'gg->eg->ee->fe->ee'},
    '++-': {'ESA1': 'gg->eg->fg->fe->ee', 'ESA2': 'gg->eg->fg->eg->gg'},
}


_AXES = {'x': np.array([1., 0, 0]), 'y': np.array([0, 1., 0]),
         'z': np.array([0, 0, 1.])}
FOURTH_ORDER_INVARIANTS = [((0, 1), (2, 3)), ((0, 2), (1, 3)),
                           ((0, 3), (1, 2))]


def _pol(p):
    """polarization.py:14-36"""
    if isinstance(p, str):
        return _AXES[p]
    if np.isscalar(p):
        return np.array([np.cos(p), np.sin(p), 0])
    return np.asarray(p, float).reshape(3)


def invariant_weights_4th_order(polarizations):
    """polarization.py:55-70"""
    e = np.array([_pol(p) for p in polarizations])
    cos = e @ e.T
    prods = np.array([cos[a] * cos[b] for a, b in FOURTH_ORDER_INVARIANTS])
    return (5 * np.eye(3) - np.ones((3, 3))) @ prods / 30


def invariant_polarizations(invariant):
    """polarization.py:73-84"""
    import itertools
    return [''.join(ax) for ax in itertools.product('xyz', repeat=4)
            if all(ax[a] == ax[b] for a, b in invariant)]


def ensemble_average(func, model, ensemble_size, random_orientations=False):
    """decorators.py:40-64: serial mean over sample_ensemble members."""
    if ensemble_size is None:
        return func(model)
    total = None
    for member in model.sample_ensemble(ensemble_size, random_orientations):
        ticks, signal = func(member)
        total = signal if total is None else total + signal
    return ticks, total / ensemble_size


def iso2(func, polarization):
    """decorators.py:99-125: (p0.p1)/3 * (xx + yy + zz)."""
    p = np.array([_pol(x) for x in polarization])
    total = None
    for axes in ('xx', 'yy', 'zz'):
        ticks, signal = func(axes)
        total = signal if total is None else total + signal
    return ticks, total * (np.dot(*p) / 3.0)


def iso4(func, polarization):
    """decorators.py:67-96: weighted sum over the three 4th-order invariants."""
    weights = invariant_weights_4th_order(polarization)
    cache, total, ticks = {}, None, None
    for invariant, weight in zip(FOURTH_ORDER_INVARIANTS, weights):
        if weight > 1e-8:
            for p in invariant_polarizations(invariant):
                if p not in cache:
                    ticks, cache[p] = func(p)
                total = (weight * cache[p] if total is None
                         else total + weight * cache[p])
    return ticks, total


def simulate_dynamics(model, initial_state, duration=None, times=None,
                      liouville_subspace='ee', **kw):
    """eom.py:11-26 (note np.outer(psi.conj(), psi): reference quirk 10)."""
    eom = model.equation_of_motion(liouville_subspace)
    initial_state = np.asarray(initial_state)
    if initial_state.ndim == 1:
        initial_state = np.outer(initial_state.conj(), initial_state)
    y0 = model.density_matrix_to_state_vector(initial_state, liouville_subspace)
    t = np.arange(0, duration, model.time_step) if times is None else times
    states = integrate(eom, y0, t, **kw)
    return t, model.state_vector_to_density_matrix(states)


def simulate_with_fields(model, pulses, geometry='-+', polarization='xx',
                         time_extra=0, times=None,
                         liouville_subspace='gg,ge,eg,ee', **kw):
    """eom.py:79-105: d/dt y = L y + sum_p (-i E_p(t)) [V_p, y]."""
    eom = model.equation_of_motion(liouville_subspace)
    V = [model.dipole_operator(liouville_subspace, p, g)
         for p, g in zip(polarization, geometry)]

    def rhs(t, y):
        dy = eom(t, y)
        for pulse, sign, Vi in zip(pulses, geometry, V):
            E = pulse(t, model.rw_freq)
            if sign == '+':
                E = np.conj(E)
            dy = dy + (-1j * E) * Vi.commutator(y)
        return dy

    y0 = model.thermal_state(liouville_subspace)
    t0 = min(p.t_init for p in pulses)
    tf = max(p.t_final for p in pulses)
    t = (np.arange(t0, tf + time_extra, model.time_step) if times is None
         else tf + times)
    return t, integrate(rhs, y0, t, t0=t0, **kw)


def linear_response(model, path, time_max, initial_state=None,
                    polarization='xx', **kw):
    """response.py:13-43."""
    subspaces = path.split('->')
    if initial_state is None:
        initial_state = model.thermal_state(subspaces[0])
    t = np.arange(0, time_max, model.time_step)
    signal = 0
    for sim_subspace in subspaces[1].split(','):
        V = [model.dipole_operator('{}->{}'.format(a, b), p, s)
             for a, b, p, s in zip(subspaces[:-1], subspaces[1:],
                                   polarization, '+-')]
        V_rho0 = np.apply_along_axis(V[0].commutator, -1, initial_state)
        try:
            eom = model.equation_of_motion(sim_subspace, heisenberg_picture=True)
        except NotImplementedError:
            eom = model.equation_of_motion(sim_subspace)
            signal = signal - integrate(eom, V_rho0, t,
                                        save_func=V[1].expectation_value, **kw)
        else:
            V_Gt = integrate(eom, -V[1].bra_vector, t, **kw)
            signal = signal + np.tensordot(V_rho0, V_Gt, (-1, -1))
    return t, signal


def _symmetrize(t, x, axis):
    """simulate/utils.py:128-151"""
    t, x = np.asarray(t), np.asarray(x)
    T = max(t[-1], -t[0])
    dt = t[1] - t[0]
    n_plus = int((T - t[-1]) / dt) + 1
    n_minus = int((T + t[0]) / dt) + 1
    t_sym = np.concatenate([t[0] - dt * np.arange(1, n_minus)[::-1], t,
                            t[-1] + dt * np.arange(1, n_plus)])
    shape = list(x.shape)
    shape[axis] = t_sym.size
    x_sym = np.zeros(shape, dtype=x.dtype)
    start = np.searchsorted(t_sym, t[0])
    sl = [slice(None)] * x.ndim
    sl[axis] = slice(start, start + t.size)
    x_sym[tuple(sl)] = x
    return t_sym, x_sym


def fourier_transform(t, x, axis=-1, rw_freq=0, unit_convert=1, sign=1):
    """simulate/utils.py:154-219 (angular convention)."""
    from scipy.fftpack import fft, fftshift, ifftshift, fftfreq
    unit_convert = unit_convert / (2 * np.pi)
    t, x = _symmetrize(t, x, axis % np.ndim(x))
    axis = axis % x.ndim
    dt = t[1] - t[0]
    f = fftshift(fftfreq(x.shape[axis], dt * unit_convert))
    X = fftshift(fft(ifftshift(x * dt, axes=axis), axis=axis), axes=axis)
    if sign == 1:
        f = -f[::-1]
        X = np.flip(X, axis=axis)
    return f + rw_freq, X


def absorption_spectra(model, time_max, correlation_decay_time=None,
                       polarization='xx', **kw):
    """response.py:145-154."""
    t, x = linear_response(model, 'gg->eg->gg', time_max,
                           polarization=polarization, **kw)
    if correlation_decay_time is not None:
        x = x * np.exp(-t / correlation_decay_time)
    f, X = fourier_transform(t, -x, rw_freq=model.rw_freq,
                             unit_convert=model.unit_convert)
    return f, X.real


def third_order_response(model, coherence_time_max, population_time_max=None,
                         population_times=None, geometry='-++',
                         polarization='xxxx', include_signal=None, **kw):
    """response.py:267-337."""
    t1 = np.arange(0, coherence_time_max, model.time_step)
    t2 = (np.arange(0, population_time_max, model.time_step)
          if population_times is None else np.asarray(population_times))
    t3 = t1.copy()
    rho0 = model.thermal_state('gg')
    total = 0
    for name, path in THIRD_ORDER_PATHWAYS[geometry].items():
        if include_signal is not None and name not in include_signal:
            continue
        ss = path.split('->')
        V = [model.dipole_operator('{}->{}'.format(a, b), p, s)
             for a, b, p, s in zip(ss[:-1], ss[1:], polarization,
                                   geometry + '-')]
        eom = [model.equation_of_motion(s) for s in ss[1:-1]]
        V_rho0 = V[0].commutator(rho0)
        V_rho1 = integrate(eom[0], V_rho0, t1, save_func=V[1].commutator, **kw)
        V_rho2 = integrate(eom[1], V_rho1, t2, t0=0, save_func=V[2].commutator,
                           **kw)
        try:
            eom_h = model.equation_of_motion(ss[3], heisenberg_picture=True)
        except NotImplementedError:
            total = total + integrate(eom[2], V_rho2, t3,
                                      save_func=V[3].expectation_value, **kw)
        else:
            V_Gt3 = integrate(eom_h, V[3].bra_vector, t3, **kw)
            total = total + np.einsum('ci,abi', V_Gt3, V_rho2)
    return (t1, t2, t3), total
