"""
ORACLE (test infrastructure) -- the integrator loop.

Restates reference ``qspectra/simulate/utils.py:16-109``: a functional wrapper
over ``scipy.integrate.ode('zvode')`` stepping to each requested output time,
with the ``save_func`` epilogue and the leading-axis batch loop.
"""
import numpy as np
import scipy.integrate

__all__ = ['IntegratorError', 'integrate', 'TIGHT', 'counting']

#: integrator settings that define "parity accuracy" 
TIGHT = dict(rtol=1e-10, atol=1e-12, nsteps=100000)


class IntegratorError(Exception):
    pass


def counting(f):
    """Wrap an RHS callable with a call counter (``.calls``)."""
    def g(t, y):
        g.calls += 1
        return f(t, y)
    g.calls = 0
    return g


def _integrate_one(f, y0, t, t0, method_name, save_func, kwargs):
    # utils.py:16-50
    if t0 is None:
        t0 = t[0]
    solver = (scipy.integrate.ode(f) if method_name == 'zvode'
              else scipy.integrate.complex_ode(f))
    solver.set_integrator(method_name, **kwargs)
    solver.set_initial_value(y0, t0)
    first = np.asarray(save_func(y0))
    out = np.empty((len(t),) + first.shape, dtype=first.dtype)
    start = 0
    if t[0] == t0:
        out[0] = first
        start = 1
    for i in range(start, len(t)):
        if not solver.successful():
            raise IntegratorError('integration failed at time {}'.format(t[i]))
        out[i] = save_func(solver.integrate(t[i]))
    return out


def integrate(f, y0, t, t0=None, method_name='zvode', save_func=None,
              **kwargs):
    """utils.py:53-109 -- loops over all leading axes of ``y0``."""
    y0 = np.asarray(y0)
    if save_func is None:
        save_func = lambda x: x
    if y0.ndim == 1:
        return _integrate_one(f, y0, t, t0, method_name, save_func, kwargs)
    return np.array([integrate(f, y, t, t0, method_name, save_func, **kwargs)
                     for y in y0])
