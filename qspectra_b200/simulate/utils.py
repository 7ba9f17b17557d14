"""
``integrate`` with the reference's signature, executed by the fused device
integrators, plus the FFT post-processing helpers.

Contract: reference ``qspectra/simulate/utils.py`` -- ``integrate`` :53-109
(leading axes of ``y0`` are batch axes; ``y[0] = save_func(y0)`` when
``t[0] == t0``, otherwise the solver first runs from ``t0`` to ``t[0]``),
``_symmetrize`` :128-151, ``fourier_transform`` :154-219, ``bound_signal``
:222-250.
"""
import numpy as np

from .. import _capi
from ..engine import DeviceEOM, LinearMap, IntegratorError  # noqa: F401


class DrivenEOM(object):
    """f(t, y) = L y + sum_p (-i E_p(t)) C_p y: a device generator plus pulse
    terms (reference simulate/eom.py:87-94).  Built by simulate_with_fields."""

    def __init__(self, eom, pulses, pulse_ops):
        self.eom = eom
        self.pulses = pulses          # [(scale, detuning, t_peak, inv2s2, conj)]
        self.pulse_ops = pulse_ops    # (n_sets, n_pulses, d, d)
        self.dim = eom.dim

    def __call__(self, t, y):
        dy = self.eom(t, y)
        ops = np.asarray(self.pulse_ops)
        ops = ops[0] if ops.ndim == 4 else ops
        for (scale, det, tp, inv, conj), Cp in zip(self.pulses, ops):
            E = scale * np.exp(1j * det * (t - tp) - (t - tp) ** 2 * inv)
            if conj:
                E = np.conj(E)
            dy = dy + (-1j * E) * Cp.dot(y)
        return dy


def integrate(f, y0, t, t0=None, method_name='zvode', f_params=None,
              save_func=None, **kwargs):
    """Solve dy/dt = f(t, y) on the GPU and return ``save_func(y(t_i))``.

    ``f`` must come from a GPU model's ``equation_of_motion`` (a DeviceEOM) or
    be a DrivenEOM; arbitrary Python callables cannot run inside the fused
    integrator and are rejected (there is no CPU fallback).

    method_name : 'zvode' (default -> engine default: adaptive Taylor for
        constant generators, DOPRI5 otherwise), 'taylor', 'rk4', 'dopri5'.
    kwargs : rtol, atol (as for scipy's solvers), rk4_substeps, generators
        (per-column generator index for batched ensembles).  `nsteps` is
        accepted silently (the device integrators have no step budget per
        interval); any other scipy option (max_step, first_step, order, ...)
        has no counterpart here and raises a warning instead of being dropped
        without notice.  Propagator stepping ('expm', the default for constant
        dense generators on uniform grids) is accurate to ~1e-13 regardless of
        rtol/atol; pass method_name='taylor' to have rtol control the series.
        A non-finite state makes every device integrator fail with
        IntegratorError.
    """
    ignored = sorted(set(kwargs) - {'rtol', 'atol', 'rk4_substeps', 'generators', 'nsteps'})
    if ignored:
        import warnings
        warnings.warn('integrate: options %s have no device counterpart and are ignored'
                      % ', '.join(ignored), RuntimeWarning, stacklevel=2)
    if f_params:
        raise NotImplementedError('f_params are not supported on the device')
    pulses = pulse_ops = None
    eom = f
    if isinstance(f, DrivenEOM):
        eom, pulses, pulse_ops = f.eom, f.pulses, f.pulse_ops
    if not isinstance(eom, DeviceEOM):
        raise TypeError('integrate needs an equation of motion produced by a '
                        'qspectra_b200 model (got %r); host callables cannot '
                        'run in the fused GPU integrator' % (f,))
    y0 = np.asarray(y0, dtype=complex)
    t = np.asarray(t, dtype=float)
    lead = y0.shape[:-1]
    opts = {k: kwargs[k] for k in ('rtol', 'atol', 'rk4_substeps', 'generators')
            if k in kwargs}
    host_post = None
    save = None
    if save_func is not None:
        if isinstance(save_func, LinearMap):
            if save_func.ado0_only:
                save, host_post = ('ado0',), save_func
            else:
                save = save_func
        elif isinstance(save_func, tuple):
            save = save_func
        else:
            host_post = save_func       # user callable: applied on the host
    out = eom.propagate(y0.reshape(-1, y0.shape[-1]), t, t0=t0,
                        method=method_name, save=save, pulses=pulses,
                        pulse_ops=pulse_ops, **opts)
    if host_post is not None:
        if isinstance(host_post, LinearMap):     # HEOM expectation value on ADO 0
            out = np.tensordot(out, host_post.matrix, axes=(-1, -1))
        else:
            first = np.asarray(host_post(out[0, 0]))
            res = np.empty(out.shape[:2] + first.shape, dtype=first.dtype)
            for b in range(out.shape[0]):
                for i in range(out.shape[1]):
                    res[b, i] = host_post(out[b, i])
            out = res
    elif isinstance(save, LinearMap) and save.matrix.ndim == 1:
        out = out[..., 0]
    return out.reshape(lead + out.shape[1:])


# ------------------------------------------------------------ post-processing
def slice_along_axis(start=None, stop=None, step=None, axis=0, ndim=1):
    axis = axis % ndim
    return tuple(slice(start, stop, step) if n == axis else slice(None)
                 for n in range(ndim))


def is_constant(x, atol=1e-7, positive=None):
    x = np.asarray(x)
    flat = np.max(np.abs(x - x[0])) < atol
    if positive is None:
        return bool(flat)
    return bool(flat and np.all((x > 0) == positive))


def _symmetrize(t, x, axis=-1):
    """Zero-pad so that the time axis is symmetric around t = 0."""
    t = np.asarray(t)
    x = np.asarray(x)
    if not is_constant(np.diff(t), positive=True):
        raise ValueError('sample times must differ by a positive constant')
    axis = axis % x.ndim
    T = max(t[-1], -t[0])
    dt = t[1] - t[0]
    n_after = int((T - t[-1]) / dt) + 1
    n_before = int((T + t[0]) / dt) + 1
    t_sym = np.concatenate([t[0] - dt * np.arange(1, n_before)[::-1], t,
                            t[-1] + dt * np.arange(1, n_after)])
    shape = list(x.shape)
    shape[axis] = t_sym.size
    x_sym = np.zeros(shape, dtype=x.dtype)
    start, end = (np.searchsorted(t_sym, ti) for ti in (t[0], t[-1]))
    x_sym[slice_along_axis(start, end + 1, axis=axis, ndim=x.ndim)] = x
    return t_sym, x_sym


def fourier_transform(t, x, axis=-1, rw_freq=0, unit_convert=1, sign=1,
                      convention='angular'):
    """X(w) = int exp(+-i (w - w0) t) x(t) dt by FFT of the zero-padded,
    symmetrised signal; returns (frequencies, X)."""
    t = np.asarray(t)
    on_device = not isinstance(x, np.ndarray) and hasattr(x, 'is_cuda') and x.is_cuda
    if on_device and _dft_grid(t):
        return _fourier_transform_device(t, x, axis, rw_freq, unit_convert, sign,
                                         convention)
    x = _capi.to_host(x) if on_device else np.asarray(x)
    if t.ndim != 1:
        raise ValueError('t must be one dimensional')
    if t.size != x.shape[axis]:
        raise ValueError('t must have the same length as the shape of x along '
                         'the given axis')
    if sign not in (-1, +1):
        raise ValueError('sign must be +1 or -1')
    if convention == 'angular':
        unit_convert = unit_convert / (2 * np.pi)
    elif convention != 'linear':
        raise ValueError("convention must be 'angular' or 'linear'")
    t, x = _symmetrize(t, x, axis)
    axis = axis % x.ndim
    dt = t[1] - t[0]
    f = np.fft.fftshift(np.fft.fftfreq(x.shape[axis], dt * unit_convert))
    X = np.fft.fftshift(np.fft.fft(np.fft.ifftshift(x * dt, axes=axis),
                                   axis=axis), axes=axis)
    if sign == 1:
        f = -f[::-1]
        X = np.flip(X, axis=axis)
    return f + rw_freq, X


def _dft_grid(t):
    """True when `_symmetrize` would pad t = 0, dt, ... to exactly 2n - 1 points
    (n - 1 before, none after) -- the case kernel K7 evaluates directly."""
    if t.ndim != 1 or t.size < 2 or t.size > 4096 or t[0] != 0:
        return False
    if not is_constant(np.diff(t), positive=True):
        return False
    dt = t[1] - t[0]
    T = max(t[-1], -t[0])
    return (int((T - t[-1]) / dt) + 1 == 1) and (int((T + t[0]) / dt) + 1 == t.size)


def _fourier_transform_device(t, x_dev, axis, rw_freq, unit_convert, sign, convention):
    """`fourier_transform` for a CUDA tensor: same frequencies, transform by the
    device kernel (engine.fourier_transform); the result stays on the device."""
    from .. import engine
    if t.size != x_dev.shape[axis]:
        raise ValueError('t must have the same length as the shape of x along '
                         'the given axis')
    if sign not in (-1, +1):
        raise ValueError('sign must be +1 or -1')
    if convention == 'angular':
        unit_convert = unit_convert / (2 * np.pi)
    elif convention != 'linear':
        raise ValueError("convention must be 'angular' or 'linear'")
    dt = t[1] - t[0]
    f = np.fft.fftshift(np.fft.fftfreq(2 * t.size - 1, dt * unit_convert))
    if sign == 1:
        f = -f[::-1]
    return f + rw_freq, engine.fourier_transform(x_dev, axis, dt, sign)


def bound_signal(ticks, signal, bounds, axis=0):
    ticks = np.asarray(ticks)
    signal = np.asarray(signal)
    if signal.shape[axis] != len(ticks):
        raise ValueError('ticks must have same shape as signal along given '
                         'axis')
    i0, i1 = sorted(int(np.argmin(np.abs(ticks - b))) for b in bounds)
    return ticks[i0:i1 + 1], signal[slice_along_axis(i0, i1 + 1, axis=axis,
                                                      ndim=signal.ndim)]
