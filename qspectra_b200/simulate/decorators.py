"""
The three embarrassingly parallel outer loops of the reference
(simulate/decorators.py:40-125) as generic decorators: disorder ensemble,
2nd-order and 4th-order isotropic averages.  The engine's own entry points
batch these axes on the device where they can (see eom.py / response.py);
the decorators remain the general, always-correct form and are what
third-party DynamicalModel code would use.
"""
from functools import wraps
import inspect

import numpy as np

from ..polarization import (check_polarizations, invariant_weights_4th_order,
                            invariant_polarizations, FOURTH_ORDER_INVARIANTS)
from ..utils import ZeroArray


def _get_call_args(func, *args, **kwargs):
    """Bind the call to a flat keyword dict (so ``func(**call_args)`` is the
    same call) -- the reference uses the removed ``inspect.getargspec``."""
    spec = inspect.getfullargspec(func)
    if spec.varargs is not None:
        raise NotImplementedError('%s cannot include positional-only '
                                  'arguments (i.e., of the form *args)' % func)
    call_args = inspect.getcallargs(func, *args, **kwargs)
    if spec.varkw is not None:
        call_args.update(call_args.pop(spec.varkw, {}))
    return call_args


def optional_ensemble_average(func):
    @wraps(func)
    def wrapper(dynamical_model, *args, **kwargs):
        ensemble_size = kwargs.pop('ensemble_size', None)
        random_orientations = kwargs.pop('ensemble_random_orientations', False)
        if ensemble_size is None:
            return func(dynamical_model, *args, **kwargs)
        total = ZeroArray()
        for member in dynamical_model.sample_ensemble(ensemble_size,
                                                      random_orientations):
            ticks, signal = func(member, *args, **kwargs)
            total += signal
        total /= ensemble_size
        return ticks, total
    return wrapper


def optional_4th_order_isotropic_average(func):
    @wraps(func)
    def wrapper(*args, **kwargs):
        if not kwargs.pop('exact_isotropic_average', False):
            return func(*args, **kwargs)
        kwargs = _get_call_args(func, *args, **kwargs)
        weights = invariant_weights_4th_order(kwargs.pop('polarization'))
        signals, total, t = {}, ZeroArray(), None
        for invariant, weight in zip(FOURTH_ORDER_INVARIANTS, weights):
            if weight > 1e-8:
                for p in invariant_polarizations(invariant):
                    if p not in signals:
                        t, signals[p] = func(polarization=p, **kwargs)
                    total += weight * signals[p]
        return t, total
    return wrapper


def optional_2nd_order_isotropic_average(func):
    @wraps(func)
    def wrapper(*args, **kwargs):
        if not kwargs.pop('exact_isotropic_average', False):
            return func(*args, **kwargs)
        kwargs = _get_call_args(func, *args, **kwargs)
        polarizations = check_polarizations(kwargs.pop('polarization'), 2)
        weight = np.dot(*polarizations)
        total, t = ZeroArray(), None
        for p in ('xx', 'yy', 'zz'):
            t, signal = func(polarization=p, **kwargs)
            total += signal
        total *= weight / 3.0
        return t, total
    return wrapper
