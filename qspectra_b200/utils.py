"""Small host-side plumbing: per-instance memoisation, accumulator seed,
reproducible RNG (semantics of reference qspectra/utils.py:8-112)."""
import copy
import functools

import numpy as np

_CACHE_ATTR = '_qsx_memo'


class ZeroArray(object):
    """Additive identity of unknown shape: ``z += x`` -> ``x``, ``z -= x`` -> ``-x``."""

    def __iadd__(self, other):
        return other

    def __isub__(self, other):
        return -other


def ndarray_list(arrays, length):
    """Stack an iterable of equally shaped arrays of known count."""
    out = None
    for n, a in enumerate(arrays):
        a = np.asarray(a)
        if out is None:
            out = np.empty((length,) + a.shape, a.dtype)
        out[n] = a
    return out


def imemoize(method):
    """Cache a method's return value on the instance (hashable args only)."""
    @functools.wraps(method)
    def wrapper(self, *args, **kwargs):
        memo = self.__dict__.setdefault(_CACHE_ATTR, {})
        key = (method.__name__, args, frozenset(kwargs.items()))
        try:
            return memo[key]
        except KeyError:
            val = memo[key] = method(self, *args, **kwargs)
            return val
    return wrapper


def memoized_property(method):
    return property(imemoize(method))


def copy_with_new_cache(obj):
    """Shallow copy that forgets everything memoised on the original."""
    new = copy.copy(obj)
    new.__dict__.pop(_CACHE_ATTR, None)
    return new


def check_random_state(seed):
    if seed is None:
        return np.random.mtrand._rand
    if isinstance(seed, np.random.RandomState):
        return seed
    return np.random.RandomState(seed)


def simple_repr(obj, names):
    body = ', '.join('%s=%r' % (k, getattr(obj, k, None)) for k in names)
    return '%s(%s)' % (type(obj).__name__, body)
