"""
Lab-frame polarisations and isotropic-average invariants.

Contract: reference ``qspectra/polarization.py`` (polarization_vector :14-36,
4th-order invariant weights :55-70, invariant polarisation lists :73-84, Arvo
random rotation :87-102).
"""
from itertools import product
from numbers import Number

import numpy as np

from .utils import check_random_state

COORD_MAP = {'x': np.array([1, 0, 0]),
             'y': np.array([0, 1, 0]),
             'z': np.array([0, 0, 1])}

FOURTH_ORDER_INVARIANTS = [((0, 1), (2, 3)),
                           ((0, 2), (1, 3)),
                           ((0, 3), (1, 2))]

MAGIC_ANGLE = np.arccos(1 / np.sqrt(3))


def polarization_vector(p):
    """'x'/'y'/'z', an in-plane angle, or any 3-vector -> length-3 float array."""
    try:
        if isinstance(p, str):
            return COORD_MAP[p]
        if isinstance(p, Number):
            return np.array([np.cos(p), np.sin(p), 0])
        vec = np.asanyarray(p, float).reshape(-1)
        if vec.size != 3:
            raise ValueError
        return vec
    except Exception:
        raise ValueError('invalid polarization {}'.format(p))


def check_polarizations(p, length):
    vecs = np.array([polarization_vector(x) for x in p])
    if len(vecs) != length:
        raise ValueError('%s polarizations required' % length)
    return vecs


def invariant_weights_4th_order(polarizations):
    """Weights of <xxyy>, <xyxy>, <xyyx> for four lab-frame polarisations."""
    e = check_polarizations(polarizations, 4)
    cosines = e @ e.T
    prods = np.array([cosines[a] * cosines[b] for a, b in FOURTH_ORDER_INVARIANTS])
    return (5 * np.eye(3) - np.ones((3, 3))) @ prods / 30


def invariant_polarizations(invariant):
    if invariant not in FOURTH_ORDER_INVARIANTS:
        raise ValueError('`invariant` is not one of the three 4th order '
                         'tensor invariants %r' % (FOURTH_ORDER_INVARIANTS,))
    return [''.join(axes) for axes in product('xyz', repeat=4)
            if all(axes[a] == axes[b] for a, b in invariant)]


def random_rotation_matrix(random_state=None):
    """Uniform random rotation (Arvo 1992); consumes exactly ``rand(3)``."""
    x1, x2, x3 = check_random_state(random_state).rand(3)
    theta, phi = 2 * np.pi * x1, 2 * np.pi * x2
    R = np.array([[np.cos(theta), np.sin(theta), 0],
                  [-np.sin(theta), np.cos(theta), 0],
                  [0, 0, 1]])
    v = np.array([np.cos(phi) * np.sqrt(x3), np.sin(phi) * np.sqrt(x3),
                  np.sqrt(1 - x3)])
    householder = np.identity(3) - 2 * np.outer(v, v)
    return -householder @ R
