"""
Response-function spectroscopy on the device integrators.

Contract: reference ``qspectra/simulate/response.py`` -- ``linear_response``
:13-100, ``absorption_spectra`` :103-154, ``impulsive_probe`` :174-244,
Liouville pathway tables :157-159 / :252-264, ``third_order_response``
:267-427, ``two_dimensional_spectra`` :430-455.

Each ``integrate`` call below is one fused device propagation of a batch of
columns: e.g. the t2 stage of the third-order response propagates all n_t1
columns of V_rho1 under one generator in a single kernel, where the reference
runs n_t1 serial ZVODE solves (utils.py:103-109).
"""
import numpy as np

from .decorators import (optional_ensemble_average,
                         optional_2nd_order_isotropic_average,
                         optional_4th_order_isotropic_average)
from .utils import integrate, fourier_transform
from ..utils import ZeroArray
from .. import _capi, engine


@optional_ensemble_average
@optional_2nd_order_isotropic_average
def _linear_response(dynamical_model, liouv_space_path, time_max,
                     initial_state=None, polarization='xx', **integrate_kwargs):
    subspaces = liouv_space_path.split('->')
    if initial_state is None:
        initial_state = dynamical_model.thermal_state(subspaces[0])
    initial_state = np.asarray(initial_state)
    t = np.arange(0, time_max, dynamical_model.time_step)
    signal = ZeroArray()
    for sim_subspace in subspaces[1].split(','):
        V = [dynamical_model.dipole_operator('{}->{}'.format(a, b), polar, trans)
             for a, b, polar, trans in zip(subspaces[:-1], subspaces[1:],
                                           polarization, '+-')]
        V_rho0 = V[0].commutator(initial_state)
        try:
            # Heisenberg picture: one propagation of the detection operator
            # serves every initial state
            eom = dynamical_model.equation_of_motion(sim_subspace,
                                                     heisenberg_picture=True)
        except NotImplementedError:
            eom = dynamical_model.equation_of_motion(sim_subspace)
            signal -= integrate(eom, V_rho0, t,
                                save_func=V[1].expectation_value,
                                **integrate_kwargs)
        else:
            V_Gt = integrate(eom, -V[1].bra_vector, t, **integrate_kwargs)
            signal += np.tensordot(V_rho0, V_Gt, (-1, -1))
    return (t, signal)


def linear_response(dynamical_model, liouv_space_path, time_max,
                    initial_state=None, polarization='xx', ensemble_size=None,
                    ensemble_random_orientations=False,
                    exact_isotropic_average=False, **integrate_kwargs):
    """Linear response function along a Liouville path 'ab->cd->ef'."""
    return _linear_response(
        dynamical_model, liouv_space_path, time_max, initial_state,
        polarization, ensemble_size=ensemble_size,
        ensemble_random_orientations=ensemble_random_orientations,
        exact_isotropic_average=exact_isotropic_average, **integrate_kwargs)


def absorption_spectra(dynamical_model, time_max, correlation_decay_time=None,
                       polarization='xx', ensemble_size=None,
                       ensemble_random_orientations=False,
                       exact_isotropic_average=False, **integrate_kwargs):
    """(frequencies, real absorption signal)."""
    (t, x) = linear_response(
        dynamical_model, 'gg->eg->gg', time_max, polarization=polarization,
        ensemble_size=ensemble_size,
        ensemble_random_orientations=ensemble_random_orientations,
        exact_isotropic_average=exact_isotropic_average, **integrate_kwargs)
    if correlation_decay_time is not None:
        x = x * np.exp(-t / correlation_decay_time)
    (f, X) = fourier_transform(t, -x, rw_freq=dynamical_model.rw_freq,
                               unit_convert=dynamical_model.unit_convert)
    return (f, X.real)


PUMP_PROBE_PATHWAYS = {'GSB': 'gg->eg->gg',
                       'ESE': 'ee->eg->gg',
                       'ESA': 'ee->fe->ee'}


def _parse_pathways(possible_pathways, include_signal):
    selected = [path for name, path in possible_pathways.items()
                if include_signal is None or name in include_signal]
    if not selected:
        raise ValueError('at least one Liouville space pathway must be '
                         'selected, i.e., include_signal must include at least '
                         'one of %r' % list(possible_pathways.keys()))
    return selected


def impulsive_probe(dynamical_model, state, time_max, polarization='xx',
                    initial_liouv_subspace='gg,ge,eg,ee',
                    include_signal='GSB,ESE,ESA', ensemble_size=None,
                    ensemble_random_orientations=False,
                    exact_isotropic_average=False, **integrate_kwargs):
    """Probe the 2nd-order part of ``state`` with an impulsive probe pulse;
    returns (frequencies, complex signal field)."""
    state = np.asarray(state)
    initial_state = state - dynamical_model.thermal_state(initial_liouv_subspace)
    total_signal = ZeroArray()
    for path in _parse_pathways(PUMP_PROBE_PATHWAYS, include_signal):
        first = path.split('->')[0]
        portion = np.apply_along_axis(
            lambda s: dynamical_model.map_between_subspaces(
                s, initial_liouv_subspace, first), -1, initial_state)
        (t, signal) = linear_response(
            dynamical_model, path, time_max, portion, polarization,
            ensemble_size=ensemble_size,
            ensemble_random_orientations=ensemble_random_orientations,
            exact_isotropic_average=exact_isotropic_average, **integrate_kwargs)
        total_signal += signal
    return fourier_transform(t, total_signal, rw_freq=dynamical_model.rw_freq,
                             unit_convert=dynamical_model.unit_convert)


# Liouville-space pathways of Abramavicius et al., Chem. Rev. 109, 2350 (2009),
# figs. 4-6 (same tables as the reference)
THIRD_ORDER_PATHWAYS = {
    '-++': {'ESE': 'gg->ge->ee->eg->gg',      # photon echo
            'GSB': 'gg->ge->gg->eg->gg',
            'ESA': 'gg->ge->ee->fe->ee'},
    '+-+': {'ESE': 'gg->eg->ee->eg->gg',      # non-rephasing
            'GSB': 'gg->eg->gg->eg->gg',
            'ESA': 'gg->eg->ee->fe->ee'},
    '++-': {'ESA1': 'gg->eg->fg->fe->ee',     # double-quantum coherence
            'ESA2': 'gg->eg->fg->eg->gg'},
}


@optional_ensemble_average
@optional_4th_order_isotropic_average
def _third_order_response(dynamical_model, coherence_time_max,
                          population_time_max, population_times, geometry,
                          polarization, include_signal, **integrate_kwargs):
    t1 = np.arange(0, coherence_time_max, dynamical_model.time_step)
    t2 = (np.arange(0, population_time_max, dynamical_model.time_step)
          if population_times is None
          else np.asarray(population_times, dtype=float))
    t3 = np.arange(0, coherence_time_max, dynamical_model.time_step)

    initial_state = dynamical_model.thermal_state('gg')
    total_signal = ZeroArray()
    for path in _parse_pathways(THIRD_ORDER_PATHWAYS[geometry], include_signal):
        subspaces = path.split('->')
        V = [dynamical_model.dipole_operator('{}->{}'.format(a, b), polar, trans)
             for a, b, polar, trans in zip(subspaces[:-1], subspaces[1:],
                                           polarization, geometry + '-')]
        eom = [dynamical_model.equation_of_motion(s) for s in subspaces[1:-1]]
        V_rho0 = V[0].commutator(initial_state)
        # t1: one column; t2: n_t1 columns under one generator (batched)
        V_rho1 = integrate(eom[0], V_rho0, t1, save_func=V[1].commutator,
                           **integrate_kwargs)
        V_rho2 = integrate(eom[1], V_rho1, t2, t0=0, save_func=V[2].commutator,
                           **integrate_kwargs)
        try:
            eom_heisen = dynamical_model.equation_of_motion(
                subspaces[3], heisenberg_picture=True)
        except NotImplementedError:
            total_signal += integrate(eom[2], V_rho2, t3,
                                      save_func=V[3].expectation_value,
                                      **integrate_kwargs)
        else:
            V_Gt3 = integrate(eom_heisen, V[3].bra_vector, t3,
                              **integrate_kwargs)
            total_signal += np.einsum('ci,abi', V_Gt3, V_rho2)
    return (t1, t2, t3), total_signal


def _polarization_variants(polarization, exact_isotropic_average):
    """[(weight, polarization)] whose weighted sum is the requested signal: the
    polarization itself, or the Cartesian configurations of the fourth-order
    isotropic average with their invariant weights summed per configuration
    (decorators.py:64-101, same 1e-8 weight cut-off)."""
    if not exact_isotropic_average:
        return [(1.0, polarization)]
    from collections import OrderedDict
    from ..polarization import (FOURTH_ORDER_INVARIANTS, invariant_polarizations,
                                invariant_weights_4th_order)
    acc = OrderedDict()
    for invariant, weight in zip(FOURTH_ORDER_INVARIANTS,
                                 invariant_weights_4th_order(polarization)):
        if weight > 1e-8:
            for p in invariant_polarizations(invariant):
                acc[p] = acc.get(p, 0.0) + weight
    return [(w, p) for p, w in acc.items()]


def _third_order_response_batched(dynamical_model, coherence_time_max,
                                  population_time_max, population_times,
                                  geometry, polarization, include_signal,
                                  ensemble_size, random_orientations,
                                  member_offset, normalize,
                                  exact_isotropic_average=False,
                                  **integrate_kwargs):
    """Third-order response of a dense-generator model with ALL independent units
    on the device at once: ensemble members (`ensemble_size`, or the model itself
    when None) x polarisation configurations of the isotropic average (up to 21).
    Per pathway three batched propagations -- t1: one column per unit, t2: n_t1
    columns per unit under that member's generator, t3: one Heisenberg column per
    unit -- and one weighted contraction on the device.  The reference runs
    ensemble_size x 21 x (1 + n_t1 + 1) serial ZVODE solves per pathway
    (decorators.py:55-60, 86-92; utils.py:103-109).  Returns a CUDA tensor."""
    from .. import _capi
    torch = _capi.torch_cuda()
    model = dynamical_model
    t1 = np.arange(0, coherence_time_max, model.time_step)
    t2 = (np.arange(0, population_time_max, model.time_step)
          if population_times is None
          else np.asarray(population_times, dtype=float))
    t3 = t1.copy()
    opts = {k: integrate_kwargs[k] for k in ('rtol', 'atol', 'rk4_substeps')
            if k in integrate_kwargs}
    method = integrate_kwargs.get('method_name', 'zvode')
    paths = _parse_pathways(THIRD_ORDER_PATHWAYS[geometry], include_signal)
    variants = _polarization_variants(polarization, exact_isotropic_average)
    nv = len(variants)
    wv = torch.tensor([w for w, _ in variants], dtype=torch.complex128, device='cuda')
    single = ensemble_size is None
    n_members = 1 if single else ensemble_size
    total = torch.zeros((len(t1), len(t2), len(t3)), dtype=torch.complex128,
                        device='cuda')
    # bound the (units x t1 x t2 x M3) intermediate to ~2 GB per chunk
    n_big = max(len(model.liouville_subspace_index(p.split('->')[3])) for p in paths)
    chunk = max(1, int(2e9 // (len(t1) * len(t2) * n_big * 16 * nv)))
    rho0 = model.thermal_state('gg')
    for lo in range(0, n_members, chunk):
        E = min(chunk, n_members - lo)
        first = member_offset + lo
        # dipole operators per member when the members differ in them: random orientations,
        # or evolution in each member's own eigenbasis (reference decorators.py:55-60 builds
        # every member's operators from the sampled model)
        own_basis = getattr(model, 'evolve_basis', 'site') == 'eigen'
        dip = ([model.sample(first + n, random_orientations) for n in range(E)]
               if (random_orientations or own_basis) and not single else [model])
        nm = len(dip)
        gens = np.repeat(np.arange(E), nv)                  # generator of unit (e, v)
        sidx = (np.arange(E * nv) if nm == E else np.tile(np.arange(nv), E))
        eoms = {}

        def eom_for(subspace, heisenberg=False):
            # pathways share stage subspaces: build each batch of generators once
            key = (subspace, heisenberg)
            if key not in eoms:
                eoms[key] = (model.equation_of_motion(subspace, heisenberg_picture=heisenberg)
                             if single else
                             model.ensemble_eom(E, random_orientations, subspace,
                                                heisenberg_picture=heisenberg,
                                                member0=first))
            return eoms[key]

        for path in paths:
            ss = path.split('->')
            # V[m][v][i]: dipole operator of interaction i for member m, configuration v
            V = [[[m.dipole_operator('{}->{}'.format(a, b), polar, trans)
                   for a, b, polar, trans in zip(ss[:-1], ss[1:], pol, geometry + '-')]
                  for _, pol in variants] for m in dip]

            def per_unit(f):
                """(E * nv, ...) array of f(V[m][v]) with members broadcast if shared"""
                a = np.array([[f(V[m][v]) for v in range(nv)] for m in range(nm)])
                if nm != E:
                    a = np.broadcast_to(a, (E,) + a.shape[1:])
                return np.ascontiguousarray(a).reshape((E * nv,) + a.shape[2:])

            stack = lambda i: np.array([[V[m][v][i].commutator.matrix for v in range(nv)]
                                        for m in range(nm)]).reshape((nm * nv,) + V[0][0][i].commutator.matrix.shape)
            eom_a, eom_b, eom_c = eom_for(ss[1]), eom_for(ss[2]), eom_for(ss[3], True)
            out1 = eom_a.propagate(per_unit(lambda v: v[0].commutator(rho0)), t1,
                                   method=method, save=stack(1), generators=gens,
                                   save_index=sidx, return_device=True, **opts)
            out2 = eom_b.propagate(out1.reshape(E * nv * len(t1), -1), t2, t0=0,
                                   method=method, save=stack(2),
                                   generators=np.repeat(gens, len(t1)),
                                   save_index=np.repeat(sidx, len(t1)),
                                   return_device=True, **opts)
            out3 = eom_c.propagate(per_unit(lambda v: v[3].bra_vector), t3, method=method,
                                   generators=gens, return_device=True, **opts)
            # K6: total[a, b, c] += sum_{e, v} w_v sum_i out2[e, v, a, b, i] out3[e, v, c, i]
            engine.response_contract(out2.reshape(E * nv, len(t1) * len(t2), -1),
                                     out3.reshape(E * nv, len(t3), -1), wv.repeat(E), total)
    if normalize and not single:
        total = total / ensemble_size
    return (t1, t2, t3), total


def _batchable(dynamical_model):
    from ..dynamics.liouville_space import LiouvilleSpaceModel
    return isinstance(dynamical_model, LiouvilleSpaceModel)


def third_order_response(dynamical_model, coherence_time_max,
                         population_time_max=None, population_times=None,
                         geometry='-++', polarization='xxxx',
                         include_signal=None, ensemble_size=None,
                         ensemble_random_orientations=False,
                         exact_isotropic_average=False, **integrate_kwargs):
    """Third-order response ((t1, t2, t3), signal[t1, t2, t3]) in the rotating
    wave approximation, summed over the selected Liouville pathways.  Disorder
    ensembles of dense-generator models are propagated as one device batch."""
    if _batchable(dynamical_model):
        ticks, total = _third_order_response_batched(
            dynamical_model, coherence_time_max, population_time_max,
            population_times, geometry, polarization, include_signal,
            ensemble_size, ensemble_random_orientations,
            integrate_kwargs.pop('member_offset', 0), True,
            exact_isotropic_average=exact_isotropic_average, **integrate_kwargs)
        return ticks, _capi.to_host(total)
    return _third_order_response(
        dynamical_model, coherence_time_max, population_time_max,
        population_times, geometry, polarization, include_signal,
        ensemble_size=ensemble_size,
        ensemble_random_orientations=ensemble_random_orientations,
        exact_isotropic_average=exact_isotropic_average, **integrate_kwargs)


def two_dimensional_spectra(dynamical_model, coherence_time_max,
                            population_time_max=None, population_times=None,
                            geometry='-++', polarization='xxxx',
                            include_signal=None, ensemble_size=None,
                            ensemble_random_orientations=False,
                            exact_isotropic_average=False,
                            **integrate_kwargs):
    """2D spectrum: Fourier transform of the third-order response over t1
    (sign -1) and t3."""
    from .. import _capi
    if _batchable(dynamical_model):
        # the (ensemble- and orientation-summed) signal never leaves the device before the transforms
        (t1, t2, t3), X = _third_order_response_batched(
            dynamical_model, coherence_time_max, population_time_max,
            population_times, geometry, polarization, include_signal,
            ensemble_size, ensemble_random_orientations,
            integrate_kwargs.pop('member_offset', 0), True,
            exact_isotropic_average=exact_isotropic_average, **integrate_kwargs)
    else:
        (t1, t2, t3), X = third_order_response(
            dynamical_model, coherence_time_max, population_time_max,
            population_times, geometry, polarization, include_signal,
            ensemble_size, ensemble_random_orientations, exact_isotropic_average,
            **integrate_kwargs)
        X = _capi.to_device(np.ascontiguousarray(X, dtype=complex))
    rw_freq = dynamical_model.rw_freq
    unit_convert = dynamical_model.unit_convert
    f1, X_ftt = fourier_transform(t1, X, 0, rw_freq=rw_freq, sign=-1,
                                  unit_convert=unit_convert)
    f3, X_ftf = fourier_transform(t3, X_ftt, 2, rw_freq=rw_freq,
                                  unit_convert=unit_convert)
    if not isinstance(X_ftf, np.ndarray):
        X_ftf = _capi.to_host(X_ftf)
    return (f1, t2, f3), X_ftf
