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


# Liouville-space pathways (Abramavicius et al., Chem. Rev. 109, 2350 (2009), figs. 4-6;
# reference response.py:157-159, 252-264)
PUMP_PROBE_PATHWAYS = {'GSB': 'gg->eg->gg',
                       'ESE': 'ee->eg->gg',
                       'ESA': 'ee->fe->ee'}
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


def _parse_pathways(possible_pathways, include_signal):
    selected = [path for name, path in possible_pathways.items()
                if include_signal is None or name in include_signal]
    if not selected:
        raise ValueError('at least one Liouville space pathway must be '
                         'selected, i.e., include_signal must include at least '
                         'one of %r' % list(possible_pathways.keys()))
    return selected


def _interactions(model, subspaces, polarization, transitions):
    """Dipole operators of the successive interactions of a Liouville path."""
    return [model.dipole_operator('{}->{}'.format(a, b), polar, trans)
            for a, b, polar, trans in zip(subspaces[:-1], subspaces[1:], polarization,
                                          transitions)]


# ---------------------------------------------------------------- linear response
# Batchable models (dense generators, HEOM): every independent unit -- ensemble member x
# Cartesian configuration of the 2nd-order isotropic average -- is one column of ONE
# Heisenberg-picture propagation of the detection operator, and the signal
#     S[s, t] = sum_u w_u sum_i (V_0 rho_s)_u[i] (G(t)^T V_1)_u[i]
# is one launch of the K6 contraction kernel.  The reference runs
# ensemble_size x 3 serial ZVODE solves (decorators.py:40-64, 99-125; response.py:13-43).
def _linear_variants(polarization, exact_isotropic_average):
    if not exact_isotropic_average:
        return [(1.0, polarization)]
    from ..polarization import check_polarizations
    weight = float(np.dot(*check_polarizations(polarization, 2))) / 3.0
    return [(weight, p) for p in ('xx', 'yy', 'zz')]


def _linear_response_batched(model, liouv_space_path, time_max, initial_state, polarization,
                             ensemble_size, random_orientations, exact_isotropic_average,
                             member_offset=0, **integrate_kwargs):
    torch = _capi.torch_cuda()
    subspaces = liouv_space_path.split('->')
    t = np.arange(0, time_max, model.time_step)
    opts = {k: integrate_kwargs[k] for k in ('rtol', 'atol', 'rk4_substeps')
            if k in integrate_kwargs}
    method = integrate_kwargs.get('method_name', 'zvode')
    variants = _linear_variants(polarization, exact_isotropic_average)
    nv = len(variants)
    single = ensemble_size is None
    n_members = 1 if single else ensemble_size
    own_basis = getattr(model, 'evolve_basis', 'site') == 'eigen'
    per_member = (random_orientations or own_basis) and not single
    states = None if initial_state is None else np.asarray(initial_state)
    lead = () if states is None or states.ndim == 1 else states.shape[:-1]
    n_states = int(np.prod(lead)) if lead else 1
    total = torch.zeros((n_states, len(t)), dtype=torch.complex128, device='cuda')
    wv = np.array([w for w, _ in variants], dtype=complex)
    chunk = 4096
    for lo in range(0, n_members, chunk):
        E = min(chunk, n_members - lo)
        first = member_offset + lo
        dip = ([model.sample(first + n, random_orientations) for n in range(E)]
               if per_member else [model])
        gens = np.repeat(np.arange(E), nv)
        for sim_subspace in subspaces[1].split(','):
            path = [subspaces[0], sim_subspace, subspaces[2]]
            V = [[_interactions(m, path, pol, '+-') for _, pol in variants] for m in dip]
            rho = [m.thermal_state(subspaces[0]) if states is None else states for m in dip]
            # X[u, s, i] = (V_0 rho_s)[i] of unit u = (member, configuration)
            X = np.array([[np.atleast_2d(V[m][v][0].commutator(rho[m])).reshape(n_states, -1)
                           for v in range(nv)] for m in range(len(dip))])
            if len(dip) != E:
                X = np.broadcast_to(X, (E,) + X.shape[1:])
            X = np.ascontiguousarray(X).reshape(E * nv, n_states, -1)
            bra = np.array([[-V[m][v][1].bra_vector for v in range(nv)]
                            for m in range(len(dip))])
            if len(dip) != E:
                bra = np.broadcast_to(bra, (E,) + bra.shape[1:])
            eom = (model.equation_of_motion(sim_subspace, heisenberg_picture=True) if single
                   else model.ensemble_eom(E, random_orientations, sim_subspace,
                                           heisenberg_picture=True, member0=first))
            G = eom.propagate(np.ascontiguousarray(bra).reshape(E * nv, -1), t, method=method,
                              generators=gens, return_device=True, **opts)
            engine.response_contract(_capi.to_device(X), G,
                                     _capi.to_device(np.tile(wv, E)), total)
    signal = _capi.to_host(total) / n_members
    return t, signal.reshape(lead + (len(t),)) if lead else signal[0]


@optional_ensemble_average
@optional_2nd_order_isotropic_average
def _linear_response_forward(dynamical_model, liouv_space_path, time_max,
                             initial_state=None, polarization='xx', **integrate_kwargs):
    """Member-by-member form for models without a Heisenberg picture (nonlinear equations of
    motion such as ZOFE): propagate V_0 rho forward and read <V_1> at every output time."""
    subspaces = liouv_space_path.split('->')
    rho0 = (dynamical_model.thermal_state(subspaces[0]) if initial_state is None
            else np.asarray(initial_state))
    t = np.arange(0, time_max, dynamical_model.time_step)
    signal = ZeroArray()
    for sim_subspace in subspaces[1].split(','):
        V = _interactions(dynamical_model, [subspaces[0], sim_subspace, subspaces[2]],
                          polarization, '+-')
        signal -= integrate(dynamical_model.equation_of_motion(sim_subspace),
                            V[0].commutator(rho0), t, save_func=V[1].expectation_value,
                            **integrate_kwargs)
    return t, signal


def linear_response(dynamical_model, liouv_space_path, time_max,
                    initial_state=None, polarization='xx', ensemble_size=None,
                    ensemble_random_orientations=False,
                    exact_isotropic_average=False, **integrate_kwargs):
    """Linear response function along a Liouville path 'ab->cd->ef'."""
    if _batchable(dynamical_model):
        return _linear_response_batched(
            dynamical_model, liouv_space_path, time_max, initial_state, polarization,
            ensemble_size, ensemble_random_orientations, exact_isotropic_average,
            **integrate_kwargs)
    return _linear_response_forward(
        dynamical_model, liouv_space_path, time_max, initial_state,
        polarization, ensemble_size=ensemble_size,
        ensemble_random_orientations=ensemble_random_orientations,
        exact_isotropic_average=exact_isotropic_average, **integrate_kwargs)


def absorption_spectra(dynamical_model, time_max, correlation_decay_time=None,
                       polarization='xx', ensemble_size=None,
                       ensemble_random_orientations=False,
                       exact_isotropic_average=False, **integrate_kwargs):
    """(frequencies, real absorption signal) = Fourier transform of the 'gg->eg->gg' linear
    response, optionally apodised with exp(-t / correlation_decay_time)."""
    t, x = linear_response(
        dynamical_model, 'gg->eg->gg', time_max, polarization=polarization,
        ensemble_size=ensemble_size,
        ensemble_random_orientations=ensemble_random_orientations,
        exact_isotropic_average=exact_isotropic_average, **integrate_kwargs)
    if correlation_decay_time is not None:
        x = x * np.exp(-t / correlation_decay_time)
    f, X = fourier_transform(t, -x, rw_freq=dynamical_model.rw_freq,
                             unit_convert=dynamical_model.unit_convert)
    return f, X.real


def impulsive_probe(dynamical_model, state, time_max, polarization='xx',
                    initial_liouv_subspace='gg,ge,eg,ee',
                    include_signal='GSB,ESE,ESA', ensemble_size=None,
                    ensemble_random_orientations=False,
                    exact_isotropic_average=False, **integrate_kwargs):
    """Probe the 2nd-order part of ``state`` (any leading batch axes) with an impulsive
    pulse; returns (frequencies, complex signal field)."""
    model = dynamical_model
    second_order = np.asarray(state) - model.thermal_state(initial_liouv_subspace)
    flat = second_order.reshape(-1, second_order.shape[-1])
    total, t = ZeroArray(), None
    for path in _parse_pathways(PUMP_PROBE_PATHWAYS, include_signal):
        start = path.split('->')[0]
        part = np.array([model.map_between_subspaces(row, initial_liouv_subspace, start)
                         for row in flat]).reshape(second_order.shape[:-1] + (-1,))
        t, signal = linear_response(
            model, path, time_max, part, polarization, ensemble_size=ensemble_size,
            ensemble_random_orientations=ensemble_random_orientations,
            exact_isotropic_average=exact_isotropic_average, **integrate_kwargs)
        total += signal
    return fourier_transform(t, total, rw_freq=model.rw_freq,
                             unit_convert=model.unit_convert)


# ---------------------------------------------------------- third-order response
@optional_ensemble_average
@optional_4th_order_isotropic_average
def _third_order_response_forward(dynamical_model, coherence_time_max,
                                  population_time_max, population_times, geometry,
                                  polarization, include_signal, **integrate_kwargs):
    """Member-by-member form for models without a Heisenberg picture: three forward
    propagations per pathway, the detection operator read at every t3."""
    model = dynamical_model
    t1 = np.arange(0, coherence_time_max, model.time_step)
    t2 = (np.arange(0, population_time_max, model.time_step) if population_times is None
          else np.asarray(population_times, dtype=float))
    t3 = t1.copy()
    rho0 = model.thermal_state('gg')
    total = ZeroArray()
    for path in _parse_pathways(THIRD_ORDER_PATHWAYS[geometry], include_signal):
        subspaces = path.split('->')
        V = _interactions(model, subspaces, polarization, geometry + '-')
        state = V[0].commutator(rho0)
        # the t2 stage propagates all n_t1 columns in one batch, the t3 stage n_t1 x n_t2
        for stage, (grid, t0) in enumerate(((t1, None), (t2, 0), (t3, 0))):
            last = stage == 2
            state = integrate(model.equation_of_motion(subspaces[stage + 1]), state, grid,
                              t0=t0, save_func=(V[3].expectation_value if last
                                                else V[stage + 1].commutator),
                              **integrate_kwargs)
        total += state
    return (t1, t2, t3), total


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
    # configurations that end in the same detection polarisation share their t3 stage (the
    # Heisenberg-propagated detection vector depends on the last field only): order them by it,
    # so that each member's configurations fall into contiguous groups (at most three)
    variants.sort(key=lambda wp: str(wp[1][3]))
    nv = len(variants)
    grp_lo = [v for v in range(nv) if v == 0 or str(variants[v][1][3]) != str(variants[v - 1][1][3])]
    grp_n = [(grp_lo[i + 1] if i + 1 < len(grp_lo) else nv) - lo for i, lo in enumerate(grp_lo)]
    ng = len(grp_lo)
    wv = torch.tensor([w for w, _ in variants], dtype=torch.complex128, device='cuda')
    single = ensemble_size is None
    n_members = 1 if single else ensemble_size
    total = torch.zeros((len(t1), len(t2), len(t3)), dtype=torch.complex128,
                        device='cuda')
    # bound the (units x t1 x t2 x M3) intermediate to ~2 GB per chunk
    n_big = max(len(model.liouville_subspace_index(p.split('->')[3])) for p in paths) \
        * getattr(model, 'ado_count', 1)
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
            # t3 stage: one Heisenberg column per member and detection polarisation
            y3 = np.array([[V[m][lo][3].bra_vector for lo in grp_lo] for m in range(nm)])
            if nm != E:
                y3 = np.broadcast_to(y3, (E,) + y3.shape[1:])
            y3 = np.ascontiguousarray(y3).reshape((E * ng,) + y3.shape[2:])
            out3 = eom_c.propagate(y3, t3, method=method, generators=np.repeat(np.arange(E), ng),
                                   return_device=True, **opts)
            # K6: total[a, b, c] += sum_{e, g} sum_i (sum_{v in g} w_v out2[e, v, a, b, i]) out3[e, g, c, i]
            g_first = (np.arange(E)[:, None] * nv + np.asarray(grp_lo)[None, :]).ravel()
            g_count = np.tile(np.asarray(grp_n), E)
            engine.response_contract(out2.reshape(E * nv, len(t1) * len(t2), -1),
                                     out3.reshape(E * ng, len(t3), -1), wv.repeat(E), total,
                                     group_first=g_first, group_count=g_count)
    if normalize and not single:
        total = total / ensemble_size
    return (t1, t2, t3), total


def _batchable(dynamical_model):
    """Models whose generators live on the device as linear maps with a Heisenberg picture:
    dense Liouvillians and HEOM hierarchies (per-ADO dipole blocks).  ZOFE is nonlinear and
    keeps the member-by-member path."""
    from ..dynamics.liouville_space import LiouvilleSpaceModel
    from ..dynamics.heom import HEOMModel
    return isinstance(dynamical_model, (LiouvilleSpaceModel, HEOMModel))


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
    return _third_order_response_forward(
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
