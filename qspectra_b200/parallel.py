"""
Multi-GPU sharding of the embarrassingly parallel outer loops:
disorder members (reference simulate/decorators.py:55-60) are block-partitioned
over the ranks of a ``torch.distributed`` job -- one process per GPU -- each
rank replays its own members' seeded disorder streams (no communication), and
the partial ensemble sums are combined with ONE reduce (NCCL over NVLink on
GPUs, gloo in the CPU tests).  A single HEOM trajectory does not shard
("replicas only").
"""
import numpy as np


def shard_members(ensemble_size, rank, world_size):
    """(first member, count) of this rank's contiguous block; the blocks tile
    range(ensemble_size) exactly and differ in size by at most one."""
    if not 0 <= rank < world_size:
        raise ValueError('rank out of range')
    base, extra = divmod(int(ensemble_size), int(world_size))
    count = base + (1 if rank < extra else 0)
    first = rank * base + min(rank, extra)
    return first, count


def world():
    """(rank, world_size) of the current torch.distributed job, (0, 1) if none."""
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
    except ImportError:
        pass
    return 0, 1


def reduce_sum(tensor, dst=None):
    """Sum a (complex or real) tensor over all ranks: all-reduce when ``dst`` is
    None, otherwise reduce to rank ``dst``.  Complex tensors are reduced through
    their real view (NCCL has no complex type)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return tensor
    buf = torch.view_as_real(tensor) if tensor.is_complex() else tensor
    buf = buf.contiguous()
    if dst is None:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM)
    else:
        dist.reduce(buf, dst=dst, op=dist.ReduceOp.SUM)
    return torch.view_as_complex(buf) if tensor.is_complex() else buf


def sharded_ensemble_mean(partial_sum_fn, ensemble_size, dst=None):
    """Generic driver: ``partial_sum_fn(first, count)`` returns this rank's SUM
    over members first .. first+count-1 as a tensor (or None when count == 0 is
    not supported by the caller); returns the global mean."""
    import torch
    rank, size = world()
    first, count = shard_members(ensemble_size, rank, size)
    part = partial_sum_fn(first, count)
    if not isinstance(part, torch.Tensor):
        part = torch.from_numpy(np.ascontiguousarray(part))
    total = reduce_sum(part, dst)
    return total / ensemble_size


def simulate_dynamics_sharded(dynamical_model, initial_state, duration=None,
                              times=None, liouville_subspace='ee',
                              ensemble_size=None, dst=None, **integrate_kwargs):
    """``simulate_dynamics`` with the disorder ensemble sharded over the ranks
    of the current torch.distributed job; every rank (or only ``dst``) receives
    the ensemble-averaged density matrices."""
    import torch
    from . import _capi, engine
    from .simulate.eom import ensemble_mean
    rank, size = world()
    first, count = shard_members(ensemble_size, rank, size)
    initial_state = np.asarray(initial_state)
    if initial_state.ndim == 1:
        initial_state = np.outer(initial_state.conj(), initial_state)
    t = (np.arange(0, duration, dynamical_model.time_step)
         if times is None else np.asarray(times, dtype=float))
    y0 = dynamical_model.density_matrix_to_state_vector(initial_state,
                                                        liouville_subspace)
    save = dynamical_model.dynamics_save
    if count > 0:
        eom = dynamical_model.ensemble_eom(count, False, liouville_subspace,
                                           member0=first)
        part = ensemble_mean(eom, y0, t, count, save, return_device=True,
                             scale=1.0 / ensemble_size, **integrate_kwargs)
    else:
        # more ranks than members: contribute zeros of the saved width ('ado0' keeps the
        # density-matrix part of a HEOM / ZOFE state: one entry per subspace element)
        torch_ = _capi.torch_cuda()
        if save is None:
            dim = y0.size
        elif hasattr(dynamical_model, 'lspace_model'):
            dim = dynamical_model.lspace_model.liouville_subspace_index(liouville_subspace).size
        else:
            dim = dynamical_model.hamiltonian.n_states(dynamical_model.hilbert_subspace) ** 2
        part = torch_.zeros((len(t), dim), dtype=torch_.complex128, device='cuda')
    total = reduce_sum(part, dst)
    states = _capi.to_host(total)
    return t, dynamical_model.saved_states_to_density_matrix(states)


def ensemble_signal_sharded(func, dynamical_model, ensemble_size,
                            random_orientations=False, dst=None):
    """Ensemble average of any simulate-layer signal with the members sharded
    over the ranks: ``func(member_model) -> (ticks, signal)`` is evaluated for
    this rank's block of members (same member numbering as the reference's
    serial loop, decorators.py:55-60) and the partial sums are combined with one
    reduce.  Used for disorder-averaged third-order / 2D responses."""
    import torch
    rank, size = world()
    first, count = shard_members(ensemble_size, rank, size)
    ticks, total = None, None
    for n in range(first, first + count):
        ticks, signal = func(dynamical_model.sample(n, random_orientations))
        total = signal if total is None else total + signal
    if total is None:           # more ranks than members: contribute zeros
        ticks, probe = func(dynamical_model.sample(0, random_orientations))
        total = np.zeros_like(probe)
    part = torch.from_numpy(np.ascontiguousarray(total))
    if torch.cuda.is_available() and size > 1:
        part = part.cuda()
    total = reduce_sum(part, dst)
    return ticks, (total / ensemble_size).cpu().numpy()


def third_order_response_sharded(dynamical_model, coherence_time_max,
                                 ensemble_size, population_time_max=None,
                                 population_times=None, geometry='-++',
                                 polarization='xxxx', include_signal=None,
                                 ensemble_random_orientations=False,
                                 exact_isotropic_average=False, dst=None,
                                 **integrate_kwargs):
    """``third_order_response`` (reference response.py:340-427) for a disorder
    ensemble sharded over the GPUs of the current torch.distributed job."""
    from .simulate.response import (third_order_response, _batchable,
                                    _third_order_response_batched)
    if _batchable(dynamical_model):
        # each rank propagates its block of members as one device batch
        import torch
        rank, size = world()
        first, count = shard_members(ensemble_size, rank, size)
        ticks, part = _third_order_response_batched(
            dynamical_model, coherence_time_max, population_time_max,
            population_times, geometry, polarization, include_signal,
            max(count, 1), ensemble_random_orientations, first, False,
            exact_isotropic_average=exact_isotropic_average, **integrate_kwargs)
        if count == 0:
            part = torch.zeros_like(part)
        total = reduce_sum(part, dst)
        from . import _capi
        return ticks, _capi.to_host(total / ensemble_size)

    def one(member):
        return third_order_response(
            member, coherence_time_max, population_time_max, population_times,
            geometry, polarization, include_signal,
            exact_isotropic_average=exact_isotropic_average, **integrate_kwargs)

    return ensemble_signal_sharded(one, dynamical_model, ensemble_size,
                                   ensemble_random_orientations, dst)


def two_dimensional_spectra_sharded(dynamical_model, coherence_time_max,
                                    ensemble_size, population_time_max=None,
                                    population_times=None, geometry='-++',
                                    polarization='xxxx', include_signal=None,
                                    ensemble_random_orientations=False,
                                    exact_isotropic_average=False, dst=None,
                                    **integrate_kwargs):
    """``two_dimensional_spectra`` (reference response.py:430-455) for a disorder
    ensemble of a dense-generator model: every rank propagates its block of
    members as one device batch, the partial third-order signals meet in ONE
    NCCL reduce, and the two Fourier transforms (kernel K7) run on the reduced
    signal."""
    import torch
    from .simulate.response import _batchable, _third_order_response_batched
    from .simulate.utils import fourier_transform
    if not _batchable(dynamical_model):
        raise NotImplementedError('sharded 2D spectra need a dense-generator model')
    rank, size = world()
    first, count = shard_members(ensemble_size, rank, size)
    (t1, t2, t3), part = _third_order_response_batched(
        dynamical_model, coherence_time_max, population_time_max,
        population_times, geometry, polarization, include_signal,
        max(count, 1), ensemble_random_orientations, first, False,
        exact_isotropic_average=exact_isotropic_average, **integrate_kwargs)
    if count == 0:
        part = torch.zeros_like(part)
    X = reduce_sum(part, dst) / ensemble_size
    rw_freq, unit_convert = dynamical_model.rw_freq, dynamical_model.unit_convert
    f1, X = fourier_transform(t1, X, 0, rw_freq=rw_freq, sign=-1, unit_convert=unit_convert)
    f3, X = fourier_transform(t3, X, 2, rw_freq=rw_freq, unit_convert=unit_convert)
    from . import _capi
    return (f1, t2, f3), (_capi.to_host(X) if not isinstance(X, np.ndarray) else X)
