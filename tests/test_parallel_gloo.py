"""
N > 1 host logic on CPU: world_size-2 ``gloo`` job that shards a disorder
ensemble over ranks exactly as the GPU path does (member blocks + one reduce),
with the CPU oracle standing in for the device propagation.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
import qspectra_b200 as qb
from qspectra_b200 import parallel, systems


def test_shard_members_tiles_the_ensemble():
    for E in (0, 1, 7, 10, 10000):
        for world in (1, 2, 3, 8):
            blocks = [parallel.shard_members(E, r, world) for r in range(world)]
            covered = [n for first, count in blocks for n in range(first, first + count)]
            assert covered == list(range(E))
            counts = [c for _, c in blocks]
            assert max(counts) - min(counts) <= 1
    with pytest.raises(ValueError):
        parallel.shard_members(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, E, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    model = oracle.OracleRedfield(systems.fmo(n_sites=3), hilbert_subspace='e',
                                  unit_convert=qb.CM_FS, secular=False)
    t = np.arange(0, 60, model.time_step)

    def partial(first, count):
        total = np.zeros((len(t), 9), dtype=complex)
        base = model.hamiltonian
        for n in range(first, first + count):
            member = oracle.OracleRedfield.__new__(oracle.OracleRedfield)
            member.__dict__.update(model.__dict__)
            member.hamiltonian = base.sample(n)
            y0 = member.density_matrix_to_state_vector(
                np.diag([1., 0, 0]).astype(complex), 'ee')
            total += oracle.integrate(member.equation_of_motion('ee'), y0, t,
                                      **oracle.TIGHT)
        return total

    assert parallel.world() == (rank, world)
    mean = parallel.sharded_ensemble_mean(partial, E)
    # complex tensors go through the real view; all ranks hold the same mean
    gathered = [torch.zeros_like(torch.view_as_real(mean)) for _ in range(world)]
    dist.all_gather(gathered, torch.view_as_real(mean).contiguous())
    assert all(torch.equal(g, gathered[0]) for g in gathered)
    if rank == 0:
        np.save(out, mean.numpy())
    dist.destroy_process_group()


def test_sharded_ensemble_matches_serial(tmp_path):
    E, world = 5, 2
    out = str(tmp_path / 'mean.npy')
    mp.spawn(_worker, args=(world, _free_port(), E, out), nprocs=world, join=True)
    got = np.load(out)
    model = oracle.OracleRedfield(systems.fmo(n_sites=3), hilbert_subspace='e',
                                  unit_convert=qb.CM_FS, secular=False)
    _, ref = oracle.ensemble_average(
        lambda m: (None, oracle.integrate(
            m.equation_of_motion('ee'),
            m.density_matrix_to_state_vector(np.diag([1., 0, 0]).astype(complex), 'ee'),
            np.arange(0, 60, model.time_step), **oracle.TIGHT)), model, E)
    assert np.linalg.norm(got - ref) / np.linalg.norm(ref) < 1e-13
