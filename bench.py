#!/usr/bin/env python
"""
Benchmark of the Liouville-space propagation hot path (BASELINE.json metric:
"HEOM/Redfield state-steps/sec ... % HBM/FP64 roofline").

    python bench.py --gpus N --steps K --warmup W            # this engine
    python bench.py --impl reference --gpus N --steps K ...  # CPU reference arm

Workload (BASELINE.json configs[1]): FMO 7-site non-secular Redfield population
dynamics in the 49-dimensional 'ee' Liouville subspace, 1e4 static-disorder
realisations (FWHM 100 cm^-1, reference sampler, seed 0), 1 ps on the model's
own 197-point output grid.  One "step" = the whole ensemble propagated over the
whole grid.  With N GPUs every rank propagates its own 1e4 members (weak
scaling) and the ensemble mean is combined with one NCCL reduce.

Units: a *state-step* is one ensemble member advanced by one accepted
integrator step = one output interval here (the default integrator for a
constant generator on a uniform grid builds exp(L dt) once per member on the
FP64 tensor cores and applies it once per interval; `rhs_per_state_step` counts
the matrix-vector applications of the stepping phase, the propagator build is
reported under `roofline`).  The CPU arms report grid-steps (one member advanced
by one output interval at rtol=1e-10), i.e. the same unit.

Extra sub-objects of the JSON line (same run, same box):
  heom          depth-8 / depth-4 / 512-member depth-4 HEOM legs (HBM roofline of the
                hierarchy kernel), an end-to-end `simulate_dynamics(HEOMModel)` number and
                the reference's own HEOM timed beside it;
  scaling_extra strong scaling of the headline workload (`--members` in total) and the
                disorder x orientation averaged 2D spectrum (BASELINE configs[3]) sharded
                over the N ranks, so that the driver's N = 1, 2, 4, 8 lines carry three
                scaling curves;
  parity        members {0, E/2 - 1, E - 1} of the timed ensemble against the CPU oracle.
"""
import os

# BLAS threads of the CPU arms: one per worker process (set before numpy is imported)
for _v in ('OMP_NUM_THREADS', 'OPENBLAS_NUM_THREADS', 'MKL_NUM_THREADS'):
    os.environ.setdefault(_v, '1')

import argparse
import json
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'redfield_state_steps_per_sec'
UNIT = 'state-steps/s'
DURATION_FS = 1000.0
TIGHT = dict(rtol=1e-10, atol=1e-12, nsteps=100000)


def parse():
    p = argparse.ArgumentParser()
    p.add_argument('--gpus', type=int, default=1)
    p.add_argument('--steps', type=int, default=10)
    p.add_argument('--warmup', type=int, default=3)
    p.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    p.add_argument('--members', type=int, default=10000)
    p.add_argument('--e2e-steps', type=int, default=5)
    p.add_argument('--cpu-members', type=int, default=0,
                   help='members per CPU step (0: 4096 for the reference arm, ~10-20 s of '
                        'work for the cpu_baseline leg)')
    p.add_argument('--no-heom', action='store_true')
    p.add_argument('--no-cpu', action='store_true')
    p.add_argument('--no-extra', action='store_true',
                   help='skip the strong-scaling and 2D-spectra legs')
    return p.parse_args()


def workload_config(members, n_gpus):
    return {'workload': 'FMO 7-site non-secular Redfield population dynamics, '
                        "'ee' subspace (M=49), %d static-disorder members per GPU, "
                        '1 ps / 197 output points' % members,
            'members_per_gpu': members, 'state_dim': 49, 'grid_points': 197,
            'integrator': 'propagator stepping in Hermitian coordinates (populations, Re and Im of '
                          'the coherences), where the generator of a Hermiticity-preserving master '
                          'equation is a real 49 x 49 matrix: change of coordinates, exp(G dt) per '
                          'member on the FP64 tensor cores (|A| <= 1/2 scaling, fixed degree-12 Taylor '
                          'polynomial in Paterson-Stockmeyer form, squarings), then u <- P u per '
                          'output interval; coordinates, propagators and trajectories are rebuilt in '
                          'every timed step (QSX_NO_HERMITIAN_FORM=1: the complex path)',
            'parallelism': 'ensemble members sharded over %d GPU(s), one NCCL '
                           'reduce' % n_gpus,
            'launch': 'the six kernels of a step are replayed as one CUDA graph (engine.CapturedEnsembleStep); '
                      'kernel times for the roofline come from an untimed pass through the eager calls '
                      '(QSX_BENCH_NO_GRAPH=1: eager launches in the timed region too)',
            'l2_policy': 'inputs larger than L2 (1e4 generators = 384 MB vs 126 MB L2)'}


# --------------------------------------------------------------- CPU arms
# Each worker process holds one model built once (the reference's own classes when
# oracle/_ref is present, the oracle port otherwise) and runs the reference's serial
# per-member path -- model.hamiltonian.sample(n) -> Redfield tensor -> ZVODE at rtol=1e-10 --
# for a block of members.
_W = {}


def _worker_init(kind):
    from qspectra_b200 import systems
    if kind == 'reference':
        from oracle import vendor_ref
        q = vendor_ref.load()
        from qspectra.utils import copy_with_new_cache
        model = q.RedfieldModel(systems.fmo(ns=q), hilbert_subspace='e',
                                unit_convert=q.CM_FS, secular=False)
        _W.update(kind=kind, q=q, model=model, copy=copy_with_new_cache)
    else:
        import oracle
        import qspectra_b200 as qb
        model = oracle.OracleRedfield(systems.fmo(), hilbert_subspace='e',
                                      unit_convert=qb.CM_FS, secular=False)
        _W.update(kind=kind, oracle=oracle, model=model)


def _member_trajectory(n, duration):
    """(t, rho[t, 7, 7]) of ensemble member n through the arm's own public API."""
    model = _W['model']
    psi0 = np.eye(7)[0]
    if _W['kind'] == 'reference':
        member = _W['copy'](model)                      # dynamics/base.py:120-128
        member.hamiltonian = model.hamiltonian.sample(n)
        return _W['q'].simulate_dynamics(member, psi0, duration, **TIGHT)
    oracle = _W['oracle']
    member = model.__class__.__new__(model.__class__)
    member.__dict__.update(model.__dict__)
    member.hamiltonian = model.hamiltonian.sample(n)
    return oracle.simulate_dynamics(member, psi0, duration, **oracle.TIGHT)


def _worker_chunk(job):
    lo, hi, duration = job
    total, nt = 0, 0
    for n in range(lo, hi):
        t, rho = _member_trajectory(n, duration)
        total = total + rho
        nt = len(t)
    return total, nt


class CpuArm(object):
    """Persistent pool of worker processes (one BLAS thread each)."""

    def __init__(self, workers=None):
        import multiprocessing as mp
        from oracle import vendor_ref
        self.kind = 'reference' if vendor_ref.load() is not None else 'port'
        self.workers = workers or os.cpu_count() or 1
        self.pool = mp.get_context('fork').Pool(self.workers, _worker_init, (self.kind,))

    def close(self):
        self.pool.close()
        self.pool.join()

    def step(self, n_members, first=0, duration=DURATION_FS):
        """Ensemble of `n_members` members (first .. first + n - 1): wall seconds, grid points"""
        chunks = min(n_members, self.workers * 4)          # small chunks balance the workers
        b = np.linspace(0, n_members, chunks + 1).astype(int)
        jobs = [(first + int(b[i]), first + int(b[i + 1]), duration) for i in range(chunks)
                if b[i + 1] > b[i]]
        t0 = time.perf_counter()
        res = self.pool.map(_worker_chunk, jobs, chunksize=1)
        wall = time.perf_counter() - t0
        return wall, res[0][1]

    def serial_rate(self, n_members=8, duration=DURATION_FS):
        """grid-steps/s of ONE worker (the reference itself is serial)"""
        t0 = time.perf_counter()
        res = self.pool.map(_worker_chunk, [(0, n_members, duration)], chunksize=1)
        wall = time.perf_counter() - t0
        return n_members * (res[0][1] - 1) / wall


def run_reference_arm(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    arm = CpuArm()
    n = args.cpu_members or 4096
    for _ in range(min(args.warmup, 1)):
        arm.step(max(arm.workers * 4, n // 8))
    serial = arm.serial_rate()
    walls = []
    nt = 197
    for _ in range(args.steps):
        wall, nt = arm.step(n)
        walls.append(wall)
    arm.close()
    value = float(args.steps * n * (nt - 1) / np.sum(walls))
    what = ('the unmodified reference (oracle/_ref): RedfieldModel + simulate_dynamics per member'
            if arm.kind == 'reference' else 'oracle port of the reference path')
    line = {'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT,
            'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': 1e3 * float(np.mean(walls)), 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'complex128',
            'data': 'synthetic', 'config': workload_config(args.members, args.gpus),
            'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': arm.workers,
                             'kind': arm.kind,
                             'sample': '%d members x %d output intervals per step (%s: generator '
                                       'rebuild + ZVODE rtol=1e-10 per member), %d worker '
                                       'processes with one BLAS thread each'
                                       % (n, nt - 1, what, arm.workers),
                             'serial_value': serial,
                             'step_spread': float((np.max(walls) - np.min(walls)) / np.mean(walls))},
            'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0,
                    'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line))


def heom_cpu_baseline(n_intervals=10):
    """The reference's own HEOM (config 2: FMO, K = 1, depth 4) on one host core: generator
    build, RHS applications/s and output intervals/s at rtol = 1e-10."""
    from oracle import vendor_ref
    from qspectra_b200 import systems
    q = vendor_ref.load()
    out = {'cores': 1}
    if q is not None:
        t0 = time.perf_counter()
        m = q.HEOMModel(systems.fmo(ns=q), hilbert_subspace='e', unit_convert=q.CM_FS,
                        level_cutoff=4, K=1)
        f = m.equation_of_motion('ee')
        out['kind'] = 'reference'
        y0 = m.density_matrix_to_state_vector(np.diag(np.eye(7)[0]).astype(complex), 'ee')
        integrate = q.simulate.utils.integrate
        dt = m.time_step
    else:
        import oracle
        import qspectra_b200 as qb
        t0 = time.perf_counter()
        m = oracle.OracleHEOM(systems.fmo(), hilbert_subspace='e', unit_convert=qb.CM_FS,
                              level_cutoff=4, K=1)
        f = m.equation_of_motion('ee')
        out['kind'] = 'port'
        y0 = m.density_matrix_to_state_vector(np.diag(np.eye(7)[0]).astype(complex), 'ee')
        integrate = oracle.integrate
        dt = m.time_step
    out['build_s'] = time.perf_counter() - t0
    calls = [0]

    def counted(t, y):
        calls[0] += 1
        return f(t, y)
    t = dt * np.arange(n_intervals + 1)
    t0 = time.perf_counter()
    integrate(counted, y0, t, **TIGHT)
    wall = time.perf_counter() - t0
    out.update(state_steps_per_s=n_intervals / wall, rhs_per_s=calls[0] / wall,
               rhs_per_state_step=calls[0] / n_intervals, wall_s=wall,
               sample='FMO HEOM K=1 level_cutoff=4 (680 ADOs x 49), %d output intervals, '
                      'ZVODE rtol=1e-10, one process' % n_intervals)
    return out


# ------------------------------------------------------------------ clocks
class ClockSampler(object):
    QUERY = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,'
             'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.QUERY,
                 '--format=csv,noheader,nounits', '-lms', '25'],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.perf_counter(), line.strip()))

    def count_between(self, t_a, t_b):
        return sum(1 for ts, _ in list(self.samples) if t_a <= ts <= t_b)

    def stop(self, t_a=None, t_b=None):
        """Summary of the samples that arrived inside [t_a, t_b] (all if not given)."""
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for ts, s in self.samples:
            if t_a is not None and not (t_a <= ts <= t_b):
                continue
            parts = [x.strip() for x in s.split(',')]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, flag in zip(names, parts[3:7]):
                if flag.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': float(np.median(sm)) if sm else None,
                'sm_max_mhz': float(np.max(mx)) if mx else None,
                'samples': len(sm), 'reasons': sorted(reasons)}


# ------------------------------------------------------------------- our arm
def measure_fp64_peak(torch):
    """cuBLAS DGEMM ceiling on this GPU (the FP64 roofline denominator;
    MEASURED_PEAKS.json only holds bf16 and HBM copy numbers)."""
    n = 6144
    a = torch.randn(n, n, dtype=torch.float64, device='cuda')
    b = torch.randn(n, n, dtype=torch.float64, device='cuda')
    torch.matmul(a, b)
    best = 0.0
    for _ in range(4):
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        torch.matmul(a, b)
        e1.record()
        torch.cuda.synchronize()
        best = max(best, 2.0 * n ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    del a, b
    return best


def ncu_traffic(key):
    """dram bytes per launch of a kernel from the committed ncu capture, or None."""
    path = os.path.join(ROOT, 'profiles', 'traffic.json')
    try:
        with open(path) as fh:
            return json.load(fh).get(key)
    except (OSError, ValueError):
        return None


def measured_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as fh:
            return json.load(fh), 'measured (MEASURED_PEAKS.json)'
    return {'hbm_gbs': 6650.0}, 'fallback (B200_PROFILING.md)'


def _heom_roofline(rhs_per_s, D, rhs, key):
    peaks, src = measured_peaks()
    achieved = rhs_per_s * 32.0 * D / 1e9
    per_rhs = ncu_traffic(key)
    return {'bound': 'hbm', 'achieved': achieved, 'peak': peaks['hbm_gbs'], 'unit': 'GB/s',
            'frac': achieved / peaks['hbm_gbs'],
            'traffic': per_rhs * rhs if per_rhs is not None else None,
            'traffic_per_rhs': per_rhs,
            'achieved_is': 'rhs_per_s x 32 D (algorithmic bytes per RHS application)',
            'peak_source': src, 'algorithmic_bytes_per_rhs': 32 * D}


def heom_leg(torch, qb, systems, engine, with_cpu):
    """BASELINE configs[4]/[2]: FMO HEOM K=1 depth 8 (116 280 ADOs x 49) and depth 4, one
    trajectory each, and a 512-member depth-4 disorder ensemble: RHS applications/s and HBM
    roofline of the hierarchy kernel through the default integrator of HEOMModel (product-form
    Taylor propagator, csrc/heom_row.cuh, where the row tile applies)."""
    out = {}
    for depth, n_int in ((8, 12), (4, 40)):
        model = qb.HEOMModel(systems.fmo(), hilbert_subspace='e',
                             unit_convert=qb.CM_FS, level_cutoff=depth, K=1)
        t0 = time.perf_counter()
        eom = model.equation_of_motion('ee')
        build_s = time.perf_counter() - t0
        y0 = model.density_matrix_to_state_vector(
            np.diag(np.eye(7)[0]).astype(complex), 'ee')
        t = model.time_step * np.arange(n_int + 1)
        y0_dev = torch.from_numpy(y0).cuda().reshape(1, -1)
        eom.propagate(y0_dev, t[:2], save=('ado0',), return_device=True)   # warm-up
        best = None
        for _ in range(3):
            eom.propagate(y0_dev, t, save=('ado0',), return_device=True)
            if best is None or eom.last['kernel_ms'] < best['kernel_ms']:
                best = dict(eom.last)
        D = eom.dim
        rhs_per_s = best['rhs'] / (best['kernel_ms'] * 1e-3)
        out['depth%d' % depth] = {
            'workload': 'FMO 7-site HEOM K=1 level_cutoff=%d: %d ADOs x 49, D=%d, '
                        'free evolution, %d output intervals' % (depth, eom.n_ado, D, n_int),
            'integrator': best['method'],
            'rhs_per_s': rhs_per_s,
            'state_steps_per_s': best['steps'] / (best['kernel_ms'] * 1e-3),
            'rhs_per_state_step': best['rhs'] / max(1, best['steps']),
            'kernel_ms': best['kernel_ms'], 'setup_s': build_s,
            'roofline': _heom_roofline(rhs_per_s, D, best['rhs'], 'heom_depth%d_per_rhs' % depth)}
        if depth == 4:
            # end to end through the public API: host model and initial state in, host
            # density matrices out (handle creation included, as in the reference's call)
            qb.simulate_dynamics(model, np.eye(7)[0], n_int * model.time_step)
            h0, d0 = _transfer()
            times = []
            for _ in range(3):
                t0 = time.perf_counter()
                qb.simulate_dynamics(model, np.eye(7)[0], n_int * model.time_step)
                torch.cuda.synchronize()
                times.append(time.perf_counter() - t0)
            h1, d1 = _transfer()
            nt = len(np.arange(0, n_int * model.time_step, model.time_step))
            out['depth4']['e2e'] = {
                'value': (nt - 1) / float(np.mean(times)), 'unit': UNIT,
                'seconds_per_call': float(np.mean(times)),
                'h2d_bytes_per_step': (h1 - h0) // 3, 'd2h_bytes_per_step': (d1 - d0) // 3,
                'path': 'simulate_dynamics(HEOMModel(level_cutoff=4), psi0, %d intervals) from '
                        'host objects' % (nt - 1)}
        del eom, model
    # BASELINE configs[2] batched: the depth-4 hierarchy is L2-resident for one trajectory,
    # so the HBM roofline of the hierarchy kernel is measured on a 512-member disorder
    # ensemble (one column and one Hamiltonian per member, 273 MB per state vector)
    E = 512
    model = qb.HEOMModel(systems.fmo(), hilbert_subspace='e', unit_convert=qb.CM_FS,
                         level_cutoff=4, K=1)
    t0 = time.perf_counter()
    eom = model.ensemble_eom(E, False, 'ee')
    build_s = time.perf_counter() - t0
    y0 = model.density_matrix_to_state_vector(np.diag(np.eye(7)[0]).astype(complex), 'ee')
    y0_dev = torch.from_numpy(y0).cuda().reshape(1, -1).expand(E, -1).contiguous()
    t = model.time_step * np.arange(11)
    best = None
    for _ in range(3):
        eom.propagate(y0_dev, t, save=('ado0',), generators=np.arange(E), return_device=True)
        if best is None or eom.last['kernel_ms'] < best['kernel_ms']:
            best = dict(eom.last)
    rhs_per_s = best['rhs'] / (best['kernel_ms'] * 1e-3)
    out['depth4_ensemble512'] = {
        'workload': 'FMO 7-site HEOM K=1 level_cutoff=4, %d static-disorder members: %d ADOs x 49 '
                    'per member, 10 output intervals' % (E, eom.n_ado),
        'integrator': best['method'],
        'rhs_per_s': rhs_per_s,
        'state_steps_per_s': best['steps'] / (best['kernel_ms'] * 1e-3),
        'rhs_per_state_step': best['rhs'] / max(1, best['steps']),
        'kernel_ms': best['kernel_ms'], 'setup_s': build_s,
        'roofline': _heom_roofline(rhs_per_s, eom.dim, best['rhs'], 'heom_depth4_ens512_per_rhs')}
    del eom, model
    # BASELINE configs[4], second half: vibronic dimer with explicit modes (2 sites x 2 modes x 2
    # levels: 8 'e' states, M = 64; row tile Cfg<8, 2, 4>: states grouped per site)
    model = qb.HEOMModel(systems.jonas_dimer(), hilbert_subspace='e', unit_convert=qb.CM_FS,
                         level_cutoff=10, K=1)
    t0 = time.perf_counter()
    eom = model.equation_of_motion('ee')
    build_s = time.perf_counter() - t0
    psi = np.zeros(eom.M, dtype=complex)
    psi[0] = 1.0
    y0_dev = torch.from_numpy(model._pad(psi)).cuda().reshape(1, -1)
    t = model.time_step * np.arange(21)
    best = None
    for _ in range(3):
        eom.propagate(y0_dev, t, save=('ado0',), return_device=True)
        if best is None or eom.last['kernel_ms'] < best['kernel_ms']:
            best = dict(eom.last)
    rhs_per_s = best['rhs'] / (best['kernel_ms'] * 1e-3)
    out['vibronic_dimer'] = {
        'workload': 'vibronic (Jonas) dimer HEOM, 2 explicit modes x 2 levels: 8 states in e, '
                    'M = %d, K=1 level_cutoff=10: %d ADOs, D=%d, 20 output intervals'
                    % (eom.M, eom.n_ado, eom.dim),
        'integrator': best['method'], 'rhs_per_s': rhs_per_s,
        'state_steps_per_s': best['steps'] / (best['kernel_ms'] * 1e-3),
        'rhs_per_state_step': best['rhs'] / max(1, best['steps']),
        'kernel_ms': best['kernel_ms'], 'setup_s': build_s,
        'roofline': _heom_roofline(rhs_per_s, eom.dim, best['rhs'], 'heom_vibronic_per_rhs')}
    # the same hierarchy as a 64-column batch (the throughput case: response-function columns,
    # polarisation configurations) on the shaped row tile Cfg<8, 2, 4>
    yb = y0_dev.expand(64, -1).contiguous()
    best = None
    for _ in range(3):
        eom.propagate(yb, t, save=('ado0',), return_device=True)
        if best is None or eom.last['kernel_ms'] < best['kernel_ms']:
            best = dict(eom.last)
    rhs_per_s = best['rhs'] / (best['kernel_ms'] * 1e-3)
    out['vibronic_dimer_batch64'] = {
        'workload': 'the same vibronic-dimer hierarchy, 64 columns in one launch, 20 output intervals',
        'integrator': best['method'], 'rhs_per_s': rhs_per_s,
        'state_steps_per_s': best['steps'] / (best['kernel_ms'] * 1e-3),
        'kernel_ms': best['kernel_ms'],
        'roofline': _heom_roofline(rhs_per_s, eom.dim, best['rhs'], 'heom_vibronic_batch64_per_rhs')}
    del eom, model
    if with_cpu:
        out['cpu_baseline'] = heom_cpu_baseline()
    return out


def zofe_leg(torch, qb, systems, fp64_peak):
    """K3: ZOFE master equation, FMO 'e' with the 16-pseudomode bath, a 592-member disorder
    ensemble (one CTA per trajectory, two per SM: two full waves): RHS applications/s against the FP64 ceiling."""
    E = 592
    model = qb.ZOFEModel(systems.fmo(bath='pseudomode'), hilbert_subspace='e',
                         unit_convert=qb.CM_FS)
    eom = model.ensemble_eom(E, False, 'ee')
    y0 = model.density_matrix_to_state_vector(np.diag(np.eye(7)[0]).astype(complex), 'ee')
    y0_dev = torch.from_numpy(y0).cuda().reshape(1, -1).expand(E, -1).contiguous()
    t = model.time_step * np.arange(21)
    best = None
    for _ in range(3):
        eom.propagate(y0_dev, t, save=('ado0',), generators=np.arange(E), return_device=True)
        if best is None or eom.last['kernel_ms'] < best['kernel_ms']:
            best = dict(eom.last)
    n, P, S = 7, 16, 7
    flops = 8.0 * n ** 3 * (2 * P * S + 5 * S + 2)
    rhs_per_s = best['rhs'] / (best['kernel_ms'] * 1e-3)
    tf = rhs_per_s * flops / 1e12
    return {'workload': 'ZOFE FMO e (7 states, 16 pseudomodes x 7 sites), %d disorder members, '
                        '20 output intervals, DOPRI5 rtol=1e-10' % E,
            'integrator': best['method'], 'rhs_per_s': rhs_per_s,
            'state_steps_per_s': best['steps'] / (best['kernel_ms'] * 1e-3),
            'grid_steps_per_s': E * 20 / (best['kernel_ms'] * 1e-3),
            'kernel_ms': best['kernel_ms'],
            'roofline': {'bound': 'tensor', 'achieved': tf, 'peak': fp64_peak, 'unit': 'TFLOP/s',
                         'frac': tf / fp64_peak, 'traffic': None,
                         'algorithmic_flops_per_rhs': flops,
                         'peak_source': 'cuBLAS FP64 GEMM measured in this run'}}


def _transfer():
    from qspectra_b200 import _capi
    return _capi.transfer_bytes()


def parity_check(qb, systems, eom, y0, t, members):
    """Selected members of the timed ensemble, propagated by the same device handle and
    integrator as the timed step, against the CPU oracle (checker only)."""
    import oracle
    from qspectra_b200 import _capi
    sel = np.asarray(members)
    y0_dev = _capi.to_device(y0).reshape(1, -1).expand(len(sel), -1).contiguous()
    got = _capi.to_host(eom.propagate(y0_dev, t, generators=sel, return_device=True))
    om = oracle.OracleRedfield(systems.fmo(), hilbert_subspace='e', unit_convert=qb.CM_FS,
                               secular=False)
    worst = 0.0
    for i, n in enumerate(sel):
        member = om.__class__.__new__(om.__class__)
        member.__dict__.update(om.__dict__)
        member.hamiltonian = om.hamiltonian.sample(int(n))
        f = member.equation_of_motion('ee')
        ref = oracle.integrate(f, y0, t, **oracle.TIGHT)
        worst = max(worst, float(np.linalg.norm(got[i] - ref) / np.linalg.norm(ref)))
    return {'members': [int(n) for n in sel], 'rel_l2_max': worst, 'tolerance': 1e-8,
            'checker': 'oracle (generator + ZVODE rtol=1e-10) on the same members'}


def spectra2d_leg(torch, dist, qb, systems, parallel, world, rank, members_per_gpu=32):
    """BASELINE configs[3]: FMO third-order response / 2D spectrum ('gef', t1 = t3 = 197 points,
    5 population times), exact isotropic average (21 polarisation configurations) x static
    disorder, members sharded over the ranks, one NCCL reduce, device Fourier transforms."""
    model = qb.RedfieldModel(systems.fmo(), hilbert_subspace='gef', unit_convert=qb.CM_FS,
                             secular=False)
    E = members_per_gpu * world
    kw = dict(population_times=np.linspace(0, 1000, 5), geometry='-++', polarization='xxxx',
              exact_isotropic_average=True, dst=0)

    def once():
        return parallel.two_dimensional_spectra_sharded(model, 1000, E, **kw)
    once()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    times = []
    for _ in range(2):
        t0 = time.perf_counter()
        (f1, t2, f3), X = once()
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
    sec = float(np.mean(times))
    if world > 1:
        tmax = torch.tensor([sec], dtype=torch.float64, device='cuda')
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        sec = float(tmax.item())
    n_t = 197
    # propagated columns per member and configuration, three pathways: t1 stage 1, t2 stage n_t1,
    # t3 stage 1; each advanced over its grid
    units = E * 21
    return {'workload': 'FMO 2D spectrum (gef, 197 x 5 x 197, -++, xxxx, exact isotropic average: '
                        '21 configurations) x %d disorder members (%d per GPU)' % (E, members_per_gpu),
            'seconds_per_spectrum': sec,
            'member_configurations_per_s': units / sec,
            'scaling': 'weak', 'finite': bool(np.isfinite(X).all()) if rank == 0 else None}


def run_ours(args):
    import torch
    import torch.distributed as dist
    import qspectra_b200 as qb
    from qspectra_b200 import systems, engine, _capi, parallel

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    if world > 1:
        import datetime
        dist.init_process_group('nccl', device_id=torch.device('cuda', local),
                                timeout=datetime.timedelta(seconds=180))
    E = args.members
    model = qb.RedfieldModel(systems.fmo(), hilbert_subspace='e',
                             unit_convert=qb.CM_FS, secular=False)
    t = np.arange(0, DURATION_FS, model.time_step)
    psi0 = np.eye(7)[0]
    y0 = model.density_matrix_to_state_vector(np.outer(psi0, psi0).astype(complex), 'ee')

    def reduce_across(mean_dev):
        if world > 1:
            buf = torch.view_as_real(mean_dev)
            dist.reduce(buf, dst=0, op=dist.ReduceOp.SUM)
        return mean_dev

    def timed_resident(n_members, member0, total_members, steps, warmup):
        """device-resident leg: generators already in HBM; returns (ms per step, stats, result)"""
        eom = model.ensemble_eom(n_members, False, 'ee', member0=member0)
        y0_dev = _capi.to_device(y0).reshape(1, -1).expand(n_members, -1).contiguous()
        gens = np.arange(n_members)

        def local_step():
            eom.__dict__.pop('_propagators', None)    # rebuild exp(L dt) every step
            # the initial state is a density matrix (Hermitian): dense generators step it in real
            # coordinates and the member sum runs over the real rows (what simulate_dynamics does
            # after checking the host state)
            out = eom.propagate(y0_dev, t, generators=gens, return_device=True,
                                hermitian_state=True, packed=True)
            return engine.reduce_members(out, 1.0 / total_members)

        # The same step as one CUDA graph (engine.CapturedEnsembleStep: change of coordinates + real
        # propagators rebuilt by every replay, packing, stepping, member mean, conversion of the mean);
        # the eager form above is kept for the instrumented pass that times the kernels one by one.
        captured = None
        if not os.environ.get('QSX_BENCH_NO_GRAPH'):
            try:
                captured = engine.CapturedEnsembleStep(eom, y0_dev, t, 1.0 / total_members)
                ref_mean = local_step()
                if float((captured.run() - ref_mean).abs().max()) > 1e-14:
                    raise RuntimeError('captured step differs from the eager step')
                captured.verify()
            except Exception as exc:                 # keep the bench alive on the eager path
                sys.stderr.write('bench: CUDA-graph step unavailable (%r), eager launches\n' % (exc,))
                captured = None

        def step():
            return reduce_across(captured.run() if captured is not None else local_step())
        for _ in range(warmup):
            step()
        # untimed rehearsal of the timed block with the same object lifetimes (the propagators of all
        # K steps stay alive until their device times are collected): the caching allocator grows
        # here, not inside the timed region (cudaMalloc synchronises the device)
        engine.PropagationStats.keep_alive = True
        for _ in range(steps):
            step()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        engine.PropagationStats.reset()
        engine.PropagationStats.keep_alive = True      # every timed step's device times are collected
        l0 = _capi.kernel_launches()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        torch.cuda.synchronize()
        t_a = time.perf_counter()
        e0.record()
        for _ in range(steps):
            result = step()
        e1.record()
        torch.cuda.synchronize()
        t_b = time.perf_counter()
        if world > 1:
            dist.barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            tmax = torch.tensor([ms], dtype=torch.float64, device='cuda')
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            ms = float(tmax.item())
        st = engine.PropagationStats
        launches = _capi.kernel_launches() - l0
        if captured is not None:
            # graph replays carry no per-kernel events: the kernels are timed one by one in an
            # untimed pass of the same step through the eager calls
            launches = captured.launches_per_run * steps
            captured.verify()
            st.reset()
            st.keep_alive = True
            for _ in range(steps):
                local_step()
            torch.cuda.synchronize()
        st.flush()          # collect the deferred device times of the (instrumented) steps
        st.keep_alive = False
        stats = dict(expm_ms=st.expm_ms, expm_gemms=st.expm_gemms, expm_builds=st.expm_builds,
                     form_ms=st.form_ms, hermitian_builds=st.hermitian_builds,
                     kernel_ms=st.kernel_ms / max(1, st.propagations),
                     rhs=st.rhs_evaluations / max(1, st.propagations),
                     steps=st.accepted_steps / max(1, st.propagations),
                     launches=launches, window=(t_a, t_b), graph=captured is not None)
        rank_local = (lambda: captured.run()) if captured is not None else local_step
        return ms / steps, stats, result, eom, rank_local

    # nvidia-smi needs ~0.1-0.2 s before its first sample: start it ahead of the warm-up and keep
    # only the samples that arrive inside the timed window
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_per_step, stats, result, eom, step = timed_resident(E, rank * E, E * world,
                                                           args.steps, args.warmup)
    t_a, t_b = stats['window']
    clocks_window = 'timed region'
    if rank == 0 and sampler.proc is not None and sampler.count_between(t_a, t_b) < 3:
        # the timed region is shorter than a few sampling periods: keep the identical load
        # running (untimed) until the sampler has seen it at least a few times
        t_end = time.perf_counter() + 0.6
        while time.perf_counter() < t_end:
            step()                      # rank-local part of the step: no collective here
            torch.cuda.synchronize()
        t_b = time.perf_counter()
        clocks_window = ('timed region + 0.6 s of the identical step repeated untimed (timed '
                         'region shorter than 3 sampling periods)')
    if world > 1:
        dist.barrier()
    clocks = sampler.stop(t_a, t_b) if rank == 0 else None
    if clocks is not None:
        clocks['window'] = clocks_window
    value = world * stats['steps'] / (ms_per_step * 1e-3)
    trace_err, parity = None, None
    if rank == 0:
        rho_dev = _capi.to_host(result).reshape(len(t), 7, 7)
        trace_err = float(np.abs(np.einsum('tii->t', rho_dev) - 1).max())
        parity = parity_check(qb, systems, eom, y0, t, [0, E // 2 - 1, E - 1])
    del eom, step
    torch.cuda.empty_cache()

    # ---- end-to-end leg through the public API, host buffers ------------------
    def e2e_once():
        # the call a user makes: host arrays in, host arrays out
        if world == 1:
            _, rho = qb.simulate_dynamics(model, psi0, DURATION_FS, ensemble_size=E)
        else:
            # one process per GPU: contiguous member blocks, one NCCL reduce to rank 0
            _, rho = parallel.simulate_dynamics_sharded(
                model, psi0, DURATION_FS, ensemble_size=E * world, dst=0)
        return rho

    e2e_once()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e2e_times = []
    h0, d0 = _capi.transfer_bytes()
    for _ in range(args.e2e_steps):
        t0 = time.perf_counter()
        e2e_once()
        torch.cuda.synchronize()
        e2e_times.append(time.perf_counter() - t0)
    h1, d1 = _capi.transfer_bytes()
    if world > 1:
        dist.barrier()
    e2e_s = float(np.mean(e2e_times))
    if world > 1:
        tmax = torch.tensor([e2e_s], dtype=torch.float64, device='cuda')
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        e2e_s = float(tmax.item())
    e2e_value = world * E * (len(t) - 1) / e2e_s
    # bytes that crossed PCIe per call on this rank, counted where the copies are issued
    # (library: qsx_transfer_bytes; Python side: _capi.to_device / to_host).  The members'
    # disorder is replayed from the seed ON THE DEVICE, so no per-member array is uploaded.
    h2d = (h1 - h0) // max(1, args.e2e_steps)
    d2h = (d1 - d0) // max(1, args.e2e_steps)

    # ---- further scaling curves (same run): strong scaling, 2D spectra ---------
    extra = None
    if not args.no_extra:
        extra = {}
        first, count = parallel.shard_members(E, rank, world)
        ms_s, st_s, _, eom_s, step_s = timed_resident(count, first, E, max(3, args.steps // 2), 2)
        del eom_s, step_s
        torch.cuda.empty_cache()
        extra['strong'] = {
            'workload': 'the headline workload with %d members IN TOTAL split over the ranks' % E,
            'value': E * (len(t) - 1) / (ms_s * 1e-3), 'unit': UNIT, 'ms_per_step': ms_s,
            'members_per_gpu': count, 'scaling': 'strong'}
        try:
            extra['spectra2d'] = spectra2d_leg(torch, dist, qb, systems, parallel, world, rank)
        except Exception as exc:       # keep the headline line if this leg cannot run
            extra['spectra2d'] = {'error': repr(exc)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks, peak_src = measured_peaks()
    fp64_peak = measure_fp64_peak(torch)
    # kernels of the step: change of coordinates (Hermitian form), tensor-core propagator build,
    # stepping, member sum.  Algorithmic flops: in Hermitian coordinates the generators are REAL,
    # so a product is 2 M^3 and a step 2 M^2 flops (M = 49, padding not counted); on the complex
    # path (QSX_NO_HERMITIAN_FORM=1) 8 M^3 and 8 M^2.
    real_form = stats['hermitian_builds'] > 0
    builds = max(1, stats['expm_builds'])
    form_ms = stats['form_ms'] / builds
    expm_ms = stats['expm_ms'] / builds - form_ms
    flops_per_launch = (2.0 if real_form else 8.0) * 49 ** 3 * stats['expm_gemms'] / builds
    achieved_tf = flops_per_launch / (expm_ms * 1e-3) / 1e12 if expm_ms > 0 else 0.0
    kernel_ms, rhs_per_launch = stats['kernel_ms'], stats['rhs']
    map_flops = (2.0 if real_form else 8.0) * 49 * 49 * rhs_per_launch
    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_per_step,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'float64 (Hermitian coordinates of the complex128 density matrices)' if real_form else 'complex128',
        'data': 'synthetic',
        'config': workload_config(E, world),
        'rhs_per_state_step': rhs_per_launch / max(1.0, stats['steps']),
        'rhs_per_s': world * rhs_per_launch / (ms_per_step * 1e-3),
        'grid_steps_per_s': world * E * (len(t) - 1) / (ms_per_step * 1e-3),
        'trace_error': trace_err,
        'parity': parity,
        'clocks': clocks,
        'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': int(h2d),
                'd2h_bytes_per_step': int(d2h), 'seconds_per_step': e2e_s,
                'seconds_each_step': [round(x, 5) for x in e2e_times],
                'bytes_are': 'counted copies of rank 0 (qsx_transfer_bytes + _capi.to_device/to_host)',
                'path': 'simulate_dynamics(model, psi0, duration, ensemble_size) from host objects: '
                        'seed -> device replay of the seeded disorder streams -> K5 (Jacobi '
                        'eigensystems + Redfield generators) -> Hermitian coordinates -> K1\' real '
                        'propagators -> K1/K4 stepping -> K6 mean over the real rows -> complex '
                        'density matrices -> D2H'},
        'gpu_launches': int(stats['launches']),
        'roofline': None,
    }
    # the step is two kernels of about equal duration: the DMMA propagator build and the
    # propagator stepping on the FP64 pipe; `roofline` describes whichever took longer in
    # THIS run, the other is listed beside it.  Both are measured against the same FP64
    # ceiling (on B200 the DGEMM/DMMA ceiling equals the FP64 FMA ceiling).
    peak_note = ('cuBLAS FP64 GEMM 6144^3 measured in this run (cutlass d884 DMMA kernel; '
                 'FP64 is not in MEASURED_PEAKS.json)')
    map_tf = map_flops / (kernel_ms * 1e-3) / 1e12
    expm_name = ('real_expm3_kernel<7,13,fused> (change of coordinates L -> G in shared memory, then FP64 DMMA m8n8k4 on '
                 'the real generators: exp(G dt) per member, Paterson-Stockmeyer degree 12 + squarings)' if real_form else
                 'dense_expm2_kernel<7,13> (FP64 DMMA m8n8k4, three real products per complex one: exp(L dt) per member, '
                 'Paterson-Stockmeyer degree 12 + squarings)')
    map_name = ('real_map_rows_kernel<25> (u <- P u stepping in Hermitian coordinates, one warp per member, two whole '
                'rows of P per lane in registers; DFMA on the FP64 pipe, measured against the same FP64 ceiling)' if real_form else
                'dense_map_split_kernel<5,10,25,2> (y <- P y stepping with P in registers; DFMA on '
                'the FP64 pipe, measured against the same FP64 ceiling as SURVEY 8d asks '
                'for the dense L.Y contraction)')
    k_expm = {'bound': 'tensor', 'achieved': achieved_tf, 'peak': fp64_peak,
              'unit': 'TFLOP/s', 'frac': achieved_tf / fp64_peak,
              'traffic': (ncu_traffic('real_expm_per_member' if real_form else 'dense_expm_per_member') or 0) * E or None,
              'kernel': expm_name,
              'kernel_ms': expm_ms, 'share_of_step': expm_ms / ms_per_step,
              'algorithmic_flops_per_launch': flops_per_launch, 'peak_source': peak_note}
    k_map = {'bound': 'tensor', 'achieved': map_tf, 'peak': fp64_peak,
             'unit': 'TFLOP/s', 'frac': map_tf / fp64_peak,
             'traffic': (ncu_traffic('real_map_per_member' if real_form else 'dense_map_per_member') or 0) * E or None,
             'kernel': map_name,
             'kernel_ms': kernel_ms, 'share_of_step': kernel_ms / ms_per_step,
             'algorithmic_flops_per_launch': map_flops, 'peak_source': peak_note}
    first_k, second_k = (k_map, k_expm) if kernel_ms >= expm_ms else (k_expm, k_map)
    line['roofline'] = first_k
    line['other_kernels'] = {second_k['kernel']: second_k,
                             'hermitian_form_kernel_ms': (form_ms if form_ms > 0.02 else 'fused into the propagator kernel') if real_form else None,
                             'share_note': 'per step: propagator build + stepping + member '
                                           'reduction; see profiles/ for the ncu launch list'}
    if extra is not None:
        line['scaling_extra'] = extra
    if world == 1 and not args.no_heom:
        line['heom'] = heom_leg(torch, qb, systems, engine, not args.no_cpu)
        try:
            line['zofe'] = zofe_leg(torch, qb, systems, fp64_peak)
        except Exception as exc:
            line['zofe'] = {'error': repr(exc)}
    if world == 1 and not args.no_cpu:
        arm = CpuArm()
        cores = arm.workers
        arm.step(cores * 2)                                  # warm the workers
        serial = arm.serial_rate()
        # about 15 s of all-core work, at most the full ensemble
        n = args.cpu_members or int(min(E, max(16 * cores, 15.0 * cores * serial / 196.0)))
        wall, nt = arm.step(n)
        arm.close()
        line['cpu_baseline'] = {
            'value': n * (nt - 1) / wall, 'unit': UNIT, 'cores': cores, 'kind': arm.kind,
            'serial_value': serial,
            'sample': '%d of %d members x %d output intervals (%s: generator rebuild + ZVODE '
                      'rtol=1e-10 per member), %.1f s on %d worker processes, one BLAS thread each'
                      % (n, E, nt - 1, 'unmodified reference from oracle/_ref'
                         if arm.kind == 'reference' else 'oracle port', wall, cores)}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == 'reference':
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
