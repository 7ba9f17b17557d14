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
"""
import argparse
import json
import os
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


def parse():
    p = argparse.ArgumentParser()
    p.add_argument('--gpus', type=int, default=1)
    p.add_argument('--steps', type=int, default=10)
    p.add_argument('--warmup', type=int, default=3)
    p.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    p.add_argument('--members', type=int, default=10000)
    p.add_argument('--e2e-steps', type=int, default=5)
    p.add_argument('--cpu-members', type=int, default=0,
                   help='members per CPU step (0: sized for ~10-20 s)')
    p.add_argument('--no-heom', action='store_true')
    p.add_argument('--no-cpu', action='store_true')
    return p.parse_args()


def workload_config(members, n_gpus):
    return {'workload': 'FMO 7-site non-secular Redfield population dynamics, '
                        "'ee' subspace (M=49), %d static-disorder members per GPU, "
                        '1 ps / 197 output points' % members,
            'members_per_gpu': members, 'state_dim': 49, 'grid_points': 197,
            'integrator': 'propagator stepping: exp(L dt) per member on the FP64 tensor '
                          'cores (scaled Taylor series + squarings, on-device truncation '
                          'control at 1e-17), then y <- P y per output interval; the '
                          'propagators are rebuilt in every timed step',
            'parallelism': 'ensemble members sharded over %d GPU(s), one NCCL '
                           'reduce' % n_gpus,
            'l2_policy': 'inputs larger than L2 (1e4 generators = 384 MB vs 126 MB L2)'}


# --------------------------------------------------------------- CPU baseline
def _cpu_member_chunk(args):
    lo, hi, duration = args
    import oracle
    import qspectra_b200 as qb
    from qspectra_b200 import systems
    model = oracle.OracleRedfield(systems.fmo(), hilbert_subspace='e',
                                  unit_convert=qb.CM_FS, secular=False)
    total, rhs = 0, 0
    for n in range(lo, hi):
        member = model.__class__.__new__(model.__class__)
        member.__dict__.update(model.__dict__)
        member.hamiltonian = model.hamiltonian.sample(n)
        f = oracle.counting(member.equation_of_motion('ee'))
        y0 = member.density_matrix_to_state_vector(
            np.diag(np.eye(7)[0]).astype(complex), 'ee')
        t = np.arange(0, duration, member.time_step)
        states = oracle.integrate(f, y0, t, **oracle.TIGHT)
        total = total + states
        rhs += f.calls
    return total, rhs, len(t)


def cpu_reference_rate(n_members, duration=DURATION_FS, workers=None):
    """grid-steps/s of the CPU oracle (port of the reference's scipy path:
    generator rebuild per member + ZVODE at rtol=1e-10), members spread over
    all host cores with multiprocessing (the reference itself is serial)."""
    import multiprocessing as mp
    workers = workers or os.cpu_count() or 1
    workers = max(1, min(workers, n_members))
    bounds = np.linspace(0, n_members, workers + 1).astype(int)
    jobs = [(int(bounds[i]), int(bounds[i + 1]), duration) for i in range(workers)
            if bounds[i + 1] > bounds[i]]
    ctx = mp.get_context('fork')
    t0 = time.perf_counter()
    with ctx.Pool(len(jobs)) as pool:
        results = pool.map(_cpu_member_chunk, jobs)
    wall = time.perf_counter() - t0
    nt = results[0][2]
    rhs = sum(r[1] for r in results)
    return {'grid_steps_per_s': n_members * (nt - 1) / wall, 'wall_s': wall,
            'rhs_per_s': rhs / wall, 'workers': len(jobs), 'members': n_members,
            'grid_points': nt}


def run_reference_arm(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    os.environ.setdefault('OMP_NUM_THREADS', '1')
    cores = os.cpu_count() or 1
    n = args.cpu_members or max(cores, 16 * cores)
    for _ in range(max(args.warmup, 0) and 1):
        cpu_reference_rate(max(cores, n // 4))
    rates, walls = [], []
    for _ in range(args.steps):
        r = cpu_reference_rate(n)
        rates.append(r['grid_steps_per_s'])
        walls.append(r['wall_s'])
    value = float(np.sum([n * 196 for _ in rates]) / np.sum(walls))
    line = {'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT,
            'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': 1e3 * float(np.mean(walls)), 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'complex128',
            'data': 'synthetic', 'config': workload_config(args.members, args.gpus),
            'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': r['workers'],
                             'kind': 'port',
                             'sample': '%d members x 196 output intervals per step '
                                       '(generator rebuild + ZVODE rtol=1e-10 per '
                                       'member, as the reference does), %d worker '
                                       'processes' % (n, r['workers'])},
            'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0,
                    'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line))


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


def heom_leg(torch, qb, systems, engine):
    """BASELINE configs[4]/[2]: FMO HEOM K=1 depth 8 (116 280 ADOs x 49), one
    trajectory: RHS applications/s and HBM roofline of the hierarchy kernel."""
    out = {}
    peaks, src = measured_peaks()
    for depth, n_int in ((8, 2), (4, 20)):
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
        achieved = rhs_per_s * 32.0 * D / 1e9
        out['depth%d' % depth] = {
            'workload': 'FMO 7-site HEOM K=1 level_cutoff=%d: %d ADOs x 49, D=%d, '
                        'free evolution, %d output intervals' % (depth, eom.n_ado, D, n_int),
            'rhs_per_s': rhs_per_s,
            'state_steps_per_s': best['steps'] / (best['kernel_ms'] * 1e-3),
            'rhs_per_state_step': best['rhs'] / max(1, best['steps']),
            'kernel_ms': best['kernel_ms'], 'setup_s': build_s,
            'roofline': {'bound': 'hbm', 'achieved': achieved,
                         'peak': peaks['hbm_gbs'], 'unit': 'GB/s',
                         'frac': achieved / peaks['hbm_gbs'], 'traffic': ((ncu_traffic('heom_depth%d_per_rhs' % depth) or 0) * best['rhs']) or None,
                         'achieved_is': 'rhs_per_s x 32 D (algorithmic bytes per RHS application)',
                         'peak_source': src,
                         'algorithmic_bytes_per_rhs': 32 * D}}
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
    achieved = rhs_per_s * 32.0 * eom.dim / 1e9
    out['depth4_ensemble512'] = {
        'workload': 'FMO 7-site HEOM K=1 level_cutoff=4, %d static-disorder members: %d ADOs x 49 '
                    'per member, 10 output intervals' % (E, eom.n_ado),
        'rhs_per_s': rhs_per_s,
        'state_steps_per_s': best['steps'] / (best['kernel_ms'] * 1e-3),
        'rhs_per_state_step': best['rhs'] / max(1, best['steps']),
        'kernel_ms': best['kernel_ms'], 'setup_s': build_s,
        'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peaks['hbm_gbs'],
                     'unit': 'GB/s', 'frac': achieved / peaks['hbm_gbs'], 'traffic': None,
                     'achieved_is': 'rhs_per_s x 32 D (algorithmic bytes per RHS application)',
                     'peak_source': src, 'algorithmic_bytes_per_rhs': 32 * eom.dim}}
    del eom, model
    return out


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
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
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

    # ---- device-resident leg: generators already in HBM ----------------------
    eom = model.ensemble_eom(E, False, 'ee', member0=rank * E)
    y0_dev = _capi.to_device(y0).reshape(1, -1).expand(E, -1).contiguous()
    gens = np.arange(E)

    def resident_step():
        eom.__dict__.pop('_propagators', None)    # rebuild exp(L dt) every step
        out = eom.propagate(y0_dev, t, generators=gens, return_device=True)
        mean = engine.reduce_members(out, 1.0 / (E * world))
        return reduce_across(mean)

    # nvidia-smi needs ~0.1-0.2 s before its first sample: start it ahead of the warm-up and keep
    # only the samples that arrive inside the timed window
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        resident_step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    engine.PropagationStats.reset()
    launches0 = _capi.kernel_launches()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    torch.cuda.synchronize()
    t_a = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        result = resident_step()
    e1.record()
    torch.cuda.synchronize()
    t_b = time.perf_counter()
    if world > 1:
        dist.barrier()
    elapsed_ms = e0.elapsed_time(e1)
    launches = _capi.kernel_launches() - launches0
    stats = engine.PropagationStats
    stats_expm_ms, stats_expm_gemms, stats_expm_builds = stats.expm_ms, stats.expm_gemms, stats.expm_builds
    kernel_ms = stats.kernel_ms / max(1, stats.propagations)
    rhs_per_launch = stats.rhs_evaluations / max(1, stats.propagations)
    steps_per_launch = stats.accepted_steps / max(1, stats.propagations)
    clocks_window = 'timed region'
    if rank == 0 and sampler.proc is not None and sampler.count_between(t_a, t_b) < 3:
        # the timed region is shorter than a few sampling periods: keep the identical load
        # running (untimed) until the sampler has seen it at least a few times
        t_end = time.perf_counter() + 0.6
        while time.perf_counter() < t_end:
            eom.__dict__.pop('_propagators', None)
            engine.reduce_members(eom.propagate(y0_dev, t, generators=gens, return_device=True), 1.0 / (E * world))
            torch.cuda.synchronize()
        t_b = time.perf_counter()
        clocks_window = 'timed region + 0.6 s of the identical step repeated untimed (timed region shorter than 3 sampling periods)'
    clocks = sampler.stop(t_a, t_b) if rank == 0 else None
    if clocks is not None:
        clocks['window'] = clocks_window
    if world > 1:
        tmax = torch.tensor([elapsed_ms], dtype=torch.float64, device='cuda')
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        elapsed_ms = float(tmax.item())
    ms_per_step = elapsed_ms / args.steps
    value = world * steps_per_launch / (ms_per_step * 1e-3)
    trace_err = None
    if rank == 0:
        rho_dev = result.cpu().numpy().reshape(len(t), 7, 7)
        trace_err = float(np.abs(np.einsum('tii->t', rho_dev) - 1).max())

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

    del eom
    torch.cuda.empty_cache()
    e2e_once()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e2e_times = []
    for _ in range(args.e2e_steps):
        t0 = time.perf_counter()
        rho_host = e2e_once()
        torch.cuda.synchronize()
        e2e_times.append(time.perf_counter() - t0)
    if world > 1:
        dist.barrier()
    e2e_s = float(np.mean(e2e_times))
    if world > 1:
        tmax = torch.tensor([e2e_s], dtype=torch.float64, device='cuda')
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        e2e_s = float(tmax.item())
    e2e_value = world * E * (len(t) - 1) / e2e_s
    # bytes that actually cross PCIe per call: the user's inputs are the model (7x7 H, bath
    # parameters, seed), psi0 and the time grid -- the members' disorder is replayed from the
    # seed ON THE DEVICE (as the reference replays it inside the call on the host); plus the
    # per-launch column tables (3 ints per member) of the propagate entry point
    h2d = 49 * 16 + E * 4 * 3 + len(t) * 8 + 7 * 7 * 8 + 4 * 7 * 8 + 49 * 8 + 8
    d2h = len(t) * 49 * 16

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks, peak_src = measured_peaks()
    fp64_peak = measure_fp64_peak(torch)
    # dominant kernel of the step: the tensor-core propagator build (dense_expm_kernel);
    # algorithmic flops = complex M x M GEMMs x 8 M^3 (M = 49, padding not counted)
    expm_ms = stats_expm_ms / max(1, stats_expm_builds)
    flops_per_launch = 8.0 * 49 ** 3 * stats_expm_gemms / max(1, stats_expm_builds)
    achieved_tf = flops_per_launch / (expm_ms * 1e-3) / 1e12 if expm_ms > 0 else 0.0
    map_flops = 8.0 * 49 * 49 * rhs_per_launch
    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_per_step,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'complex128', 'data': 'synthetic',
        'config': workload_config(E, world),
        'rhs_per_state_step': rhs_per_launch / max(1.0, steps_per_launch),
        'rhs_per_s': world * rhs_per_launch / (ms_per_step * 1e-3),
        'grid_steps_per_s': world * E * (len(t) - 1) / (ms_per_step * 1e-3),
        'trace_error': trace_err,
        'clocks': clocks,
        'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': h2d,
                'd2h_bytes_per_step': d2h, 'seconds_per_step': e2e_s,
                'seconds_each_step': [round(x, 5) for x in e2e_times],
                'path': 'simulate_dynamics(model, psi0, duration, ensemble_size) from host objects: '
                        'seed -> device replay of the seeded disorder streams -> K5 (Jacobi '
                        'eigensystems + Redfield generators) -> K1\' propagators -> K1/K4 '
                        'stepping -> K6 mean -> D2H of the averaged density matrices'},
        'gpu_launches': int(launches),
        'roofline': None,
    }
    # the step is two kernels of about equal duration: the DMMA propagator build and the
    # propagator stepping on the FP64 pipe; `roofline` describes whichever took longer in
    # THIS run, the other is listed beside it.  Both are measured against the same FP64
    # ceiling (on B200 the DGEMM/DMMA ceiling equals the FP64 FMA ceiling).
    peak_note = ('cuBLAS FP64 GEMM 6144^3 measured in this run (cutlass d884 DMMA kernel; '
                 'FP64 is not in MEASURED_PEAKS.json)')
    map_tf = map_flops / (kernel_ms * 1e-3) / 1e12
    k_expm = {'bound': 'tensor', 'achieved': achieved_tf, 'peak': fp64_peak,
              'unit': 'TFLOP/s', 'frac': achieved_tf / fp64_peak,
              'traffic': (ncu_traffic('dense_expm_per_member') or 0) * E or None,
              'kernel': 'dense_expm2_kernel<7,13> (FP64 DMMA m8n8k4, three real products per complex one: exp(L dt) per member, '
                        'Paterson-Stockmeyer degree 14 + squarings)',
              'kernel_ms': expm_ms, 'share_of_step': expm_ms / ms_per_step,
              'algorithmic_flops_per_launch': flops_per_launch, 'peak_source': peak_note}
    k_map = {'bound': 'tensor', 'achieved': map_tf, 'peak': fp64_peak,
             'unit': 'TFLOP/s', 'frac': map_tf / fp64_peak,
             'traffic': (ncu_traffic('dense_map_per_member') or 0) * E or None,
             'kernel': 'dense_map_kernel<4,13,1,2> (y <- P y stepping with P in registers; DFMA on '
                       'the FP64 pipe, measured against the same FP64 ceiling as SURVEY 8d asks '
                       'for the dense L.Y contraction)',
             'kernel_ms': kernel_ms, 'share_of_step': kernel_ms / ms_per_step,
             'algorithmic_flops_per_launch': map_flops, 'peak_source': peak_note}
    first, second = (k_map, k_expm) if kernel_ms >= expm_ms else (k_expm, k_map)
    line['roofline'] = first
    line['other_kernels'] = {second['kernel']: second,
                             'share_note': 'per step: propagator build + stepping + member '
                                           'reduction; see profiles/ for the ncu launch list'}
    if world == 1 and not args.no_heom:
        line['heom'] = heom_leg(torch, qb, systems, engine)
    if world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        n = args.cpu_members or 160 * cores
        r = cpu_reference_rate(n)
        line['cpu_baseline'] = {
            'value': r['grid_steps_per_s'], 'unit': UNIT, 'cores': r['workers'],
            'kind': 'port',
            'sample': '%d of %d members x 196 output intervals (oracle: generator '
                      'rebuild + ZVODE rtol=1e-10 per member), %.1f s on %d worker '
                      'processes' % (n, E, r['wall_s'], r['workers'])}
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
