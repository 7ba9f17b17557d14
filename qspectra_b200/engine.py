"""
Device equations of motion: thin Python objects around the C-ABI handles.

A ``DeviceEOM`` is what ``DynamicalModel.equation_of_motion`` returns.  It is a
callable ``f(t, y)`` like the reference's closures
(dynamics/liouville_space.py:339-341, heom.py:241-244) *and* it knows how to
run the whole ``integrate`` loop (simulate/utils.py:53-109) on the GPU for a
batch of initial states.  PyTorch is used for device memory and streams only.
"""
import ctypes as C
import os

import numpy as np

from . import _capi
from ._capi import IntegratorError  # noqa: F401  (re-export)


class PropagationStats(object):
    """Counters of the most recent propagations (bench.py reads these after
    ``flush()``).  Propagator builds and propagator stepping do not synchronise
    with the host; their device times are collected lazily: ``flush()`` (or reading
    ``eom.last`` / ``prop.build_ms``) waits for the recorded events."""
    rhs_evaluations = 0
    accepted_steps = 0
    kernel_ms = 0.0
    propagations = 0
    expm_ms = 0.0
    expm_gemms = 0
    expm_builds = 0
    hermitian_builds = 0       # propagator builds that went through the Hermitian-coordinate (real) form
    form_ms = 0.0              # of expm_ms: the change of coordinates L -> G (hermitian_form_kernel)
    _pending = []          # (reference to an EOM, method name, argument) whose device times are outstanding
    MAX_PENDING = 256
    #: bench.py: keep the objects alive until flush() so that no timing is lost; by default
    #: only weak references are held (no device memory is retained for statistics) and an
    #: entry that overflows the list is resolved only if that does not block
    keep_alive = False

    @classmethod
    def reset(cls):
        cls.flush()
        cls.rhs_evaluations = 0
        cls.accepted_steps = 0
        cls.kernel_ms = 0.0
        cls.propagations = 0
        cls.expm_ms = 0.0
        cls.expm_gemms = 0
        cls.expm_builds = 0
        cls.hermitian_builds = 0
        cls.form_ms = 0.0

    @classmethod
    def defer(cls, obj, name, arg=None):
        import weakref
        cls._pending.append((obj if cls.keep_alive else weakref.ref(obj), not cls.keep_alive, name, arg))
        while len(cls._pending) > cls.MAX_PENDING:
            cls._resolve(cls._pending.pop(0), False)

    @classmethod
    def _resolve(cls, entry, blocking):
        ref, weak, name, arg = entry
        obj = ref() if weak else ref
        if obj is None or getattr(obj, '_h', None) is None:
            return
        if not blocking and not obj._stats_ready():
            return
        getattr(obj, name)(*(() if arg is None else (arg,)))

    @classmethod
    def flush(cls):
        while cls._pending:
            cls._resolve(cls._pending.pop(0), True)


class LinearMap(object):
    """A linear save_func: ``state -> matrix . state`` (the reference's
    ``SystemOperator.commutator`` etc. are exactly such maps,
    liouville_space.py:183-209).  ``per_ado`` marks HEOM block-diagonal maps."""

    def __init__(self, matrix, n_ado=1, ado0_only=False):
        self.matrix = np.asarray(matrix)
        self.n_ado = n_ado
        self.ado0_only = ado0_only

    def __call__(self, state):
        state = np.asarray(state)
        m = self.matrix
        k = m.shape[-1]
        if self.ado0_only:
            return np.tensordot(state[..., :k], m, axes=(-1, -1))
        if self.n_ado == 1:
            return np.tensordot(state, m, axes=(-1, -1))
        blocks = state.reshape(state.shape[:-1] + (self.n_ado, k))
        out = np.tensordot(blocks, m, axes=(-1, -1))
        return out.reshape(state.shape[:-1] + (-1,))

    #: reference code calls ``V.commutator(state)`` and also passes
    #: ``V.commutator`` as save_func; both work with this object.
    dot = __call__


def resolve_method(method_name, lti):
    """Map the reference's ``method_name`` to a device integrator.  'zvode'
    (the reference default, simulate/utils.py:53) selects the engine's default:
    adaptive Taylor for constant generators, DOPRI5 otherwise."""
    name = (method_name or 'zvode').lower()
    if name in ('zvode', 'vode', 'auto', 'lsoda'):
        return 'taylor' if lti else 'dopri5'
    if name in ('dop853',):
        return 'dopri5'
    if name not in _capi.METHODS:
        raise ValueError('unknown integration method %r (device methods: '
                         'taylor, rk4, dopri5, expm)' % method_name)
    if name in ('taylor', 'map', 'poly') and not lti:
        raise ValueError('%s needs a time-independent linear generator' % name)
    return name


class DeviceEOM(object):
    """Base class: a generator resident on the GPU."""
    lti = True
    n_generators = 1
    dim = 0

    def __call__(self, t, y):
        y = np.asarray(y, dtype=complex)
        out = self.apply(y.reshape(1, -1) if y.ndim == 1 else y)
        return out.reshape(y.shape)

    # subclasses: _apply_dev(y_dev, dy_dev, n, gens) and _propagate(args)
    def apply(self, y, generators=None):
        torch = _capi.torch_cuda()
        y_dev = _capi.to_device(y).reshape(-1, self.dim)
        dy = torch.empty_like(y_dev)
        _, gptr = _capi.int32_ptr(generators)
        self._apply_dev(y_dev, dy, y_dev.shape[0], gptr)
        return _capi.to_host(dy)

    def propagate(self, y0, t, t0=None, method='zvode', save=None,
                  generators=None, pulses=None, pulse_ops=None, rtol=None,
                  atol=None, rk4_substeps=None, return_device=False,
                  save_index=None, **ignored):
        """Integrate a batch: y0 (B, dim) -> (B, len(t), saved_dim).

        save      : None | LinearMap | ('ado0',) | ndarray (rows, dim[/n_ado])
                    or stack (n_generators, rows, dim)
        generators: int array (B,) choosing the generator of each column
        save_index: int array (B,) choosing the matrix of a save stack for each
                    column (dense generators; default: the column's generator)
        pulses    : list of (scale, detuning, t_peak, inv_two_sigma_sq, conj)
        pulse_ops : (n_sets, n_pulses, d, d) complex
        """
        torch = _capi.torch_cuda()
        t = np.ascontiguousarray(t, dtype=np.float64)
        if t.ndim != 1 or t.size == 0:
            raise ValueError('t must be a non-empty 1D array')
        y0_dev = _capi.to_device(y0).reshape(-1, self.dim)
        B = y0_dev.shape[0]
        lti = self.lti and not pulses
        method = resolve_method(method, lti)

        args = _capi.QsxPropagateArgs()
        args.n_columns = B
        args.n_times = t.size
        args.t_host = t.ctypes.data_as(C.POINTER(C.c_double))
        args.t0 = float(t[0] if t0 is None else t0)
        args.y0_dev = y0_dev.data_ptr()
        garr, gptr = _capi.int32_ptr(generators)
        if garr is not None and garr.shape != (B,):
            raise ValueError('generators must have one entry per column')
        args.generator_of_column_host = gptr
        args.method = _capi.METHODS[method]
        args.rtol = float(rtol) if rtol else 0.0
        args.atol = float(atol) if atol else 0.0
        args.rk4_substeps = int(rk4_substeps) if rk4_substeps else 0

        keep = []
        saved_dim = self._configure_save(args, save, keep)
        sarr, sptr = _capi.int32_ptr(save_index)
        if sarr is not None:
            if sarr.shape != (B,):
                raise ValueError('save_index must have one entry per column')
            if not isinstance(self, (DenseEOM, HeomEOM)):
                raise ValueError('per-column save matrices need a dense or HEOM generator')
            args.save_of_column_host = sptr
        n_p = len(pulses) if pulses else 0
        if n_p > _capi.MAX_PULSES:
            raise ValueError('at most %d pulses' % _capi.MAX_PULSES)
        args.n_pulses = n_p
        if n_p:
            ops = _capi.to_device(pulse_ops)
            if ops.dim() == 3:
                ops = ops.unsqueeze(0)
            args.pulse_ops_dev = ops.data_ptr()
            args.n_pulse_sets = ops.shape[0]
            keep.append(ops)
            for i, p in enumerate(pulses):
                (args.pulses[i].scale, args.pulses[i].detuning,
                 args.pulses[i].t_peak, args.pulses[i].inv_two_sigma_sq) = p[:4]
                args.pulses[i].conjugate = int(bool(p[4]))
        out = torch.empty((B, t.size, saved_dim), dtype=torch.complex128,
                          device=y0_dev.device)
        args.out_dev = out.data_ptr()
        self._propagate(args)
        PropagationStats.rhs_evaluations += int(args.rhs_evaluations)
        PropagationStats.accepted_steps += int(args.accepted_steps)
        PropagationStats.propagations += 1
        info = dict(rhs=int(args.rhs_evaluations), steps=int(args.accepted_steps),
                    kernel_ms=float(args.kernel_ms), method=method)
        if info['kernel_ms'] < 0:
            # queued without a host synchronisation (propagator stepping): the device time
            # is collected when somebody asks for it
            info['kernel_ms'] = None
            PropagationStats.defer(self, '_resolve_last', info)
        else:
            PropagationStats.kernel_ms += info['kernel_ms']
        self.last = info
        return out if return_device else _capi.to_host(out)

    def _resolve_last(self, info):
        if info.get('kernel_ms') is None:
            info['kernel_ms'] = self._last_kernel_ms()
            PropagationStats.kernel_ms += info['kernel_ms']

    def _last_kernel_ms(self):
        raise RuntimeError('%s never defers its statistics' % type(self).__name__)

    @property
    def last(self):
        """Statistics of the most recent propagation: rhs, steps, kernel_ms, method."""
        src = self.__dict__.get('_last_src')
        if src is not None:
            return dict(src.last, method='expm')
        info = self.__dict__.get('_last')
        if info is not None:
            self._resolve_last(info)
        return info

    @last.setter
    def last(self, info):
        self.__dict__['_last'] = info
        self.__dict__['_last_src'] = None

    def _configure_save(self, args, save, keep):
        if save is None:
            args.save_mode = _capi.SAVE_STATE
            return self.dim
        if isinstance(save, tuple) and save and save[0] == 'ado0':
            return self._configure_ado0(args)
        matrix = save.matrix if isinstance(save, LinearMap) else save
        return self._configure_matrix(args, matrix, save, keep)

    def _configure_ado0(self, args):
        raise ValueError('ado0 save is only defined for HEOM generators')

    def _configure_matrix(self, args, matrix, save, keep):
        S = _capi.to_device(matrix)
        if S.dim() == 1:
            S = S.unsqueeze(0)
        if S.dim() == 2:
            S = S.unsqueeze(0)
        if S.shape[-1] != self.dim:
            raise ValueError('save matrix has %d columns, state has %d'
                             % (S.shape[-1], self.dim))
        keep.append(S)
        args.save_mode = _capi.SAVE_MATRIX
        args.save_rows = S.shape[1]
        args.save_dev = S.data_ptr()
        args.n_save = S.shape[0]
        return S.shape[1]


class HermitianTrajectory(object):
    """Trajectories in Hermitian coordinates (csrc/dense_real.cu): a real CUDA tensor
    (..., n_times, MS) with MS = dim rounded up to even, the transposition permutation of the
    subspace and the propagators that produced it (their device-side Hermiticity check is
    settled when the complex form is read on the host)."""

    def __init__(self, data, perm, dim, source=None):
        self.data, self.perm, self.dim, self.source = data, perm, dim, source

    def to_complex(self):
        """complex128 CUDA tensor (..., n_times, dim) in the reference's vectorisation."""
        torch = _capi.torch_cuda()
        lead = tuple(self.data.shape[:-1])
        rows = int(np.prod(lead, dtype=np.int64))
        out = torch.empty(lead + (self.dim,), dtype=torch.complex128, device=self.data.device)
        _capi.check(_capi.lib().qsx_hermitian_unpack(
            self.data.data_ptr(), self.dim, rows, self.data.shape[-1],
            self.perm.ctypes.data_as(C.POINTER(C.c_int32)), out.data_ptr(),
            _capi.current_stream_ptr()))
        return out


class HermitianPropagators(object):
    """exp(G dt) of every generator of a DenseEOM in Hermitian coordinates, where a generator
    that commutes with Hermitian conjugation is REAL (csrc/dense_real.cu): one real tensor-core
    product per complex one, a quarter of the multiply-adds per step."""
    TOLERANCE = 1e-10
    _h = True           # PropagationStats protocol: "still alive"

    def __init__(self, eom, dt):
        torch = _capi.torch_cuda()
        lib, stream = _capi.lib(), _capi.current_stream_ptr()
        n, M = eom.n_generators, eom.dim
        self.perm = np.ascontiguousarray(eom.hermitian_perm, dtype=np.int32)
        self.dim, self.n_generators = M, n
        self.P = torch.empty((n, M, M), dtype=torch.float64, device='cuda')
        self.defect = torch.zeros(4, dtype=torch.float64, device='cuda')
        self.counter = torch.zeros(1, dtype=torch.int64, device='cuda')
        self.host = self._pinned_slot(torch)
        self.events = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        pptr = self.perm.ctypes.data_as(C.POINTER(C.c_int32))
        self.events[0].record()
        if os.environ.get('QSX_HERMITIAN_TWO_KERNELS'):
            # A/B runs: change of coordinates and series as separate launches, G through global memory
            G = torch.empty((n, M, M), dtype=torch.float64, device='cuda')
            gnorm = torch.empty(n, dtype=torch.float64, device='cuda')
            _capi.check(lib.qsx_dense_hermitian_form(eom._h, pptr, G.data_ptr(), gnorm.data_ptr(),
                                                     self.defect.data_ptr(), stream))
            self.events[3].record()
            _capi.check(lib.qsx_real_expm(G.data_ptr(), gnorm.data_ptr(), M, n, float(dt),
                                          self.P.data_ptr(), self.counter.data_ptr(), stream))
        else:
            self.events[3].record()
            _capi.check(lib.qsx_dense_hermitian_expm(eom._h, pptr, float(dt), self.P.data_ptr(),
                                                     self.defect.data_ptr(), self.counter.data_ptr(),
                                                     stream))
        self.events[1].record()
        self.last_event = self.events[1]
        PropagationStats.expm_builds += 1
        PropagationStats.hermitian_builds += 1
        PropagationStats.defer(self, '_resolve_build')

    _ring, _ring_next = None, 0

    @classmethod
    def _pinned_slot(cls, torch):
        """five pinned doubles from a process-lifetime ring (cudaHostAlloc per build would cost
        more host time than the build's launches); a slot is reused after 256 further builds"""
        if cls._ring is None:
            cls._ring = torch.zeros((256, 5), dtype=torch.float64).pin_memory()
        slot = cls._ring[cls._ring_next % 256]
        cls._ring_next += 1
        slot.zero_()
        return slot

    def snapshot(self):
        """queue the copy of the device-side check values to pinned host memory"""
        torch = _capi.torch_cuda()
        self.host[:4].copy_(self.defect, non_blocking=True)
        self.host[4:].copy_(self.counter.to(torch.float64), non_blocking=True)
        self.events[2].record()
        self.last_event = self.events[2]

    def _stats_ready(self):
        return self.last_event.query()

    def ok(self):
        """after a synchronisation that follows snapshot(): did generators and states pass the
        Hermiticity check?"""
        d = self.host.numpy()
        return bool(d[0] <= self.TOLERANCE * d[1] and d[2] <= self.TOLERANCE * max(d[3], 1e-300))

    def _resolve_build(self):
        if '_build' not in self.__dict__:
            self.snapshot()
            self.last_event.synchronize()
            ms = self.events[0].elapsed_time(self.events[1])
            gemms = int(self.host[4].item())
            self.__dict__['_build'] = (ms, gemms)
            PropagationStats.expm_ms += ms
            PropagationStats.form_ms += self.events[0].elapsed_time(self.events[3])
            PropagationStats.expm_gemms += gemms
            if not self.ok() and not self.__dict__.get('_fallback'):
                raise RuntimeError(
                    'Hermitian-coordinate propagation was used for a generator or state that is '
                    'not compatible with Hermitian conjugation (|Im G| %.3e of %.3e, |Im u0| %.3e '
                    'of %.3e); set QSX_NO_HERMITIAN_FORM=1' % tuple(self.host[:4].tolist()))
        return self.__dict__['_build']


class CapturedEnsembleStep(object):
    """One ensemble step in Hermitian coordinates as a CUDA graph: change of coordinates and real
    tensor-core propagators of every generator (rebuilt by every replay), packing of the initial
    states, stepping, member mean and the conversion of the mean to the complex vectorisation --
    six kernels behind one ``cudaGraphLaunch``.  For small ensembles (a strongly sharded job) the
    host-side launch path of the eager calls (~0.6 ms of Python / ctypes per step) is longer than
    the kernels; a replay costs ~10 us.  Buffers are owned by the object; ``run()`` returns the
    (n_times, dim) complex128 CUDA tensor that the next replay overwrites."""

    def __init__(self, eom, y0_dev, t, scale):
        torch = _capi.torch_cuda()
        if eom.hermitian_perm is None or eom.dim > DenseEOM.HERMITIAN_MAX_DIM or eom.heisenberg_picture:
            raise ValueError('generator has no Hermitian-coordinate form')
        dt = eom._uniform_step(t, None)
        if dt is None:
            raise ValueError('propagator stepping needs a uniform output grid')
        self.eom, self.dt, self.scale = eom, float(dt), float(scale)
        M, n = eom.dim, eom.n_generators
        self.M, self.MS, self.nt = M, M + (M & 1), len(t)
        self.perm = np.ascontiguousarray(eom.hermitian_perm, dtype=np.int32)
        self.y0 = _capi.to_device(y0_dev).reshape(-1, M).contiguous()
        self.B = self.y0.shape[0]
        if self.B != n:
            raise ValueError('one column per generator expected')
        f64 = dict(dtype=torch.float64, device=self.y0.device)
        self.P = torch.empty((n, M, M), **f64)
        self.check = torch.zeros(5, **f64)           # defect[0..3] + the product counter (as int64 bits)
        self.u0 = torch.empty((self.B, self.MS), **f64)
        self.out = torch.empty((self.B, self.nt, self.MS), **f64)
        self.mean_real = torch.empty((self.nt, self.MS), **f64)
        self.mean = torch.empty((self.nt, M), dtype=torch.complex128, device=self.y0.device)
        # warm-up on a side stream (scratch pools and kernel attributes settle outside the capture)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            self._enqueue()
            self._enqueue()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._enqueue()
        self.launches_per_run = 7

    def _enqueue(self):
        lib, stream = _capi.lib(), _capi.current_stream_ptr()
        pptr = self.perm.ctypes.data_as(C.POINTER(C.c_int32))
        M, MS, B, nt = self.M, self.MS, self.B, self.nt
        self.check.zero_()
        _capi.check(lib.qsx_dense_hermitian_expm(self.eom._h, pptr, self.dt, self.P.data_ptr(),
                                                 self.check.data_ptr(), self.check[4:].data_ptr(), stream))
        _capi.check(lib.qsx_hermitian_pack(self.y0.data_ptr(), M, B, pptr, MS, self.u0.data_ptr(),
                                           self.check.data_ptr(), stream))
        _capi.check(lib.qsx_real_map(self.P.data_ptr(), M, B, None, B, self.u0.data_ptr(), nt, MS,
                                     self.out.data_ptr(), stream))
        _capi.check(lib.qsx_reduce_members(self.out.data_ptr(), B, nt * MS // 2, self.scale,
                                           self.mean_real.data_ptr(), stream))
        _capi.check(lib.qsx_hermitian_unpack(self.mean_real.data_ptr(), M, nt, MS, pptr,
                                             self.mean.data_ptr(), stream))

    def run(self):
        self.graph.replay()
        return self.mean

    def verify(self):
        """synchronises; raises if generators or states failed the device-side Hermiticity check"""
        d = self.check.cpu().numpy()
        tol = HermitianPropagators.TOLERANCE
        if not (d[0] <= tol * d[1] and d[2] <= tol * max(d[3], 1e-300)):
            raise RuntimeError('generator or state is not compatible with Hermitian conjugation '
                               '(|Im G| %.3e of %.3e, |Im u0| %.3e of %.3e)' % tuple(d[:4]))
        return int(d[4:].view(np.int64)[0])          # real M x M products of the last replay


class _MapTiming(object):
    """deferred device time of one stepping call in Hermitian coordinates"""
    _h = True

    def __init__(self, e0, e1, info):
        self.e0, self.e1, self.info = e0, e1, info

    def _stats_ready(self):
        return self.e1.query()

    def _resolve(self):
        if self.info.get('kernel_ms') is None:
            self.e1.synchronize()
            self.info['kernel_ms'] = self.e0.elapsed_time(self.e1)
            PropagationStats.kernel_ms += self.info['kernel_ms']


class DenseEOM(DeviceEOM):
    """Batched dense Liouvillians L[g] (n_generators, M, M) staged on the GPU
    (kernel K1/K4, csrc/dense.cu)."""

    def __init__(self, L, heisenberg_picture=False):
        torch = _capi.torch_cuda()
        lib = _capi.lib()
        on_device = isinstance(L, torch.Tensor)
        if on_device:
            Ld = L.to(torch.complex128).cuda().contiguous()
            if Ld.dim() == 2:
                Ld = Ld.unsqueeze(0)
            shape, ptr = tuple(Ld.shape), Ld.data_ptr()
        else:
            Lh = np.ascontiguousarray(L, dtype=np.complex128)
            if Lh.ndim == 2:
                Lh = Lh[None]
            shape, ptr = Lh.shape, Lh.ctypes.data
        if len(shape) != 3 or shape[1] != shape[2]:
            raise ValueError('generators must have shape (n, M, M)')
        self.n_generators, self.dim = int(shape[0]), int(shape[1])
        self.heisenberg_picture = bool(heisenberg_picture)
        self._h = C.c_void_p()
        _capi.check(lib.qsx_dense_create(C.byref(self._h), self.dim,
                                         self.n_generators, ptr, int(on_device),
                                         int(self.heisenberg_picture),
                                         _capi.current_stream_ptr()))

    def __del__(self):
        h = getattr(self, '_h', None)
        if h and _capi is not None and _capi._lib is not None:
            _capi._lib.qsx_dense_destroy(h)
            self._h = None

    #: largest state dimension of the single-CTA tensor-core propagator kernel
    EXPM_MAX_DIM = 56
    #: wider states (e.g. FMO 'fe', 147) get exp(L dt) from the same series with one
    #: tiled tensor-core GEMM launch per product (csrc/dense_wide.cu) and are stepped by
    #: the CTA-resident kernel with P streamed from L2
    EXPM_LIBRARY_MAX_DIM = 1024

    def propagator(self, dt):
        """DenseEOM holding P_g = exp(L_g dt) for every generator (FP64 tensor
        cores, csrc/dense.cu: dense_expm_kernel); cached per dt."""
        cache = self.__dict__.setdefault('_propagators', {})
        key = float(dt)
        if key not in cache:
            torch = _capi.torch_cuda()
            prop = DenseEOM.__new__(DenseEOM)
            prop.n_generators, prop.dim = self.n_generators, self.dim
            prop.heisenberg_picture = self.heisenberg_picture
            # storage from torch's caching allocator, borrowed by the handle
            prop._storage = (torch.empty((self.n_generators, self.dim, self.dim),
                                         dtype=torch.complex128, device='cuda'),
                             torch.empty(self.n_generators, dtype=torch.float64,
                                         device='cuda'))
            prop._h = C.c_void_p()
            _capi.check(_capi.lib().qsx_dense_expm(
                self._h, key, prop._storage[0].data_ptr(),
                prop._storage[1].data_ptr(), C.byref(prop._h),
                _capi.current_stream_ptr()))
            # no host synchronisation: device time and GEMM count are read back lazily
            PropagationStats.expm_builds += 1
            PropagationStats.defer(prop, '_resolve_build')
            cache[key] = prop
        return cache[key]

    @classmethod
    def from_transposed(cls, Lt_dev):
        """Wrap a CUDA tensor that already has the engine's transposed storage
        Lt[g][c][r] = L_g[r][c] (no copy; e.g. the output of the K5 builders)."""
        torch = _capi.torch_cuda()
        obj = cls.__new__(cls)
        obj.n_generators, obj.dim = int(Lt_dev.shape[0]), int(Lt_dev.shape[1])
        obj.heisenberg_picture = False
        obj._storage = (Lt_dev, torch.empty(obj.n_generators, dtype=torch.float64,
                                            device=Lt_dev.device))
        obj._h = C.c_void_p()
        _capi.check(_capi.lib().qsx_dense_wrap(
            C.byref(obj._h), obj.dim, obj.n_generators, Lt_dev.data_ptr(),
            obj._storage[1].data_ptr(), _capi.current_stream_ptr()))
        return obj

    def _uniform_step(self, t, t0):
        """dt if `t` is a uniform grid starting at t0, else None."""
        t = np.asarray(t, dtype=float)
        if t.size < 3 or (t0 is not None and t0 != t[0]):
            return None
        d = np.diff(t)
        if d[0] <= 0 or np.abs(d - d[0]).max() > 1e-9 * abs(d[0]):
            return None
        return float(d[0])

    #: transposition permutation of the Liouville subspace (set by the models for subspaces
    #: closed under transposition, Schroedinger picture): enables Hermitian-coordinate stepping
    hermitian_perm = None
    HERMITIAN_MAX_DIM = 56

    def _hermitian_state(self, y0):
        """True when the host array y0 (..., dim) is Hermitian under the subspace's transposition."""
        if not isinstance(y0, np.ndarray) or self.hermitian_perm is None:
            return False
        y = y0.reshape(-1, self.dim)
        return bool(np.abs(y[:, self.hermitian_perm] - y.conj()).max() <= 1e-14 * max(np.abs(y).max(), 1e-300))

    def _propagate_hermitian(self, y0, t, dt, generators, return_device, packed):
        """Propagator stepping of Hermitian states in real coordinates (csrc/dense_real.cu)."""
        torch = _capi.torch_cuda()
        lib, stream = _capi.lib(), _capi.current_stream_ptr()
        M = self.dim
        MS = M + (M & 1)
        cache = self.__dict__.setdefault('_propagators', {})
        key = ('hermitian', float(dt))
        if key not in cache:
            cache[key] = HermitianPropagators(self, dt)
        hp = cache[key]
        t = np.ascontiguousarray(t, dtype=np.float64)
        y0_dev = _capi.to_device(y0).reshape(-1, M)
        B = y0_dev.shape[0]
        garr, gptr = _capi.int32_ptr(generators)
        if garr is not None:
            if garr.shape != (B,):
                raise ValueError('generators must have one entry per column')
            if B == self.n_generators and np.array_equal(garr, np.arange(B)):
                gptr = None                 # column c uses generator c: no table upload
        pptr = hp.perm.ctypes.data_as(C.POINTER(C.c_int32))
        u0 = torch.empty((B, MS), dtype=torch.float64, device=y0_dev.device)
        _capi.check(lib.qsx_hermitian_pack(y0_dev.data_ptr(), M, B, pptr, MS, u0.data_ptr(),
                                           hp.defect.data_ptr(), stream))
        out = torch.empty((B, t.size, MS), dtype=torch.float64, device=y0_dev.device)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _capi.check(lib.qsx_real_map(hp.P.data_ptr(), M, self.n_generators, gptr, B, u0.data_ptr(),
                                     t.size, MS, out.data_ptr(), stream))
        e1.record()
        n_rhs = (t.size - 1) * B
        PropagationStats.rhs_evaluations += n_rhs
        PropagationStats.accepted_steps += n_rhs
        PropagationStats.propagations += 1
        info = dict(rhs=n_rhs, steps=n_rhs, kernel_ms=None, method='expm', hermitian_form=True)
        timing = _MapTiming(e0, e1, info)
        info['_timing'] = timing
        PropagationStats.defer(timing, '_resolve')
        self.__dict__['_last'] = info
        self.__dict__['_last_src'] = None
        traj = HermitianTrajectory(out, hp.perm, M, hp)
        if packed:
            return traj
        res = traj.to_complex()
        if return_device:
            return res
        hp.snapshot()
        host = _capi.to_host(res)
        if not hp.ok():
            # not Hermiticity-compatible after all: the caller falls back to the complex path
            hp.__dict__['_fallback'] = True
            self.hermitian_perm = None
            return None
        return host

    def _resolve_last(self, info):
        timing = info.get('_timing')
        if timing is not None:
            timing._resolve()
            return
        DeviceEOM._resolve_last(self, info)

    def propagate(self, y0, t, t0=None, method='zvode', **kw):
        """Adds method 'expm' (propagator stepping: y_{i+1} = exp(L dt) y_i,
        built once per generator on the tensor cores).  It is also what the
        default ('zvode') selects for a uniform grid when it pays off, i.e.
        when the series for exp(L dt) (~16 M matrix-vector equivalents) is
        cheaper than integrating every column step by step."""
        name = (method or 'zvode').lower()
        if name in ('expm', 'zvode', 'auto') and not kw.get('pulses'):
            dt = self._uniform_step(t, t0)
            n_cols = int(np.prod(np.shape(y0)[:-1])) if np.ndim(y0) > 1 else 1
            worth = len(t) * max(1, n_cols // self.n_generators) >= 64
            if dt is not None and self.dim <= self.EXPM_LIBRARY_MAX_DIM and \
                    (name == 'expm' or worth):
                hermitian = kw.pop('hermitian_state', None)
                packed = kw.pop('packed', False)
                if (self.hermitian_perm is not None and self.dim <= self.HERMITIAN_MAX_DIM
                        and not self.heisenberg_picture and kw.get('save') is None
                        and kw.get('save_index') is None
                        and not os.environ.get('QSX_NO_HERMITIAN_FORM')
                        and (self._hermitian_state(y0) if hermitian is None else hermitian)):
                    out = self._propagate_hermitian(y0, t, dt, kw.get('generators'),
                                                    kw.get('return_device', False), packed)
                    if out is not None:
                        return out
                prop = self.propagator(dt)
                out = DeviceEOM.propagate(prop, y0, t, t0=t0, method='map', **kw)
                self.__dict__['_last_src'] = prop      # resolved (and relabelled 'expm') on access
                return out
            if name == 'expm':
                raise ValueError('expm needs a uniform output grid starting at '
                                 't0 and a state dimension <= %d'
                                 % self.EXPM_LIBRARY_MAX_DIM)
        kw.pop('hermitian_state', None)
        kw.pop('packed', None)
        return DeviceEOM.propagate(self, y0, t, t0=t0, method=method, **kw)

    def _apply_dev(self, y, dy, n, gptr):
        _capi.check(_capi.lib().qsx_dense_apply(
            self._h, y.data_ptr(), dy.data_ptr(), n, gptr,
            _capi.current_stream_ptr()))

    def _propagate(self, args):
        _capi.check(_capi.lib().qsx_dense_propagate(
            self._h, C.byref(args), _capi.current_stream_ptr()))

    def _last_kernel_ms(self):
        ms = C.c_double()
        _capi.check(_capi.lib().qsx_dense_last_kernel_ms(self._h, C.byref(ms)))
        return ms.value

    def _stats_ready(self):
        return bool(_capi.lib().qsx_dense_events_ready(self._h))

    def _resolve_build(self):
        """Device time / GEMM count of the qsx_dense_expm call that made this handle."""
        if '_build' not in self.__dict__ and getattr(self, '_h', None):
            ms, gemms = C.c_double(), C.c_uint64()
            _capi.check(_capi.lib().qsx_dense_build_stats(self._h, C.byref(ms), C.byref(gemms)))
            self.__dict__['_build'] = (ms.value, int(gemms.value))
            PropagationStats.expm_ms += ms.value
            PropagationStats.expm_gemms += int(gemms.value)
        return self.__dict__.get('_build', (0.0, 0))

    @property
    def build_ms(self):
        return self._resolve_build()[0]

    @property
    def build_gemms(self):
        return self._resolve_build()[1]


class HeomEOM(DeviceEOM):
    """HEOM hierarchy generator (kernel K2/K4, csrc/heom.cu)."""

    def __init__(self, n_sites, K, level_cutoff, subspace_index, H,
                 coupling_diag, nu, c, temp_corr, unit_convert, modified=False,
                 heisenberg_picture=False):
        lib = _capi.lib()
        _capi.torch_cuda()
        H = np.ascontiguousarray(H, dtype=np.complex128)
        if H.ndim == 2:
            H = H[None]
        idx = np.ascontiguousarray(subspace_index, dtype=np.int64)
        v = np.ascontiguousarray(coupling_diag, dtype=np.float64)
        nu = np.ascontiguousarray(nu, dtype=np.float64)
        c = np.ascontiguousarray(c, dtype=np.complex128)
        cfg = _capi.QsxHeomConfig()
        cfg.n_sites, cfg.K, cfg.level_cutoff = int(n_sites), int(K), int(level_cutoff)
        cfg.n_hilbert, cfg.M = H.shape[-1], idx.size
        cfg.subspace_index = idx.ctypes.data_as(C.POINTER(C.c_int64))
        cfg.n_members = H.shape[0]
        cfg.H = H.ctypes.data
        cfg.coupling_diag = v.ctypes.data_as(C.POINTER(C.c_double))
        cfg.nu = nu.ctypes.data_as(C.POINTER(C.c_double))
        cfg.c = c.ctypes.data
        cfg.temp_corr = float(np.real(temp_corr))
        cfg.unit_convert = float(unit_convert)
        cfg.modified = int(bool(modified))
        cfg.heisenberg = int(bool(heisenberg_picture))
        if v.shape != (n_sites, H.shape[-1]):
            raise ValueError('coupling_diag must have shape (n_sites, N)')
        self._h = C.c_void_p()
        _capi.check(lib.qsx_heom_create(C.byref(self._h), C.byref(cfg),
                                        _capi.current_stream_ptr()))
        self.n_generators = H.shape[0]
        self.M = idx.size
        self.n_ado = int(lib.qsx_heom_ado_count(self._h))
        self.dim = self.n_ado * self.M
        self.bins = n_sites * (K + 1)
        self.heisenberg_picture = bool(heisenberg_picture)

    def __del__(self):
        h = getattr(self, '_h', None)
        if h and _capi is not None and _capi._lib is not None:
            _capi._lib.qsx_heom_destroy(h)
            self._h = None

    def index_maps(self):
        idx = np.empty((self.n_ado, self.bins), dtype=np.int64)
        up = np.empty((self.n_ado, self.bins), dtype=np.int32)
        down = np.empty((self.n_ado, self.bins), dtype=np.int32)
        _capi.check(_capi.lib().qsx_heom_index_maps(
            self._h, idx.ctypes.data, up.ctypes.data, down.ctypes.data))
        return idx, up, down

    def propagate(self, y0, t, t0=None, method='zvode', **kw):
        # constant generator, no pulses: the product-form Taylor propagator
        # (QSX_METHOD_POLY, csrc/heom_row.cuh) is the default; the library runs
        # the adaptive Taylor series where the row tile does not apply
        if (method or 'zvode').lower() in ('zvode', 'vode', 'auto', 'lsoda') \
                and not kw.get('pulses'):
            method = 'poly'
        return DeviceEOM.propagate(self, y0, t, t0=t0, method=method, **kw)

    def _apply_dev(self, y, dy, n, gptr):
        _capi.check(_capi.lib().qsx_heom_apply(
            self._h, y.data_ptr(), dy.data_ptr(), n, gptr,
            _capi.current_stream_ptr()))

    def _propagate(self, args):
        _capi.check(_capi.lib().qsx_heom_propagate(
            self._h, C.byref(args), _capi.current_stream_ptr()))

    def _configure_ado0(self, args):
        args.save_mode = _capi.SAVE_ADO0
        return self.M

    def _configure_matrix(self, args, matrix, save, keep):
        S = np.asarray(matrix)
        if S.ndim == 1:
            S = S[None]
        if S.ndim == 3:
            # stack of per-ADO blocks (n_save, rows, M): column c is saved through
            # S[save_index[c]] (e.g. one dipole commutator per polarisation configuration)
            if S.shape[-1] != self.M:
                raise ValueError('HEOM save matrix must act on one ADO '
                                 '(%d columns), got %d' % (self.M, S.shape[-1]))
            Sd = _capi.to_device(np.ascontiguousarray(S))
            keep.append(Sd)
            args.save_mode = _capi.SAVE_MATRIX
            args.save_rows = S.shape[1]
            args.save_dev = Sd.data_ptr()
            args.n_save = S.shape[0]
            return self.n_ado * S.shape[1]
        if isinstance(save, LinearMap) and save.ado0_only:
            # expectation values read ADO 0 only (heom.py:55-58): embed the row
            # vector in a block that is applied to every ADO and keep block 0
            raise ValueError('use save=("ado0",) and contract on the host, or '
                             'the Heisenberg picture, for HEOM expectation '
                             'values')
        if S.shape[-1] != self.M:
            raise ValueError('HEOM save matrix must act on one ADO '
                             '(%d columns), got %d' % (self.M, S.shape[-1]))
        Sd = _capi.to_device(S)
        keep.append(Sd)
        args.save_mode = _capi.SAVE_MATRIX
        args.save_rows = S.shape[0]
        args.save_dev = Sd.data_ptr()
        args.n_save = 1
        return self.n_ado * S.shape[0]


class ZofeEOM(DeviceEOM):
    """ZOFE master equation (kernel K3/K4, csrc/zofe.cu); nonlinear, so the
    integrators are DOPRI5 (default) and RK4."""
    lti = False

    def __init__(self, H, coupling_diag, Gamma, w, unit_convert,
                 ham_hermit=False, rho_hermit=False):
        lib = _capi.lib()
        _capi.torch_cuda()
        H = np.ascontiguousarray(H, dtype=np.complex128)
        if H.ndim == 2:
            H = H[None]
        v = np.ascontiguousarray(coupling_diag, dtype=np.float64)
        Gamma = np.ascontiguousarray(Gamma, dtype=np.complex128)
        w = np.ascontiguousarray(w, dtype=np.complex128)
        if Gamma.shape != w.shape or Gamma.shape[1] != v.shape[0] \
                or v.shape[1] != H.shape[-1]:
            raise ValueError('inconsistent ZOFE array shapes')
        cfg = _capi.QsxZofeConfig()
        cfg.n_states, cfg.n_sites, cfg.n_pm = H.shape[-1], v.shape[0], Gamma.shape[0]
        cfg.n_members = H.shape[0]
        cfg.H = H.ctypes.data
        cfg.coupling_diag = v.ctypes.data_as(C.POINTER(C.c_double))
        cfg.Gamma, cfg.w = Gamma.ctypes.data, w.ctypes.data
        cfg.unit_convert = float(unit_convert)
        cfg.ham_hermit, cfg.rho_hermit = int(bool(ham_hermit)), int(bool(rho_hermit))
        self._h = C.c_void_p()
        _capi.check(lib.qsx_zofe_create(C.byref(self._h), C.byref(cfg),
                                        _capi.current_stream_ptr()))
        self.n_generators = H.shape[0]
        self.n_states = H.shape[-1]
        self.head = self.n_states ** 2
        self.dim = int(lib.qsx_zofe_state_dim(self._h))

    def __del__(self):
        h = getattr(self, '_h', None)
        if h and _capi is not None and _capi._lib is not None:
            _capi._lib.qsx_zofe_destroy(h)
            self._h = None

    def _apply_dev(self, y, dy, n, gptr):
        _capi.check(_capi.lib().qsx_zofe_apply(
            self._h, y.data_ptr(), dy.data_ptr(), n, gptr,
            _capi.current_stream_ptr()))

    def _propagate(self, args):
        _capi.check(_capi.lib().qsx_zofe_propagate(
            self._h, C.byref(args), _capi.current_stream_ptr()))

    def _configure_ado0(self, args):      # "head" = the density-matrix part
        args.save_mode = _capi.SAVE_ADO0
        return self.head

    def _configure_matrix(self, args, matrix, save, keep):
        S = np.asarray(matrix)
        if S.ndim == 1:
            S = S[None]
        if S.shape[-1] != self.head:
            raise ValueError('ZOFE save matrix must act on vec(rho) (%d columns)'
                             % self.head)
        Sd = _capi.to_device(S)
        keep.append(Sd)
        args.save_mode = _capi.SAVE_MATRIX
        args.save_rows = S.shape[0]
        args.save_dev = Sd.data_ptr()
        args.n_save = 1
        return S.shape[0]


def reduce_members(batch_dev, scale=1.0):
    """out[...] = scale * sum_m batch[m, ...] on the device (kernel K6)."""
    torch = _capi.torch_cuda()
    if isinstance(batch_dev, HermitianTrajectory):
        # the member sum commutes with the (linear) change of coordinates: reduce the real
        # rows (viewed as complex pairs, MS is even), then unpack the mean only
        data = batch_dev.data.contiguous()
        B = data.shape[0]
        red = reduce_members(torch.view_as_complex(data.reshape(B, -1, 2)), scale)
        mean = HermitianTrajectory(torch.view_as_real(red).reshape(data.shape[1:]),
                                   batch_dev.perm, batch_dev.dim, batch_dev.source)
        if batch_dev.source is not None:
            batch_dev.source.snapshot()
        return mean.to_complex()
    torch = _capi.torch_cuda()
    batch_dev = batch_dev.contiguous()
    out = torch.empty(batch_dev.shape[1:], dtype=torch.complex128,
                      device=batch_dev.device)
    _capi.check(_capi.lib().qsx_reduce_members(
        batch_dev.data_ptr(), batch_dev.shape[0], out.numel(), float(scale),
        out.data_ptr(), _capi.current_stream_ptr()))
    return out


def response_contract(x_dev, y_dev, weights_dev, total_dev, group_first=None, group_count=None):
    """Kernel K6: total[ab, c] += sum_u w[u] sum_i x[u, ab, i] y[u, c, i] on the FP64 tensor
    cores (csrc/dense_wide.cu).  x: (U, n_ab, K), y: (U, n_c, K), w: (U,), total: (n_ab, n_c),
    all complex128 CUDA tensors.  With groups, y is (G, n_c, K) and group g contracts
    sum_{j < group_count[g]} w[group_first[g] + j] x[group_first[g] + j] with y[g]: units that
    share their y operand cost one product."""
    x_dev, y_dev, weights_dev = x_dev.contiguous(), y_dev.contiguous(), weights_dev.contiguous()
    U, n_ab, K = x_dev.shape
    n_c = y_dev.shape[1]
    if weights_dev.shape != (U,) or total_dev.numel() != n_ab * n_c or not total_dev.is_contiguous():
        raise ValueError('inconsistent shapes in response_contract')
    if group_first is None:
        if y_dev.shape != (U, n_c, K):
            raise ValueError('inconsistent shapes in response_contract')
        _capi.check(_capi.lib().qsx_response_contract(
            x_dev.data_ptr(), y_dev.data_ptr(), weights_dev.data_ptr(), U, n_ab, n_c, K,
            total_dev.data_ptr(), _capi.current_stream_ptr()))
        return total_dev
    first = np.ascontiguousarray(group_first, dtype=np.int32)
    count = np.ascontiguousarray(group_count, dtype=np.int32)
    G = first.size
    if count.shape != (G,) or y_dev.shape != (G, n_c, K):
        raise ValueError('inconsistent group tables in response_contract')
    _capi.check(_capi.lib().qsx_response_contract_grouped(
        x_dev.data_ptr(), U, y_dev.data_ptr(), weights_dev.data_ptr(), G,
        first.ctypes.data_as(C.POINTER(C.c_int32)), count.ctypes.data_as(C.POINTER(C.c_int32)),
        n_ab, n_c, K, total_dev.data_ptr(), _capi.current_stream_ptr()))
    return total_dev


def fourier_transform(x_dev, axis, dt, sign):
    """Kernel K7: X[k] = dt sum_j x[j] exp(sign 2 pi i j (k - n + 1) / (2n - 1)) along
    `axis` of a CUDA tensor (the zero-padded, shifted FFT of simulate/utils.py:154-219
    for samples at t = 0, dt, ...).  Returns a CUDA tensor with 2n - 1 points on that axis."""
    torch = _capi.torch_cuda()
    x_dev = x_dev.contiguous()
    axis = axis % x_dev.dim()
    n = x_dev.shape[axis]
    outer = int(np.prod(x_dev.shape[:axis], dtype=np.int64))
    inner = int(np.prod(x_dev.shape[axis + 1:], dtype=np.int64))
    shape = list(x_dev.shape)
    shape[axis] = 2 * n - 1
    out = torch.empty(shape, dtype=torch.complex128, device=x_dev.device)
    _capi.check(_capi.lib().qsx_fourier_transform(
        x_dev.data_ptr(), outer, n, inner, float(dt), int(sign), out.data_ptr(),
        _capi.current_stream_ptr()))
    return out


def redfield_build(E, U, coupling_diag, bath_kind, temperature, reorg_energy,
                   cutoff_freq, secular, eigen_basis, unit_convert,
                   subspace_index, matsubara_cutoff=1000, transposed=False):
    """Batched Redfield generators on the device (kernel K5).

    E (m, N) float, U (m, N, N): eigen-systems of the members in the rotating
    frame.  Returns a CUDA tensor (m, M, M) = unit_convert * L[idx, idx]."""
    torch = _capi.torch_cuda()
    E_dev = _capi.to_device(np.ascontiguousarray(E, dtype=np.float64),
                            dtype=torch.float64)
    U_dev = _capi.to_device(np.ascontiguousarray(U, dtype=np.complex128))
    m, N = E_dev.shape
    v = np.ascontiguousarray(coupling_diag, dtype=np.float64)
    idx = np.ascontiguousarray(subspace_index, dtype=np.int64)
    bath = _capi.QsxBath(int(bath_kind), int(matsubara_cutoff),
                         float(temperature), float(reorg_energy),
                         float(cutoff_freq))
    out = torch.empty((m, idx.size, idx.size), dtype=torch.complex128,
                      device=E_dev.device)
    _capi.check(_capi.lib().qsx_redfield_build(
        m, N, E_dev.data_ptr(), U_dev.data_ptr(), v.shape[0],
        v.ctypes.data_as(C.POINTER(C.c_double)), C.byref(bath),
        int(bool(secular)), int(bool(eigen_basis)), float(unit_convert),
        idx.size, idx.ctypes.data_as(C.POINTER(C.c_int64)), int(bool(transposed)),
        out.data_ptr(), _capi.current_stream_ptr()))
    return out


def redfield_build_sampled(H0, site_shifts, quanta, rw_freq, coupling_diag,
                           bath_kind, temperature, reorg_energy, cutoff_freq,
                           secular, eigen_basis, unit_convert, subspace_index,
                           matsubara_cutoff=1000, transposed=False):
    """Kernel K5 with on-device Jacobi eigensystems: only the (m, n_sites)
    disorder shifts cross the PCIe bus.  H0: real symmetric lab-frame
    Hamiltonian of the un-sampled system in the Hilbert subspace."""
    torch = _capi.torch_cuda()
    H0 = np.ascontiguousarray(H0, dtype=np.float64)
    shifts = (site_shifts.contiguous() if isinstance(site_shifts, torch.Tensor) else
              _capi.to_device(np.ascontiguousarray(site_shifts, dtype=np.float64),
                              dtype=torch.float64))
    q = np.ascontiguousarray(quanta, dtype=np.float64)
    v = np.ascontiguousarray(coupling_diag, dtype=np.float64)
    idx = np.ascontiguousarray(subspace_index, dtype=np.int64)
    m, N = shifts.shape[0], H0.shape[0]
    bath = _capi.QsxBath(int(bath_kind), int(matsubara_cutoff),
                         float(temperature), float(reorg_energy),
                         float(cutoff_freq))
    out = torch.empty((m, idx.size, idx.size), dtype=torch.complex128,
                      device=shifts.device)
    dptr = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    _capi.check(_capi.lib().qsx_redfield_build_sampled(
        m, N, dptr(H0), shifts.data_ptr(), dptr(q), float(rw_freq), v.shape[0],
        dptr(v), C.byref(bath), int(bool(secular)), int(bool(eigen_basis)),
        float(unit_convert), idx.size,
        idx.ctypes.data_as(C.POINTER(C.c_int64)), int(bool(transposed)),
        out.data_ptr(), _capi.current_stream_ptr()))
    return out
