"""
ctypes binding of ``libqsx.so`` (the C ABI declared in
``include/qspectra_b200.h``).  The library is built in-tree by
``__graft_entry__.build()`` / ``make -C qspectra_b200/csrc``.  There is no CPU
fallback: if the library or a CUDA device is missing, every compute entry point
raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libqsx.so')

QSX_OK = 0
ERR_INVALID, ERR_CUDA, ERR_INTEGRATOR, ERR_UNSUPPORTED = -1, -2, -3, -4
METHOD_TAYLOR, METHOD_RK4, METHOD_DOPRI5, METHOD_MAP, METHOD_POLY = 0, 1, 2, 3, 4
SAVE_STATE, SAVE_MATRIX, SAVE_ADO0 = 0, 1, 2
MAX_PULSES = 4
METHODS = {'taylor': METHOD_TAYLOR, 'rk4': METHOD_RK4, 'dopri5': METHOD_DOPRI5,
           'map': METHOD_MAP, 'poly': METHOD_POLY}


class IntegratorError(Exception):
    """Same role as reference simulate/utils.py:12-13."""


class QsxPulse(C.Structure):
    _fields_ = [('scale', C.c_double), ('detuning', C.c_double),
                ('t_peak', C.c_double), ('inv_two_sigma_sq', C.c_double),
                ('conjugate', C.c_int32), ('_pad', C.c_int32)]


class QsxPropagateArgs(C.Structure):
    _fields_ = [('n_columns', C.c_int32), ('n_times', C.c_int32),
                ('t_host', C.POINTER(C.c_double)), ('t0', C.c_double),
                ('y0_dev', C.c_void_p),
                ('generator_of_column_host', C.POINTER(C.c_int32)),
                ('method', C.c_int32), ('rtol', C.c_double), ('atol', C.c_double),
                ('rk4_substeps', C.c_int32), ('save_mode', C.c_int32),
                ('save_rows', C.c_int32), ('save_dev', C.c_void_p),
                ('n_save', C.c_int32),
                ('save_of_column_host', C.POINTER(C.c_int32)), ('n_pulses', C.c_int32),
                ('pulses', QsxPulse * MAX_PULSES),
                ('pulse_ops_dev', C.c_void_p), ('n_pulse_sets', C.c_int32),
                ('out_dev', C.c_void_p),
                ('rhs_evaluations', C.c_uint64), ('accepted_steps', C.c_uint64),
                ('kernel_ms', C.c_double)]


class QsxHeomConfig(C.Structure):
    _fields_ = [('n_sites', C.c_int32), ('K', C.c_int32),
                ('level_cutoff', C.c_int32), ('n_hilbert', C.c_int32),
                ('M', C.c_int32), ('subspace_index', C.POINTER(C.c_int64)),
                ('n_members', C.c_int32), ('H', C.c_void_p),
                ('coupling_diag', C.POINTER(C.c_double)),
                ('nu', C.POINTER(C.c_double)), ('c', C.c_void_p),
                ('temp_corr', C.c_double), ('unit_convert', C.c_double),
                ('modified', C.c_int32), ('heisenberg', C.c_int32)]


class QsxZofeConfig(C.Structure):
    _fields_ = [('n_states', C.c_int32), ('n_sites', C.c_int32),
                ('n_pm', C.c_int32), ('n_members', C.c_int32),
                ('H', C.c_void_p), ('coupling_diag', C.POINTER(C.c_double)),
                ('Gamma', C.c_void_p), ('w', C.c_void_p),
                ('unit_convert', C.c_double), ('ham_hermit', C.c_int32),
                ('rho_hermit', C.c_int32)]


class QsxBath(C.Structure):
    _fields_ = [('kind', C.c_int32), ('matsubara_cutoff', C.c_int32),
                ('temperature', C.c_double), ('reorg_energy', C.c_double),
                ('cutoff_freq', C.c_double)]


BATH_DEBYE_COMPLEX, BATH_DEBYE_REAL = 0, 1

#: every symbol include/qspectra_b200.h declares (checked by the CPU tests)
EXPORTS = ['qsx_last_error', 'qsx_version', 'qsx_kernel_launches', 'qsx_transfer_bytes',
           'qsx_device_info', 'qsx_dense_create', 'qsx_dense_apply',
           'qsx_dense_propagate', 'qsx_dense_expm', 'qsx_dense_wrap', 'qsx_dense_build_stats',
           'qsx_dense_last_kernel_ms', 'qsx_dense_events_ready',
           'qsx_dense_destroy', 'qsx_dense_hermitian_form', 'qsx_dense_hermitian_expm', 'qsx_real_expm', 'qsx_real_map',
           'qsx_hermitian_pack', 'qsx_hermitian_unpack', 'qsx_heom_create',
           'qsx_heom_ado_count', 'qsx_heom_index_maps', 'qsx_heom_apply',
           'qsx_heom_propagate', 'qsx_heom_destroy', 'qsx_ado_count',
           'qsx_ado_enumerate', 'qsx_redfield_build', 'qsx_redfield_build_sampled',
           'qsx_reduce_members', 'qsx_fourier_transform', 'qsx_response_contract', 'qsx_response_contract_grouped',
           'qsx_sample_streams', 'qsx_sample_gauss_device', 'qsx_zofe_create', 'qsx_zofe_state_dim',
           'qsx_zofe_apply', 'qsx_zofe_propagate', 'qsx_zofe_destroy']

_lib = None


def lib():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            'qspectra_b200: %s is missing -- build it with '
            '`python -c "import __graft_entry__ as g; g.build()"` or '
            '`make -C qspectra_b200/csrc`; there is no CPU fallback'
            % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    L.qsx_last_error.restype = C.c_char_p
    L.qsx_kernel_launches.restype = C.c_uint64
    L.qsx_transfer_bytes.restype = None
    L.qsx_transfer_bytes.argtypes = [C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.qsx_ado_count.restype = C.c_int64
    L.qsx_ado_count.argtypes = [C.c_int32, C.c_int32]
    L.qsx_heom_ado_count.restype = C.c_int64
    L.qsx_heom_ado_count.argtypes = [C.c_void_p]
    L.qsx_ado_enumerate.argtypes = [C.c_int32, C.c_int32, C.c_void_p,
                                    C.c_void_p, C.c_void_p]
    L.qsx_device_info.argtypes = [C.POINTER(C.c_int32), C.POINTER(C.c_int64),
                                  C.POINTER(C.c_int32)]
    L.qsx_dense_create.argtypes = [C.POINTER(C.c_void_p), C.c_int32, C.c_int32,
                                   C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]
    L.qsx_dense_apply.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                  C.POINTER(C.c_int32), C.c_void_p]
    L.qsx_dense_propagate.argtypes = [C.c_void_p, C.POINTER(QsxPropagateArgs),
                                      C.c_void_p]
    L.qsx_dense_expm.argtypes = [C.c_void_p, C.c_double, C.c_void_p, C.c_void_p,
                                 C.POINTER(C.c_void_p), C.c_void_p]
    L.qsx_dense_hermitian_form.argtypes = [C.c_void_p, C.POINTER(C.c_int32), C.c_void_p,
                                           C.c_void_p, C.c_void_p, C.c_void_p]
    L.qsx_dense_hermitian_expm.argtypes = [C.c_void_p, C.POINTER(C.c_int32), C.c_double, C.c_void_p,
                                           C.c_void_p, C.c_void_p, C.c_void_p]
    L.qsx_real_expm.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_double,
                                C.c_void_p, C.c_void_p, C.c_void_p]
    L.qsx_real_map.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_int32),
                               C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p,
                               C.c_void_p]
    L.qsx_hermitian_pack.argtypes = [C.c_void_p, C.c_int32, C.c_int64, C.POINTER(C.c_int32),
                                     C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
    L.qsx_hermitian_unpack.argtypes = [C.c_void_p, C.c_int32, C.c_int64, C.c_int32,
                                       C.POINTER(C.c_int32), C.c_void_p, C.c_void_p]
    L.qsx_dense_build_stats.argtypes = [C.c_void_p, C.POINTER(C.c_double),
                                        C.POINTER(C.c_uint64)]
    L.qsx_dense_last_kernel_ms.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
    L.qsx_dense_events_ready.argtypes = [C.c_void_p]
    L.qsx_dense_wrap.argtypes = [C.POINTER(C.c_void_p), C.c_int32, C.c_int32,
                                 C.c_void_p, C.c_void_p, C.c_void_p]
    L.qsx_dense_destroy.argtypes = [C.c_void_p]
    L.qsx_dense_destroy.restype = None
    L.qsx_heom_create.argtypes = [C.POINTER(C.c_void_p),
                                  C.POINTER(QsxHeomConfig), C.c_void_p]
    L.qsx_heom_index_maps.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p]
    L.qsx_heom_apply.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                 C.POINTER(C.c_int32), C.c_void_p]
    L.qsx_heom_propagate.argtypes = [C.c_void_p, C.POINTER(QsxPropagateArgs),
                                     C.c_void_p]
    L.qsx_heom_destroy.argtypes = [C.c_void_p]
    L.qsx_heom_destroy.restype = None
    L.qsx_redfield_build.argtypes = [
        C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32,
        C.POINTER(C.c_double), C.POINTER(QsxBath), C.c_int32, C.c_int32,
        C.c_double, C.c_int32, C.POINTER(C.c_int64), C.c_int32, C.c_void_p,
        C.c_void_p]
    L.qsx_redfield_build_sampled.argtypes = [
        C.c_int32, C.c_int32, C.POINTER(C.c_double), C.c_void_p,
        C.POINTER(C.c_double), C.c_double, C.c_int32, C.POINTER(C.c_double),
        C.POINTER(QsxBath), C.c_int32, C.c_int32, C.c_double, C.c_int32,
        C.POINTER(C.c_int64), C.c_int32, C.c_void_p, C.c_void_p]
    L.qsx_reduce_members.argtypes = [C.c_void_p, C.c_int32, C.c_int64,
                                     C.c_double, C.c_void_p, C.c_void_p]
    L.qsx_fourier_transform.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_int64,
                                        C.c_double, C.c_int32, C.c_void_p, C.c_void_p]
    L.qsx_response_contract.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                        C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
    L.qsx_response_contract_grouped.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32,
                                                C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_int64,
                                                C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
    L.qsx_zofe_create.argtypes = [C.POINTER(C.c_void_p),
                                  C.POINTER(QsxZofeConfig), C.c_void_p]
    L.qsx_zofe_state_dim.restype = C.c_int64
    L.qsx_zofe_state_dim.argtypes = [C.c_void_p]
    L.qsx_zofe_apply.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                 C.POINTER(C.c_int32), C.c_void_p]
    L.qsx_zofe_propagate.argtypes = [C.c_void_p, C.POINTER(QsxPropagateArgs),
                                     C.c_void_p]
    L.qsx_zofe_destroy.argtypes = [C.c_void_p]
    L.qsx_zofe_destroy.restype = None
    L.qsx_sample_streams.argtypes = [C.POINTER(C.c_uint32), C.c_int32, C.c_int64,
                                     C.c_int32, C.c_int32, C.c_int32,
                                     C.c_void_p, C.c_void_p]
    L.qsx_sample_gauss_device.argtypes = [C.POINTER(C.c_uint32), C.c_int32, C.c_int64,
                                          C.c_int32, C.c_int32, C.c_double,
                                          C.c_void_p, C.c_void_p]
    _lib = L
    return L


def check(status):
    """Map a qsx_status to the reference's exception types."""
    if status == QSX_OK:
        return
    msg = lib().qsx_last_error().decode('utf-8', 'replace')
    if status == ERR_INVALID:
        raise ValueError(msg)
    if status == ERR_INTEGRATOR:
        raise IntegratorError(msg)
    if status == ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise RuntimeError('qspectra_b200 CUDA failure: ' + msg)


def kernel_launches():
    return int(lib().qsx_kernel_launches())


#: bytes moved by the Python side of the engine (tensor uploads in to_device, downloads in to_host)
_py_h2d = 0
_py_d2h = 0


def transfer_bytes():
    """(host->device, device->host) bytes moved so far by the library and by this module."""
    a, b = C.c_uint64(0), C.c_uint64(0)
    lib().qsx_transfer_bytes(C.byref(a), C.byref(b))
    return int(a.value) + _py_h2d, int(b.value) + _py_d2h


def to_host(tensor):
    """CUDA tensor -> numpy array (counted device->host copy)."""
    global _py_d2h
    if tensor.is_cuda:
        _py_d2h += tensor.numel() * tensor.element_size()
    return tensor.cpu().numpy()


def ado_enumerate(bins, level_cutoff):
    """(ado_index int64 [n, bins], up int32, down int32) from the closed-form
    enumeration -- host only, no GPU needed."""
    L = lib()
    n = int(L.qsx_ado_count(bins, level_cutoff))
    if n < 0:
        raise ValueError('bad ADO enumeration arguments')
    idx = np.empty((n, bins), dtype=np.int64)
    up = np.empty((n, bins), dtype=np.int32)
    down = np.empty((n, bins), dtype=np.int32)
    check(L.qsx_ado_enumerate(bins, level_cutoff, idx.ctypes.data,
                              up.ctypes.data, down.ctypes.data))
    return idx, up, down


# ----------------------------------------------------------------- torch glue
def torch_cuda():
    """torch, after checking that a CUDA device is usable (fail loudly)."""
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError('qspectra_b200 needs a CUDA device (B200); there is '
                           'no CPU fallback')
    return torch


def current_stream_ptr():
    torch = torch_cuda()
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def to_device(array, dtype=None):
    """numpy / torch -> contiguous CUDA tensor (complex128 by default)."""
    global _py_h2d
    torch = torch_cuda()
    if isinstance(array, torch.Tensor):
        t = array
        if dtype is not None and t.dtype != dtype:
            t = t.to(dtype)
        if not t.is_cuda:
            _py_h2d += t.numel() * t.element_size()
            t = t.contiguous().pin_memory().cuda(non_blocking=True)
        return t.contiguous()
    a = np.ascontiguousarray(array, dtype=np.complex128 if dtype is None else None)
    t = torch.from_numpy(a)
    if dtype is not None:
        t = t.to(dtype)
    _py_h2d += t.numel() * t.element_size()
    # pinned staging (torch's caching host allocator): a large copy becomes a real DMA transfer,
    # and a small one does not make the host wait for the kernels queued ahead of it
    return t.pin_memory().cuda(non_blocking=True)


def int32_ptr(values):
    if values is None:
        return None, None
    arr = np.ascontiguousarray(values, dtype=np.int32)
    return arr, arr.ctypes.data_as(C.POINTER(C.c_int32))


def sample_gauss_device(seed, member0, n_members, n_gauss, scale=1.0):
    """CUDA tensor [n_members, n_gauss] = scale * RandomState(list(seed) + [n]).randn(n_gauss)
    for n = member0 .. member0 + n_members - 1, generated on the device."""
    torch = torch_cuda()
    prefix = np.ascontiguousarray(np.atleast_1d(seed), dtype=np.int64)
    if ((prefix < 0) | (prefix > 2 ** 32 - 1)).any():
        raise ValueError('seed entries must fit in uint32')
    prefix = prefix.astype(np.uint32)
    out = torch.empty((n_members, n_gauss), dtype=torch.float64, device='cuda')
    check(lib().qsx_sample_gauss_device(
        prefix.ctypes.data_as(C.POINTER(C.c_uint32)), prefix.size, int(member0),
        int(n_members), int(n_gauss), float(scale), out.data_ptr(), current_stream_ptr()))
    return out


def sample_streams(seed, member0, n_members, n_gauss, n_uniform=0):
    """(gauss [n_members, n_gauss], uniform [n_members, n_uniform]): exactly the
    draws of RandomState(list(seed) + [n]).randn(n_gauss) then .rand(n_uniform)."""
    prefix = np.ascontiguousarray(np.atleast_1d(seed), dtype=np.int64)
    if ((prefix < 0) | (prefix > 2 ** 32 - 1)).any():
        raise ValueError('seed entries must fit in uint32')
    prefix = prefix.astype(np.uint32)
    gauss = np.empty((n_members, n_gauss), dtype=np.float64)
    uni = np.empty((n_members, n_uniform), dtype=np.float64)
    check(lib().qsx_sample_streams(
        prefix.ctypes.data_as(C.POINTER(C.c_uint32)), prefix.size, int(member0),
        int(n_members), int(n_gauss), int(n_uniform), gauss.ctypes.data,
        uni.ctypes.data))
    return gauss, uni
