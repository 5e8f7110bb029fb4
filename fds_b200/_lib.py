"""ctypes binding of libfds_b200.so (C ABI declared in include/fds_b200.h).

There is NO CPU fallback: if the library is missing or no CUDA device is present,
every device operation raises.  Build with ``python -m fds_b200.build``.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libfds_b200.so')

c_double_p = ctypes.POINTER(ctypes.c_double)
c_int64_p = ctypes.POINTER(ctypes.c_int64)

FDS_SYS_L96_RK4, FDS_SYS_L96_EULER, FDS_SYS_LORENZ63, FDS_SYS_VANDERPOL, FDS_SYS_CIRCULAR, FDS_SYS_HOST = range(6)
FDS_MATH_EXACT, FDS_MATH_FAST = 0, 1
FDS_FLAG_SIMPLE_KERNELS = 1
FDS_FLAG_EXTERNAL_STREAM = 2
FDS_BUILD_ENSEMBLE, FDS_BUILD_KEEP_TANGENTS = 1, 2
FDS_DILATION_OFF, FDS_DILATION_FD, FDS_DILATION_ANALYTIC, FDS_DILATION_GIVEN = range(4)
FDS_BUF_ENSEMBLE, FDS_BUF_PRIMAL, FDS_BUF_GRAM, FDS_BUF_JSUM, FDS_BUF_TANGENT, FDS_BUF_DXDT = range(6)


class FdsConfig(ctypes.Structure):
    """Mirror of ``struct fds_config`` (include/fds_b200.h)."""
    _fields_ = [('device', ctypes.c_int32), ('system', ctypes.c_int32), ('math', ctypes.c_int32),
                ('m', ctypes.c_int32), ('n_local', ctypes.c_int64), ('n_global', ctypes.c_int64),
                ('site_offset', ctypes.c_int64), ('rank', ctypes.c_int32), ('world', ctypes.c_int32),
                ('max_halo_steps', ctypes.c_int32), ('flags', ctypes.c_int32), ('dt', ctypes.c_double),
                ('stream', ctypes.c_void_p), ('batch', ctypes.c_int32), ('reserved', ctypes.c_int32)]


# name -> (restype, argtypes); every symbol the header declares
_CTX = ctypes.c_void_p
_I64, _I32, _D = ctypes.c_int64, ctypes.c_int, ctypes.c_double
SIGNATURES = {
    'fds_ctx_create': (_I32, [ctypes.POINTER(FdsConfig), ctypes.POINTER(_CTX)]),
    'fds_ctx_destroy': (None, [_CTX]),
    'fds_last_error': (ctypes.c_char_p, [_CTX]),
    'fds_version': (ctypes.c_char_p, []),
    'fds_set_state': (_I32, [_CTX, c_double_p]),
    'fds_get_state': (_I32, [_CTX, c_double_p]),
    'fds_set_tangents': (_I32, [_CTX, c_double_p, c_double_p]),
    'fds_get_tangents': (_I32, [_CTX, c_double_p, c_double_p]),
    'fds_host_alloc': (ctypes.c_void_p, [ctypes.c_size_t]),
    'fds_host_free': (None, [ctypes.c_void_p]),
    'fds_random_tangents': (_I32, [_CTX, ctypes.c_uint64]),
    'fds_get_ensemble': (_I32, [_CTX, c_double_p]),
    'fds_set_ensemble': (_I32, [_CTX, c_double_p, _D]),
    'fds_set_primal_history': (_I32, [_CTX, _I32, c_double_p]),
    'fds_set_dxdt_weights': (_I32, [_CTX, c_double_p]),
    'fds_set_dxdt_stencil': (_I32, [_CTX, _I32, c_double_p]),
    'fds_dilation_order': (_I32, [_CTX]),
    'fds_set_time_dilation': (_I32, [_CTX, _I32]),
    'fds_set_dxdt': (_I32, [_CTX, c_double_p]),
    'fds_set_qr_passes': (_I32, [_CTX, _I32]),
    'fds_last_condition': (_D, [_CTX]),
    'fds_single_pass_pending': (_I32, [_CTX]),
    'fds_last_qr_passes': (_I32, [_CTX]),
    'fds_run': (_I32, [_CTX, _D, _I64, c_double_p]),
    'fds_primal_advance': (_I32, [_CTX, _D, _I64, _I64, _I32]),
    'fds_halo_mark_valid': (_I32, [_CTX, _I32, _I64]),
    'fds_halo_wrap_ensemble': (_I32, [_CTX, _I64]),
    'fds_halo_wrap_primal': (_I32, [_CTX, _I64]),
    'fds_halo_layout': (_I32, [_CTX, _I32, _I64, c_int64_p, c_int64_p]),
    'fds_boundary_begin': (_I32, [_CTX, _I32]),
    'fds_boundary_dxdt': (_I32, [_CTX, _D, _I32, _D]),
    'fds_boundary_factor': (_I32, [_CTX, _I32, _I32, _D]),
    'fds_boundary_finish': (_I32, [_CTX, _I32, _D, _I32]),
    'fds_boundary_results': (_I32, [_CTX, c_double_p, c_double_p, c_double_p, c_double_p]),
    'fds_halo_steps': (_I64, [_CTX, _I64]),
    'fds_segment_advance': (_I32, [_CTX, _D, _D, _I64, _I64]),
    'fds_segment_jsums': (_I32, [_CTX, _I64, c_double_p]),
    'fds_boundary': (_I32, [_CTX, _D, _D, _I32, _I32, _I32, _I64, c_double_p, c_double_p, c_double_p, c_double_p]),
    'fds_segment': (_I32, [_CTX, _D, _D, _I64, c_double_p]),
    'fds_run_segment_host': (_I32, [_CTX, c_double_p, c_double_p, c_double_p, _D, _D, _I64, c_double_p]),
    'fds_lss_solve': (_I32, [_I32, c_double_p, c_double_p, _I64, _I32, c_double_p]),
    'fds_lss_solve_batch': (_I32, [_I32, c_double_p, c_double_p, _I64, _I32, _I64, c_double_p]),
    'fds_set_batch_params': (_I32, [_CTX, c_double_p]),
    'fds_batch_count': (_I32, [_CTX]),
    'fds_segment_sensitivities': (_I32, [_CTX, _I64, _D, c_double_p, c_double_p]),
    'fds_boundary_out_doubles': (_I64, [_CTX]),
    'fds_boundary_enqueue': (_I32, [_CTX, _D, _D, _I32, _I32, _I32, _I64, c_double_p]),
    'fds_boundary_check': (_I32, [_CTX, c_double_p, _I32]),
    'fds_segment_enqueue': (_I32, [_CTX, _D, _D, _I64, c_double_p]),
    'fds_comm_unique_id': (_I32, [ctypes.c_void_p]),
    'fds_comm_init': (_I32, [_CTX, ctypes.c_void_p]),
    'fds_comm_attach': (_I32, [_CTX, ctypes.c_void_p]),
    'fds_comm_attached': (_I32, [_CTX]),
    'fds_host_alloc_wc': (ctypes.c_void_p, [ctypes.c_size_t]),
    'fds_collectives': (_I64, [_CTX]),
    'fds_get_ensemble_rows': (_I32, [_CTX, _I64, _I64, c_double_p]),
    'fds_get_ensemble_column': (_I32, [_CTX, _I32, c_double_p]),
    'fds_strip_plan': (_I32, [_CTX, c_int64_p]),
    'fds_fp64_issue_rate': (_I32, [_I32, c_double_p, c_double_p, c_double_p]),
    'fds_buffer_ptr': (ctypes.c_void_p, [_CTX, _I32]),
    'fds_kernel_launches': (_I64, [_CTX]),
    'fds_sync': (_I32, [_CTX]),
    'fds_profile_enable': (_I32, [_CTX, _I32]),
    'fds_profile_read': (_I32, [_CTX, c_double_p, c_int64_p]),
    'fds_profile_cycles': (_I32, [_CTX, c_double_p]),
    'fds_profile_steps': (_I64, [_CTX]),
}

_lib = None


class FdsError(RuntimeError):
    pass


def load():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FdsError('%s not found: build it with `python -m fds_b200.build` '
                       '(fds_b200 has no CPU fallback)' % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here = header / library mismatch
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def dptr(a):
    """double* of a C-contiguous float64 numpy array (or NULL for None)."""
    if a is None:
        return None
    assert a.dtype.name == 'float64' and a.flags['C_CONTIGUOUS']
    return a.ctypes.data_as(c_double_p)


# --------------------------------------------------------------------------- #
# Page-locked host arrays.  Large results (the (m, N) tangent block of a checkpoint)
# are returned in pinned memory so the device -> host copy runs at PCIe rate;
# blocks are recycled through a small pool because cudaHostAlloc itself is slow.
# --------------------------------------------------------------------------- #
PINNED_MIN_BYTES = 1 << 24      # smaller arrays are ordinary numpy allocations
_POOL_MAX_FREE = 2              # free blocks kept per size
_pool = {}


class _PinnedBlock:
    """Owner of one cudaHostAlloc block; exposes it through __array_interface__."""

    def __init__(self, nbytes):
        self.nbytes = int(nbytes)
        self.ptr = None
        free = _pool.get(self.nbytes)
        self.ptr = free.pop() if free else load().fds_host_alloc(self.nbytes)
        if not self.ptr:
            raise FdsError('fds_host_alloc(%d) failed' % self.nbytes)

    def view(self, shape):
        self.__array_interface__ = {'shape': tuple(int(x) for x in shape), 'typestr': '<f8',
                                    'data': (int(self.ptr), False), 'version': 3}
        import numpy as np
        return np.asarray(self)

    def __del__(self):
        try:
            if not self.ptr:
                return
            free = _pool.setdefault(self.nbytes, [])
            if len(free) < _POOL_MAX_FREE:
                free.append(self.ptr)
            elif _lib is not None:
                _lib.fds_host_free(self.ptr)
        except Exception:
            pass


def pinned_empty(shape):
    """float64 numpy array in page-locked host memory (kept alive by the array's base)."""
    import numpy as np
    shape = (shape,) if np.isscalar(shape) else tuple(shape)
    n = 1
    for x in shape:
        n *= int(x)
    return _PinnedBlock(max(n, 1) * 8).view(shape)


def host_empty(shape):
    """Result buffer: pinned when large, plain numpy otherwise."""
    import numpy as np
    n = int(np.prod(shape)) * 8
    return pinned_empty(shape) if n >= PINNED_MIN_BYTES else np.empty(shape)


def pool_release():
    """Free every pooled pinned block."""
    for free in _pool.values():
        while free:
            load().fds_host_free(free.pop())
