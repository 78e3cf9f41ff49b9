"""ctypes binding of libpeppan_b200.so (declared in include/peppan_b200.h)."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('PB_LIB_PATH') or os.path.join(_HERE, 'libpeppan_b200.so')      # PB_LIB_PATH: kernel-tuning aid
_lib = None


class PbError(RuntimeError):
    pass


class ScoreParams(C.Structure):
    _fields_ = [('matrix', C.c_int8 * 1024), ('nsym', C.c_int32), ('gap_open', C.c_int32),
                ('gap_extend', C.c_int32)]


class SwStats(C.Structure):
    _fields_ = [('cells', C.c_double), ('cells_reverse', C.c_double),
                ('ms_h2d', C.c_float), ('ms_forward', C.c_float), ('ms_reverse', C.c_float),
                ('ms_traceback', C.c_float), ('ms_d2h', C.c_float), ('ms_total_device', C.c_float),
                ('h2d_bytes', C.c_int64), ('d2h_bytes', C.c_int64),
                ('kernel_launches', C.c_int32), ('n_s32_pairs', C.c_int32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


def load():
    """Load the CUDA library.  Raises (never falls back) if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PbError('libpeppan_b200.so is not built (run `python -c "import __graft_entry__ as g; g.build()"` '
                      'or `make -C peppan_b200/csrc`); peppan_b200 has no CPU fallback')
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    lib.pb_init.argtypes = [C.c_int, C.c_int, C.c_int, vp, C.POINTER(vp)]
    lib.pb_destroy.argtypes = [vp]
    lib.pb_destroy.restype = None
    lib.pb_last_error.argtypes = [vp]
    lib.pb_last_error.restype = C.c_char_p
    lib.pb_nccl_unique_id.argtypes = [vp]
    lib.pb_device_info.argtypes = [vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(i64)]
    lib.pb_reserve.argtypes = [vp, i64]
    lib.pb_reserve_sms.argtypes = [vp, i32]
    lib.pb_sw_batch.argtypes = [vp, vp, vp, vp, vp, i64, C.POINTER(ScoreParams), vp, vp, vp, vp, vp,
                                C.POINTER(SwStats)]
    lib.pb_sw_align_batch.argtypes = [vp, vp, vp, vp, vp, i64, C.POINTER(ScoreParams), vp, vp, vp, vp, vp, vp, vp,
                                      C.POINTER(C.POINTER(C.c_uint32)), C.POINTER(SwStats)]
    lib.pb_sw_job_create.argtypes = [vp, vp, vp, vp, vp, i64, C.POINTER(ScoreParams), C.c_int, C.POINTER(vp)]
    lib.pb_sw_job_run.argtypes = [vp, vp, C.POINTER(SwStats)]
    lib.pb_sw_job_fetch.argtypes = [vp, vp, vp, vp, vp, vp, vp]
    lib.pb_sw_job_destroy.argtypes = [vp, vp]
    lib.pb_sw_job_destroy.restype = None
    lib.pb_measure_dpx_peak.argtypes = [vp, C.c_int, C.POINTER(C.c_double)]
    lib.pb_free.argtypes = [vp]
    lib.pb_free.restype = None
    lib.pb_host_alloc.argtypes = [vp, i64, C.POINTER(vp)]
    lib.pb_host_free.argtypes = [vp, vp]
    lib.pb_host_free.restype = None
    _lib = lib
    return lib


def ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Context(object):
    """One per (process, GPU).  Create after fork, never before (CUDA + fork)."""

    def __init__(self, device=0, rank=0, world=1, nccl_uid=None):
        self.lib = load()
        self._pinned = []
        self.h = C.c_void_p()
        uid = None
        if nccl_uid is not None:
            uid = C.create_string_buffer(bytes(nccl_uid), 128)
        rc = self.lib.pb_init(device, rank, world, uid, C.byref(self.h))
        if rc != 0:
            raise PbError('pb_init failed (%d): %s' % (rc, self.lib.pb_last_error(None).decode()))
        self.device, self.rank, self.world = device, rank, world

    def check(self, rc, what):
        if rc != 0:
            raise PbError('%s failed (%d): %s' % (what, rc, self.lib.pb_last_error(self.h).decode()))

    def device_info(self):
        sm, clk, mem = C.c_int32(), C.c_int32(), C.c_int64()
        self.check(self.lib.pb_device_info(self.h, C.byref(sm), C.byref(clk), C.byref(mem)), 'pb_device_info')
        return dict(sm_count=sm.value, clock_khz=clk.value, hbm_bytes=mem.value)

    def reserve(self, nbytes):
        """pb_reserve: let the context's device memory pool hold `nbytes` now, so that later calls do not grow it"""
        self.check(self.lib.pb_reserve(self.h, int(nbytes)), 'pb_reserve')

    def reserve_sms(self, n):
        """pb_reserve_sms: keep `n` SMs free of this context's persistent kernels (for a communicator context on the same device)"""
        self.check(self.lib.pb_reserve_sms(self.h, int(n)), 'pb_reserve_sms')

    def dpx_peak(self, which=0):
        v = C.c_double()
        self.check(self.lib.pb_measure_dpx_peak(self.h, which, C.byref(v)), 'pb_measure_dpx_peak')
        return v.value

    def pinned_empty(self, shape, dtype):
        """numpy array backed by page-locked host memory (released with the context)."""
        dt = np.dtype(dtype)
        n = int(np.prod(shape))
        p = C.c_void_p()
        self.check(self.lib.pb_host_alloc(self.h, n * dt.itemsize, C.byref(p)), 'pb_host_alloc')
        self._pinned.append(p)
        buf = (C.c_char * max(n * dt.itemsize, 1)).from_address(p.value)
        return np.frombuffer(buf, dtype=dt, count=n).reshape(shape)

    def close(self):
        if self.h:
            for p in self._pinned:
                self.lib.pb_host_free(self.h, p)
            self._pinned = []
            self.lib.pb_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def nccl_unique_id():
    buf = C.create_string_buffer(128)
    rc = load().pb_nccl_unique_id(buf)
    if rc != 0:
        raise PbError('pb_nccl_unique_id failed: %s' % load().pb_last_error(None).decode())
    return buf.raw
