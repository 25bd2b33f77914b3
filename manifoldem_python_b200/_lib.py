"""ctypes binding of include/manifoldem_b200.h.

There is no CPU fallback: if the shared library is missing or a CUDA call fails, a
RuntimeError is raised (the product path must fail loudly, never route around the GPU)."""
import ctypes as C
import os
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libmanifoldem_b200.so')

SYMBOLS = [
    'mem_version', 'mem_last_error', 'mem_ctx_create', 'mem_ctx_destroy', 'mem_ctx_sync', 'mem_ctx_set_option', 'mem_ctx_launch_count',
    'mem_ctx_timer_start', 'mem_ctx_timer_stop', 'mem_ctx_kernel_time', 'mem_ctx_kernel_clock', 'mem_host_alloc', 'mem_host_free', 'mem_gather_rows_host', 'mem_dev_alloc', 'mem_dev_free', 'mem_copy_h2d', 'mem_copy_d2h',
    'mem_pd_distance_device', 'mem_pd_distance_batch_device', 'mem_pd_distance_host', 'mem_pd_last_timings', 'mem_contract_device',
    'mem_contract_knn_device', 'mem_knn_mode',
    'mem_operand_shape', 'mem_knn_device', 'mem_knn_device_f32', 'mem_graph_dense_device', 'mem_graph_compact_device', 'mem_ferguson_device',
    'mem_laplacian_dense_device', 'mem_symv_host', 'mem_nlsa_spectra_device', 'mem_nlsa_cond_device', 'mem_nlsa_supervectors_device', 'mem_nlsa_gram_small_device',
    'mem_nlsa_project_device', 'mem_nlsa_reconstruct_device', 'mem_manifold_fit_host', 'mem_lanczos_steps_device', 'mem_lanczos_ritz_device', 'mem_s2_assign_host', 'mem_s2_pairwise_host', 'mem_ctf_host',
    'mem_gather_square_device',
]


class PdParams(C.Structure):
    _fields_ = [('nS', C.c_int32), ('N', C.c_int32), ('transposed', C.c_int32), ('relion_shift', C.c_int32),
                ('filter_type', C.c_int32), ('filter_order', C.c_int32), ('filter_Qc', C.c_double),
                ('pix_size', C.c_double), ('Cs', C.c_double), ('EkV', C.c_double), ('gaussEnv', C.c_double),
                ('AmpContrast', C.c_double), ('psi_p_deg', C.c_double), ('avg_only', C.c_int32),
                ('contraction', C.c_int32), ('k_chunk_blocks', C.c_int32), ('split_k', C.c_int32),
                ('knn_k', C.c_int32), ('reserved0', C.c_int32)]


class PdIO(C.Structure):
    _fields_ = [('raw', C.c_void_p), ('flip', C.c_void_p), ('shift', C.c_void_p), ('psi_deg', C.c_void_p),
                ('df', C.c_void_p), ('msk2', C.c_void_p), ('D', C.c_void_p), ('imgAll', C.c_void_p),
                ('imgAllFlip', C.c_void_p), ('CTF', C.c_void_p), ('imgAvg', C.c_void_p), ('imgAvgFlip', C.c_void_p),
                ('imgAllIntensity', C.c_void_p), ('knn_idx', C.c_void_p), ('knn_val', C.c_void_p)]


class ContractShape(C.Structure):
    _fields_ = [('nS', C.c_int32), ('n1_blocks', C.c_int32), ('n3_blocks', C.c_int32), ('ldz', C.c_int64)]


_lib = None
_lock = threading.RLock()


def load():
    """dlopen the library and check every symbol the header declares. Needs no GPU."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError('%s is missing: run `python __graft_entry__.py` (nvcc, sm_100a) first; '
                               'there is no CPU fallback for this path' % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        missing = [s for s in SYMBOLS if not hasattr(lib, s)]
        if missing:
            raise RuntimeError('libmanifoldem_b200.so lacks symbols: %s' % missing)
        lib.mem_last_error.restype = C.c_char_p
        lib.mem_ctx_launch_count.restype = C.c_int64
        lib.mem_ctx_launch_count.argtypes = [C.c_void_p, C.c_int]
        lib.mem_ctx_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
        lib.mem_ctx_destroy.argtypes = [C.c_void_p]
        lib.mem_ctx_sync.argtypes = [C.c_void_p]
        lib.mem_ctx_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_int32]
        lib.mem_ctx_timer_start.argtypes = [C.c_void_p]
        lib.mem_ctx_timer_stop.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
        lib.mem_ctx_kernel_time.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int64),
                                            C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        lib.mem_ctx_kernel_clock.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        lib.mem_host_alloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]
        lib.mem_host_free.argtypes = [C.c_void_p]
        lib.mem_gather_rows_host.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_size_t, C.c_int32]
        lib.mem_dev_alloc.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.c_size_t]
        lib.mem_dev_free.argtypes = [C.c_void_p, C.c_void_p]
        lib.mem_copy_h2d.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
        lib.mem_copy_d2h.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
        lib.mem_pd_distance_device.argtypes = [C.c_void_p, C.POINTER(PdParams), C.POINTER(PdIO), C.c_void_p]
        lib.mem_pd_distance_host.argtypes = [C.c_void_p, C.POINTER(PdParams), C.POINTER(PdIO)]
        lib.mem_pd_last_timings.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.c_int]
        lib.mem_contract_device.argtypes = [C.c_void_p, C.POINTER(ContractShape), C.c_void_p, C.c_void_p, C.c_void_p,
                                            C.c_int32, C.c_int32, C.c_int32, C.c_void_p]
        lib.mem_contract_knn_device.argtypes = [C.c_void_p, C.POINTER(ContractShape), C.c_void_p, C.c_void_p, C.c_void_p,
                                                C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                                C.c_void_p]
        lib.mem_knn_mode.argtypes = [C.c_int32]
        lib.mem_operand_shape.argtypes = [C.c_void_p, C.c_int32, C.POINTER(ContractShape)]
        lib.mem_knn_device.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.mem_knn_device_f32.argtypes = lib.mem_knn_device.argtypes
        lib.mem_graph_compact_device.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.POINTER(C.c_int64)]
        lib.mem_graph_dense_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p,
                                               C.c_void_p]
        lib.mem_ferguson_device.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_double,
                                            C.c_void_p]
        lib.mem_laplacian_dense_device.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_double, C.c_void_p, C.c_void_p]
        lib.mem_symv_host.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
        lib.mem_gather_square_device.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int32,
                                                 C.c_void_p, C.c_void_p]
        lib.mem_ctf_host.argtypes = [C.c_void_p, C.POINTER(PdParams), C.c_void_p, C.c_void_p]
        lib.mem_s2_assign_host.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p]
        lib.mem_lanczos_steps_device.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32,
                                                 C.c_int32, C.c_void_p]
        lib.mem_lanczos_ritz_device.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p,
                                                C.c_void_p]
        vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
        lib.mem_nlsa_spectra_device.argtypes = [vp, vp, vp, i32, i32, vp, vp, vp]
        lib.mem_nlsa_cond_device.argtypes = [vp, vp, i32, i32, vp, i32, i32, vp, vp]
        lib.mem_nlsa_supervectors_device.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, i32, vp, vp, vp]
        lib.mem_nlsa_gram_small_device.argtypes = [vp, vp, i64, i32, vp, vp]
        lib.mem_nlsa_project_device.argtypes = [vp, vp, i64, i32, vp, vp, i32, i32, vp, vp]
        lib.mem_nlsa_reconstruct_device.argtypes = [vp, vp, i32, i32, i32, vp, i32, i32, vp, vp, vp]
        lib.mem_manifold_fit_host.argtypes = [vp, vp, i32, vp, vp, i32, C.c_double, C.c_double, vp]
        lib.mem_pd_distance_batch_device.argtypes = [vp, C.POINTER(PdParams), C.POINTER(PdIO), i32, vp, vp, vp]
        lib.mem_s2_pairwise_host.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
        _lib = lib
        return lib


def check(rc):
    if rc != 0:
        raise RuntimeError('manifoldem_b200: ' + (load().mem_last_error() or b'?').decode())


class PinnedArray:
    """NumPy view over cudaMallocHost memory (host side of the host-buffer entry points)."""

    def __init__(self, shape, dtype):
        self.lib = load()
        self.shape = tuple(int(s) for s in np.atleast_1d(shape))
        self.dtype = np.dtype(dtype)
        self.nbytes = int(np.prod(self.shape)) * self.dtype.itemsize
        p = C.c_void_p()
        check(self.lib.mem_host_alloc(C.byref(p), max(self.nbytes, 1)))
        self.ptr = p.value
        buf = (C.c_uint8 * max(self.nbytes, 1)).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=self.dtype, count=int(np.prod(self.shape))).reshape(self.shape)

    def free(self):
        if self.ptr:
            self.array = None
            self.lib.mem_host_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class DeviceArray:
    """Plain device allocation owned through the C ABI (tests / bench keep inputs resident with it)."""

    def __init__(self, ctx, shape, dtype, src=None):
        self.lib = load()
        self.ctx = ctx
        self.shape = tuple(int(s) for s in np.atleast_1d(shape))
        self.dtype = np.dtype(dtype)
        self.nbytes = int(np.prod(self.shape)) * self.dtype.itemsize
        p = C.c_void_p()
        check(self.lib.mem_dev_alloc(ctx.handle, C.byref(p), max(self.nbytes, 1)))
        self.ptr = p.value
        if src is not None:
            self.upload(src)

    def upload(self, src):
        src = np.ascontiguousarray(src, dtype=self.dtype)
        assert src.nbytes == self.nbytes, (src.shape, self.shape)
        check(self.lib.mem_copy_h2d(self.ctx.handle, self.ptr, src.ctypes.data, self.nbytes))

    def download(self):
        out = np.empty(self.shape, dtype=self.dtype)
        check(self.lib.mem_copy_d2h(self.ctx.handle, out.ctypes.data, self.ptr, self.nbytes))
        return out

    def free(self):
        if self.ptr:
            # a context that was closed before its arrays: plain cudaFree (handle None), never a dangling context pointer
            self.lib.mem_dev_free(self.ctx.handle if self.ctx.handle else None, self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Context:
    """One CUDA context/stream/workspace per device (thread-safe to create from any host thread)."""

    def __init__(self, device=0):
        self.lib = load()
        h = C.c_void_p()
        check(self.lib.mem_ctx_create(int(device), C.byref(h)))
        self.handle = h.value
        self.device = int(device)

    def sync(self):
        check(self.lib.mem_ctx_sync(self.handle))

    def set_option(self, name, value):
        check(self.lib.mem_ctx_set_option(self.handle, name.encode(), int(value)))

    def launches(self, reset=False):
        return int(self.lib.mem_ctx_launch_count(self.handle, 1 if reset else 0))

    def timings(self):
        t = (C.c_float * 8)()
        check(self.lib.mem_pd_last_timings(self.handle, t, 8))
        keys = ['ingest_lowpass', 'align', 'fft_ctf_operands', 'flip_avg', 'contraction', 'device_total', 'h2d', 'd2h']
        return dict(zip(keys, [float(x) for x in t]))

    def timer_start(self):
        check(self.lib.mem_ctx_timer_start(self.handle))

    def timer_stop(self):
        ms = C.c_float()
        check(self.lib.mem_ctx_timer_stop(self.handle, C.byref(ms)))
        return float(ms.value)

    def kernel_time(self, reset=True):
        """(total ms, launches, work items, K blocks) of the tcgen05 contraction launches since the last reset."""
        tot, n, items, kb = C.c_double(), C.c_int64(), C.c_int32(), C.c_int32()
        check(self.lib.mem_ctx_kernel_time(self.handle, 1 if reset else 0, C.byref(tot), C.byref(n), C.byref(items),
                                           C.byref(kb)))
        return float(tot.value), int(n.value), int(items.value), int(kb.value)

    def kernel_clock(self):
        """(SM MHz, CTA-0 wall ms) of the last tcgen05 contraction launch, from the in-kernel clock probe."""
        mhz, ms = C.c_double(), C.c_double()
        check(self.lib.mem_ctx_kernel_clock(self.handle, C.byref(mhz), C.byref(ms)))
        return float(mhz.value), float(ms.value)

    def close(self):
        if self.handle:
            self.lib.mem_ctx_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_default_ctx = {}


def default_context(device=0):
    load()
    with _lock:                      # lookup and creation under one lock: two threads never create two contexts
        ctx = _default_ctx.get(device)
        if ctx is None:
            ctx = Context(device)
            _default_ctx[device] = ctx
        return ctx
