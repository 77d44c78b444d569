"""ctypes mirror of include/gravity_cuda.h (the C ABI of libgravity_cuda).

Names, argument meaning and error behaviour follow the header one to one; a
negative return code becomes GravityError carrying gravity_cuda_last_error().
There is no fallback of any kind: if the shared library is missing this module
raises at first use, and on a box without an sm_100 GPU `GravityCuda(...)`
raises with the library's own message.
"""
import ctypes as C
import os

import numpy as np

KERNEL_AUTO, KERNEL_TILED, KERNEL_STRICT = 0, 1, 2
NCCL_ID_BYTES = 128
P2P_BLOB_BYTES = 256

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_dp = C.POINTER(C.c_double)


class GravityError(RuntimeError):
    pass


def _preload_bundled_nccl():
    """libgravity_cuda links libnccl.so.2.  A process can hold only one library of that soname: if
    the system's NCCL gets in first, a later `import torch` fails, because torch's CUDA library needs
    symbols of the newer NCCL that ships with it (site-packages/nvidia/nccl).  So when that bundled
    NCCL exists it is loaded first -- the C ABI of libnccl.so.2 is backward compatible, and it is the
    one in use anyway whenever torch was imported before this module (bench.py, the rank workers)."""
    import importlib.util
    try:
        spec = importlib.util.find_spec("nvidia.nccl")
    except (ImportError, ValueError):
        spec = None
    for base in (list(spec.submodule_search_locations) if spec and spec.submodule_search_locations else []):
        cand = os.path.join(base, "lib", "libnccl.so.2")
        if os.path.exists(cand):
            try:
                C.CDLL(cand, mode=C.RTLD_GLOBAL)
            except OSError:
                pass
            return


def load_cuda_library():
    """Load proj_entropy_b200/libgravity_cuda.so and declare every prototype."""
    global _LIB
    if _LIB is not None:
        return _LIB
    # GRAVITY_CUDA_LIB lets developer tools A/B a differently compiled build
    path = os.environ.get("GRAVITY_CUDA_LIB") or os.path.join(_HERE, "libgravity_cuda.so")
    if not os.path.exists(path):
        raise GravityError("%s is not built (run `make` or __graft_entry__.build()); "
                           "there is no fallback implementation" % path)
    _preload_bundled_nccl()
    lib = C.CDLL(path, mode=C.RTLD_GLOBAL)
    H = C.c_void_p
    sigs = {
        "gravity_cuda_device_count": (C.c_int, []),
        "gravity_cuda_last_error": (C.c_char_p, [H]),
        "gravity_cuda_nccl_unique_id": (C.c_int, [C.c_void_p, C.c_size_t]),
        "gravity_cuda_create": (C.c_int, [C.POINTER(H), C.c_int64, C.c_int]),
        "gravity_cuda_create_rank": (C.c_int, [C.POINTER(H), C.c_int64, C.c_int, C.c_int, C.c_int,
                                               C.c_void_p, C.c_size_t]),
        "gravity_cuda_p2p_export": (C.c_int, [H, C.c_void_p, C.c_size_t]),
        "gravity_cuda_p2p_connect": (C.c_int, [H, C.c_void_p, C.c_size_t]),
        "gravity_cuda_destroy": (None, [H]),
        "gravity_cuda_set_option": (C.c_int, [H, C.c_char_p, C.c_int64]),
        "gravity_cuda_upload": (C.c_int, [H, _dp, _dp, _dp, _dp, _dp]),
        "gravity_cuda_download": (C.c_int, [H, _dp, _dp, _dp, _dp]),
        "gravity_cuda_step": (C.c_int, [H, C.c_int32, C.c_int64]),
        "gravity_cuda_step_async": (C.c_int, [H, C.c_int32, C.c_int64]),
        "gravity_cuda_sync": (C.c_int, [H]),
        "gravity_cuda_step_host": (C.c_int, [H, _dp, _dp, _dp, _dp, _dp, C.c_int32, C.c_int64]),
        "gravity_cuda_accel": (C.c_int, [H, C.c_int32, _dp, _dp]),
        "gravity_cuda_step_objects": (C.c_int, [H, C.POINTER(C.c_void_p), C.c_int64, C.c_size_t, C.c_size_t,
                                                C.c_size_t, C.c_size_t, C.c_size_t, C.c_int32, C.c_int64]),
        "gravity_cuda_path_enable": (C.c_int, [H, C.c_int64]),
        "gravity_cuda_step_path": (C.c_int, [H, C.c_int32, C.c_int64, C.c_int64, C.POINTER(C.c_int32),
                                             C.POINTER(C.c_int64)]),
        "gravity_cuda_path_read": (C.c_int, [H, C.c_int64, C.c_int64, _dp, _dp]),
        "gravity_cuda_path_samples_in": (C.c_int64, [C.c_int32, C.c_int64, C.c_int64, C.POINTER(C.c_int32)]),
        "gravity_cuda_trace_read": (C.c_int, [H, C.POINTER(C.c_uint64), C.c_int64, C.POINTER(C.c_int64)]),
        "gravity_cuda_energy": (C.c_int, [H, _dp, _dp]),
        "gravity_cuda_timer_start": (C.c_int, [H]),
        "gravity_cuda_timer_stop": (C.c_int, [H, _dp]),
        "gravity_cuda_launch_count": (C.c_int, [H, C.POINTER(C.c_int64)]),
        "gravity_cuda_kernel_time": (C.c_int, [H, _dp, C.POINTER(C.c_int64)]),
        "gravity_cuda_kernel_times": (C.c_int, [H, _dp, C.c_int64, C.POINTER(C.c_int64)]),
        "gravity_cuda_plan_probe": (C.c_int, [C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                              C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_int64]),
        "gravity_cuda_fp64_peak": (C.c_int, [C.c_int, _dp, _dp]),
    }
    for name, (res, args) in sigs.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    lib._declared = sorted(sigs)
    _LIB = lib
    return lib


def _f64(a, n=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if n is not None and a.shape != (n,):
        raise ValueError("expected %d doubles, got shape %s" % (n, a.shape))
    return a


def _ptr(a):
    return a.ctypes.data_as(_dp)


def device_count():
    return load_cuda_library().gravity_cuda_device_count()


def nccl_unique_id():
    lib = load_cuda_library()
    buf = (C.c_ubyte * NCCL_ID_BYTES)()
    if lib.gravity_cuda_nccl_unique_id(buf, NCCL_ID_BYTES) != 0:
        raise GravityError(lib.gravity_cuda_last_error(None).decode())
    return bytes(buf)


def fp64_peak(device=0):
    """(TFLOP/s of the FP64 FMA pipe measured by a DFMA microbenchmark, SM MHz)."""
    lib = load_cuda_library()
    tf, mhz = C.c_double(), C.c_double()
    if lib.gravity_cuda_fp64_peak(device, C.byref(tf), C.byref(mhz)) != 0:
        raise GravityError(lib.gravity_cuda_last_error(None).decode())
    return tf.value, mhz.value


def path_samples_in(dt, nsteps, sim_time=0, last_day=0):
    """(PATH samples that nsteps steps of size dt take from sim_time, last_day afterwards): the
    commit block's day test as the kernels run it, computed on the host (no GPU needed)."""
    lib = load_cuda_library()
    ld = C.c_int32(int(last_day))
    cnt = lib.gravity_cuda_path_samples_in(int(dt), int(nsteps), int(sim_time), C.byref(ld))
    if cnt < 0:
        raise GravityError("path_samples_in: dt >= 1, sim_time >= 0 and nsteps >= 0 are required")
    return cnt, ld.value


def plan_probe(n, nranks=1, rank=0, num_sms=148, threads=0, bodies_per_thread=0, ctas_per_sm=0):
    """The tiled kernel's launch plan, computed on the host (no GPU needed).
    Returns (info dict, segments array [n_segments, 5] = cta, iblock, tile_begin, tile_end, partial_row)."""
    lib = load_cuda_library()
    info = np.zeros(10, dtype=np.int32)
    ip = C.POINTER(C.c_int32)
    if lib.gravity_cuda_plan_probe(n, nranks, rank, num_sms, threads, bodies_per_thread, ctas_per_sm,
                                   info.ctypes.data_as(ip), None, 0) != 0:
        raise GravityError(lib.gravity_cuda_last_error(None).decode())
    segs = np.zeros((max(1, int(info[4])), 5), dtype=np.int32)
    lib.gravity_cuda_plan_probe(n, nranks, rank, num_sms, threads, bodies_per_thread, ctas_per_sm,
                                info.ctypes.data_as(ip), segs.ctypes.data_as(ip), int(info[4]))
    keys = ["threads", "bodies_per_thread", "ctas_per_sm", "n_cta", "n_segments", "n_iblock", "n_tile",
            "local_tile_begin", "local_tile_end", "n_local"]
    return dict(zip(keys, (int(v) for v in info))), segs[:int(info[4])]


class GravityCuda:
    """One gravity_cuda_t handle.

    GravityCuda(n, ngpus=k)                         -> gravity_cuda_create
    GravityCuda(n, rank=r, nranks=p, device=d, nccl_id=b) -> gravity_cuda_create_rank
    """

    def __init__(self, n, ngpus=1, rank=None, nranks=None, device=0, nccl_id=None, **options):
        self._lib = load_cuda_library()
        self._h = C.c_void_p()
        self.n = int(n)
        if rank is None:
            rc = self._lib.gravity_cuda_create(C.byref(self._h), self.n, int(ngpus))
        else:
            idbuf = None
            if nccl_id is not None:
                idbuf = (C.c_ubyte * NCCL_ID_BYTES).from_buffer_copy(nccl_id)
            rc = self._lib.gravity_cuda_create_rank(C.byref(self._h), self.n, int(rank), int(nranks), int(device),
                                                    idbuf, NCCL_ID_BYTES if idbuf is not None else 0)
        if rc != 0:
            self._h = C.c_void_p()
            raise GravityError(self._lib.gravity_cuda_last_error(None).decode())
        for k, v in options.items():
            self.set_option(k, v)

    # -- plumbing -----------------------------------------------------------
    def _check(self, rc):
        if rc != 0:
            raise GravityError(self._lib.gravity_cuda_last_error(self._h).decode())

    def close(self):
        if self._h:
            self._lib.gravity_cuda_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- the C ABI, one method per entry point --------------------------------
    def set_option(self, key, value):
        self._check(self._lib.gravity_cuda_set_option(self._h, key.encode(), int(value)))

    def p2p_export(self):
        buf = (C.c_ubyte * P2P_BLOB_BYTES)()
        self._check(self._lib.gravity_cuda_p2p_export(self._h, buf, P2P_BLOB_BYTES))
        return bytes(buf)

    def p2p_connect(self, blobs=None):
        """blobs: the P2P_BLOB_BYTES blobs of all ranks concatenated in rank order
        (rank mode), or None for a single-process handle."""
        if blobs is None:
            self._check(self._lib.gravity_cuda_p2p_connect(self._h, None, 0))
        else:
            buf = (C.c_ubyte * len(blobs)).from_buffer_copy(blobs)
            self._check(self._lib.gravity_cuda_p2p_connect(self._h, buf, len(blobs)))

    def upload(self, mass, x, y, vx, vy):
        """mass may be None after the first upload (the masses stay)."""
        n = self.n
        arrs = [None if a is None else _f64(a, n) for a in (mass, x, y, vx, vy)]
        self._check(self._lib.gravity_cuda_upload(self._h, *[None if a is None else _ptr(a) for a in arrs]))

    def upload_ptrs(self, mass, x, y, vx, vy):
        """Raw host addresses (e.g. pinned torch tensors' data_ptr()); mass may be 0/None."""
        cast = lambda p: C.cast(C.c_void_p(int(p)), _dp) if p else None
        self._check(self._lib.gravity_cuda_upload(self._h, cast(mass), cast(x), cast(y), cast(vx), cast(vy)))

    def download_ptrs(self, x, y, vx, vy):
        cast = lambda p: C.cast(C.c_void_p(int(p)), _dp) if p else None
        self._check(self._lib.gravity_cuda_download(self._h, cast(x), cast(y), cast(vx), cast(vy)))

    def download(self):
        # zeros, not empty: with download_scope=1 only the caller's own shard is written
        out = [np.zeros(self.n, dtype=np.float64) for _ in range(4)]
        self._check(self._lib.gravity_cuda_download(self._h, *[_ptr(a) for a in out]))
        return tuple(out)

    def step(self, dt, nsteps=1):
        self._check(self._lib.gravity_cuda_step(self._h, int(dt), int(nsteps)))

    def step_async(self, dt, nsteps=1):
        self._check(self._lib.gravity_cuda_step_async(self._h, int(dt), int(nsteps)))

    def sync(self):
        self._check(self._lib.gravity_cuda_sync(self._h))

    def path_enable(self, capacity):
        self._check(self._lib.gravity_cuda_path_enable(self._h, int(capacity)))

    def step_path(self, dt, nsteps, sim_time, last_day):
        """-> (last_day after the batch, samples taken since path_enable)"""
        ld, tail = C.c_int32(int(last_day)), C.c_int64()
        self._check(self._lib.gravity_cuda_step_path(self._h, int(dt), int(nsteps), int(sim_time), C.byref(ld),
                                                     C.byref(tail)))
        return ld.value, tail.value

    def path_read(self, first, count):
        """-> (x, y), each [count, n]: PATH samples first..first+count-1 of every body"""
        x = np.zeros((int(count), self.n), dtype=np.float64)
        y = np.zeros((int(count), self.n), dtype=np.float64)
        self._check(self._lib.gravity_cuda_path_read(self._h, int(first), int(count), _ptr(x), _ptr(y)))
        return x, y

    def trace_read(self, cap=2048):
        """-> uint64 [n_cta, 8] globaltimer stamps of the last tiled launch (option trace=1)"""
        buf = np.zeros((cap, 8), dtype=np.uint64)
        n = C.c_int64()
        self._check(self._lib.gravity_cuda_trace_read(self._h, buf.ctypes.data_as(C.POINTER(C.c_uint64)), cap,
                                                      C.byref(n)))
        return buf[:min(cap, n.value)]

    def step_host(self, mass, x, y, vx, vy, dt, nsteps=1):
        """upload + nsteps + download in one call; x, y, vx, vy (contiguous float64 arrays) are
        updated in place, mass may be None after the first upload."""
        for a in (x, y, vx, vy):
            if not (isinstance(a, np.ndarray) and a.dtype == np.float64 and a.flags.c_contiguous and a.shape == (self.n,)):
                raise ValueError("step_host updates contiguous float64 arrays of length n in place")
        m = None if mass is None else _ptr(_f64(mass, self.n))
        self._check(self._lib.gravity_cuda_step_host(self._h, m, _ptr(x), _ptr(y), _ptr(vx), _ptr(vy), int(dt), int(nsteps)))

    def step_host_ptrs(self, mass, x, y, vx, vy, dt, nsteps=1):
        """Raw host addresses (e.g. pinned torch tensors' data_ptr()); mass may be 0/None."""
        cast = lambda p: C.cast(C.c_void_p(int(p)), _dp) if p else None
        self._check(self._lib.gravity_cuda_step_host(self._h, cast(mass), cast(x), cast(y), cast(vx), cast(vy),
                                                     int(dt), int(nsteps)))

    def accel(self, dt):
        ax = np.zeros(self.n, dtype=np.float64)
        ay = np.zeros(self.n, dtype=np.float64)
        self._check(self._lib.gravity_cuda_accel(self._h, int(dt), _ptr(ax), _ptr(ay)))
        return ax, ay

    def step_objects(self, object_ptrs, off_x, off_y, off_vx, off_vy, off_mass, dt, nsteps=1):
        n = len(object_ptrs)
        arr = (C.c_void_p * n)(*object_ptrs)
        self._check(self._lib.gravity_cuda_step_objects(self._h, arr, n, off_x, off_y, off_vx, off_vy, off_mass,
                                                        int(dt), int(nsteps)))

    def energy(self):
        ke, pe = C.c_double(), C.c_double()
        self._check(self._lib.gravity_cuda_energy(self._h, C.byref(ke), C.byref(pe)))
        return ke.value, pe.value

    def timer_start(self):
        self._check(self._lib.gravity_cuda_timer_start(self._h))

    def timer_stop(self):
        ms = C.c_double()
        self._check(self._lib.gravity_cuda_timer_stop(self._h, C.byref(ms)))
        return ms.value

    def launch_count(self):
        v = C.c_int64()
        self._check(self._lib.gravity_cuda_launch_count(self._h, C.byref(v)))
        return v.value

    def kernel_times(self, cap=4096):
        out = np.zeros(cap, dtype=np.float64)
        cnt = C.c_int64()
        self._check(self._lib.gravity_cuda_kernel_times(self._h, _ptr(out), cap, C.byref(cnt)))
        return out[:min(cap, cnt.value)]

    def kernel_time(self):
        ms, cnt = C.c_double(), C.c_int64()
        self._check(self._lib.gravity_cuda_kernel_time(self._h, C.byref(ms), C.byref(cnt)))
        return ms.value, cnt.value
