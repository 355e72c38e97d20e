"""ctypes binding of libchi_b200.so (include/chi_b200.h).

The library is the product: there is no Python or CPU fallback.  Importing
this module never touches the GPU; the first call that needs a device fails
loudly if the shared object is missing or no CUDA device is present.
"""
import ctypes
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# CHI_B200_LIB: an alternative build of the same library (kernel experiments)
LIB_PATH = os.environ.get('CHI_B200_LIB') or os.path.join(
    HERE, 'libchi_b200.so')
COMM_ID_BYTES = 128

c_i32p = ctypes.POINTER(ctypes.c_int32)
c_i64p = ctypes.POINTER(ctypes.c_int64)
c_f64p = ctypes.POINTER(ctypes.c_double)
c_vp = ctypes.c_void_p

ERR_KINDS = {'gaussian': 0, 'lognormal': 1, 'cam': 2, 'multiplicative': 3}
POP_KINDS = {'pooled': 0, 'heterogeneous': 1, 'gaussian': 2, 'lognormal': 3,
             'truncated_gaussian': 4}

_lib = None
_modules = {}
_default_device = None


def set_default_device(device):
    """CUDA device used by problems created from now on (one process per
    GPU: call with LOCAL_RANK)."""
    global _default_device
    _default_device = int(device)


def bind_to_device_numa_node(device=None):
    """Restricts this process to the CPUs of the NUMA node the GPU hangs off
    (one process per GPU on a multi-socket host): page-locked buffers that are
    allocated afterwards land in that node's memory (first touch), so the
    copies of eight ranks do not all cross the socket interconnect.  Returns
    the node, or None when the topology is not exposed (nothing changes
    then)."""
    device = default_device() if device is None else int(device)
    try:
        import torch
        props = torch.cuda.get_device_properties(device)
        bus = '%04x:%02x:%02x.0' % (
            props.pci_domain_id, props.pci_bus_id, props.pci_device_id)
        with open('/sys/bus/pci/devices/%s/numa_node' % bus) as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open('/sys/devices/system/node/node%d/cpulist' % node) as f:
            text = f.read().strip()
        cpus = set()
        for part in text.split(','):
            lo, _, hi = part.partition('-')
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except Exception:   # no sysfs / no such attribute: leave things alone
        return None


def default_device():
    if _default_device is not None:
        return _default_device
    return int(os.environ.get('CHI_B200_DEVICE', '0'))

_SIGNATURES = {
    'chi_b200_abi_version': (ctypes.c_int, []),
    'chi_b200_last_error': (ctypes.c_char_p, []),
    'chi_b200_device_count': (ctypes.c_int, []),
    'chi_b200_host_alloc': (
        ctypes.c_int, [ctypes.POINTER(c_vp), ctypes.c_int64]),
    'chi_b200_host_free': (ctypes.c_int, [c_vp]),
    'chi_b200_n_models': (ctypes.c_int, []),
    'chi_b200_model_find': (ctypes.c_int, [ctypes.c_char_p]),
    'chi_b200_model_info': (ctypes.c_int, [ctypes.c_int, c_i32p]),
    'chi_b200_model_key': (ctypes.c_char_p, [ctypes.c_int]),
    'chi_b200_model_inert_mask': (
        ctypes.c_int, [ctypes.c_int, ctypes.POINTER(ctypes.c_uint32)]),
    'chi_b200_model_variants': (
        ctypes.c_int, [ctypes.c_int, c_i32p, c_i32p, c_i32p, ctypes.c_int32]),
    'chi_b200_model_describe': (
        ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int, c_i32p]),
    'chi_b200_problem_create': (
        ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.POINTER(c_vp)]),
    'chi_b200_problem_destroy': (None, [c_vp]),
    'chi_b200_problem_set_outputs': (
        ctypes.c_int, [c_vp, ctypes.c_int32, c_i32p, c_i32p]),
    'chi_b200_problem_set_sensitivity_columns': (
        ctypes.c_int, [c_vp, ctypes.c_int32, c_i32p]),
    'chi_b200_problem_set_tolerances': (
        ctypes.c_int, [c_vp, ctypes.c_double, ctypes.c_double,
                       ctypes.c_int32]),
    'chi_b200_problem_set_variant': (ctypes.c_int, [c_vp, ctypes.c_int32]),
    'chi_b200_problem_set_fixed_parameters': (
        ctypes.c_int, [c_vp, ctypes.c_int32, c_i32p, c_f64p]),
    'chi_b200_problem_set_n_ids': (ctypes.c_int, [c_vp, ctypes.c_int64]),
    'chi_b200_problem_upload': (
        ctypes.c_int, [c_vp, ctypes.c_int64, c_i64p, c_f64p, c_f64p, c_i32p,
                       c_i64p, c_f64p]),
    'chi_b200_simulate': (
        ctypes.c_int, [c_vp, ctypes.c_int64, c_i32p, c_f64p, ctypes.c_int32,
                       c_f64p, c_f64p, c_i32p]),
    'chi_b200_loglik': (
        ctypes.c_int, [c_vp, ctypes.c_int64, c_i32p, c_f64p, ctypes.c_int32,
                       c_f64p, c_f64p, c_i32p, c_i32p, c_i32p]),
    'chi_b200_loglik_dev': (
        ctypes.c_int, [c_vp, ctypes.c_int64, c_vp, c_vp, ctypes.c_int32,
                       c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    'chi_b200_error_model': (
        ctypes.c_int, [ctypes.c_int32, ctypes.c_int32, c_f64p,
                       ctypes.c_int64, ctypes.c_int32, c_f64p, c_f64p,
                       c_f64p, ctypes.c_int32, c_f64p, c_f64p]),
    'chi_b200_covariate_linear': (
        ctypes.c_int, [ctypes.c_int32, ctypes.c_int64, ctypes.c_int32,
                       ctypes.c_int32, ctypes.c_int32, ctypes.c_int32,
                       c_i32p, c_i32p, c_f64p, c_f64p, c_f64p, c_f64p,
                       c_f64p, c_f64p, c_f64p]),
    'chi_b200_population_set': (
        ctypes.c_int, [c_vp, ctypes.c_int32, c_i32p, c_i32p, c_i32p, c_i32p,
                       c_i32p, c_i32p, c_i32p, c_f64p]),
    'chi_b200_population_sizes': (ctypes.c_int, [c_vp, c_i64p]),
    'chi_b200_population_eval': (
        ctypes.c_int, [c_vp, c_f64p, c_f64p, c_f64p, ctypes.c_int32, c_f64p,
                       c_f64p, c_f64p, c_f64p]),
    'chi_b200_hier_logpdf': (
        ctypes.c_int, [c_vp, ctypes.c_int32, c_f64p, ctypes.c_int32, c_f64p,
                       c_f64p, c_i32p]),
    'chi_b200_hier_logpdf_dev': (
        ctypes.c_int, [c_vp, ctypes.c_int32, c_vp, ctypes.c_int32, c_vp,
                       c_vp, c_vp, c_vp]),
    'chi_b200_problem_set_shard': (
        ctypes.c_int, [c_vp, ctypes.c_int64, ctypes.c_int64]),
    'chi_b200_comm_unique_id': (
        ctypes.c_int, [ctypes.POINTER(ctypes.c_uint8), ctypes.c_int32]),
    'chi_b200_problem_comm_init': (
        ctypes.c_int, [c_vp, ctypes.c_int32, ctypes.c_int32,
                       ctypes.POINTER(ctypes.c_uint8)]),
    'chi_b200_problem_comm_destroy': (ctypes.c_int, [c_vp]),
    'chi_b200_problem_counters': (ctypes.c_int, [c_vp, c_i64p]),
    'chi_b200_problem_set_timing': (ctypes.c_int, [c_vp, ctypes.c_int32]),
    'chi_b200_problem_set_snapshot_budget': (
        ctypes.c_int, [c_vp, ctypes.c_int64]),
    'chi_b200_problem_set_batching': (
        ctypes.c_int, [c_vp, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32]),
    'chi_b200_hier_logpdf_begin': (
        ctypes.c_int, [c_vp, ctypes.c_int32, c_f64p, ctypes.c_int32, c_f64p,
                       c_f64p, c_i32p]),
    'chi_b200_problem_synchronize': (ctypes.c_int, [c_vp]),
    'chi_b200_problem_set_small_batch': (
        ctypes.c_int, [c_vp, ctypes.c_int64]),
    'chi_b200_problem_timings': (ctypes.c_int, [c_vp, c_f64p]),
    'chi_b200_fp64_peak': (ctypes.c_int, [ctypes.c_int32, c_f64p, c_f64p]),
}


def exported_symbols():
    """Names every entry point include/chi_b200.h declares."""
    return sorted(_SIGNATURES)


def lib():
    """Loads the native library (no GPU needed for loading)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            'chi_b200 native library not found at %s. Build it with '
            '`python -m chi_b200._build` (needs nvcc); there is no CPU '
            'fallback.' % LIB_PATH)
    handle = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)
    for name, (restype, argtypes) in _SIGNATURES.items():
        fn = getattr(handle, name)
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = handle
    return _lib


def check(rc):
    if rc != 0:
        msg = lib().chi_b200_last_error().decode()
        raise RuntimeError('chi_b200: ' + msg)


def require_gpu():
    n = lib().chi_b200_device_count()
    if n <= 0:
        raise RuntimeError(
            'chi_b200 needs a CUDA device (B200, sm_100a); none is visible '
            'and there is no CPU fallback.')
    return n


def load_module(path):
    """Loads a separately compiled model module (registers on load)."""
    lib()
    if path not in _modules:
        _modules[path] = ctypes.CDLL(path, mode=ctypes.RTLD_GLOBAL)
    return _modules[path]


def find_model(code):
    """Returns the registry id of a generated model, compiling and loading
    its module on first use if it is not one of the built-in models."""
    handle = lib()
    idx = handle.chi_b200_model_find(code.key.encode())
    if idx >= 0:
        return idx
    from chi_b200 import _build
    # returns at once when a module built from the same generated code and
    # integrator headers exists
    so = _build.build_model_module(code)
    load_module(so)
    idx = handle.chi_b200_model_find(code.key.encode())
    if idx < 0:
        raise RuntimeError('Model module failed to register: ' + so)
    return idx


def f64(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(c_f64p)


def i32(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(c_i32p)


def i64(a):
    a = np.ascontiguousarray(a, dtype=np.int64)
    return a, a.ctypes.data_as(c_i64p)


def out_f64(shape):
    a = np.empty(shape, dtype=np.float64)
    return a, a.ctypes.data_as(c_f64p)


def out_i32(shape):
    a = np.zeros(shape, dtype=np.int32)
    return a, a.ctypes.data_as(c_i32p)


class _PinnedBlock(object):
    def __init__(self, nbytes):
        self.ptr = c_vp()
        check(lib().chi_b200_host_alloc(ctypes.byref(self.ptr), int(nbytes)))

    def __del__(self):
        try:
            if self.ptr.value:
                lib().chi_b200_host_free(self.ptr)
        except Exception:  # pragma: no cover
            pass


def pinned_empty(shape, dtype=np.float64):
    """numpy array backed by page-locked host memory (cudaHostAlloc), so the
    library's cudaMemcpyAsync calls are true asynchronous DMA transfers."""
    dtype = np.dtype(dtype)
    n = int(np.prod(shape))
    block = _PinnedBlock(max(n * dtype.itemsize, 8))
    buf = (ctypes.c_char * (n * dtype.itemsize)).from_address(block.ptr.value)
    arr = np.frombuffer(buf, dtype=dtype, count=n).reshape(shape)
    _pinned_owner[id(arr)] = (arr, block)
    return arr


_pinned_owner = {}


class Problem(object):
    """Owner of one ``chi_b200_problem`` handle."""

    def __init__(self, model_id, device=None):
        require_gpu()
        if device is None:
            device = default_device()
        self._h = c_vp()
        check(lib().chi_b200_problem_create(
            int(model_id), int(device), ctypes.byref(self._h)))
        self.device = device

    @property
    def handle(self):
        return self._h

    def close(self):
        if self._h is not None and self._h.value:
            lib().chi_b200_problem_destroy(self._h)
            self._h = c_vp()

    def __del__(self):
        try:
            self.close()
        except Exception:  # pragma: no cover
            pass
