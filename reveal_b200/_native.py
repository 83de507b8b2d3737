"""ctypes binding of libreveal_b200.so (C-ABI declared in include/reveal_b200.h).

There is NO CPU fallback: if the CUDA library has not been built, or no CUDA
device is present, every operation raises.  Build with
``python -c "import __graft_entry__ as g; g.build()"`` (or ``reveal_b200.build.build()``).
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libreveal_b200.so")

c_i64p = ctypes.POINTER(ctypes.c_int64)
c_vp = ctypes.c_void_p


class Times(ctypes.Structure):
    _fields_ = [("h2d_ms", ctypes.c_float), ("pack_ms", ctypes.c_float), ("sa_ms", ctypes.c_float),
                ("lcp_ms", ctypes.c_float), ("so_ms", ctypes.c_float), ("total_ms", ctypes.c_float),
                ("sa_rounds", ctypes.c_int32), ("launches", ctypes.c_int32), ("sa_sorted_items", ctypes.c_int64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class KernelProfile(ctypes.Structure):
    SLOTS = ("rs_pass_kernel", "sa_place_kernel", "sweep kernels", "lcp_sparse_kernel", "sa_lead_kernel", "text passes", "stage-4 round kernels", "")
    _fields_ = [("ms", ctypes.c_double * 8), ("launches", ctypes.c_int64 * 8), ("bytes", ctypes.c_int64 * 8),
                ("launches_total", ctypes.c_int64)]

    def as_dict(self):
        return {name: {"ms": self.ms[k], "launches": self.launches[k], "bytes": self.bytes[k]} for k, name in enumerate(self.SLOTS)}


# every symbol include/reveal_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "rv_last_error": (ctypes.c_char_p, []),
    "rv_version": (ctypes.c_char_p, []),
    "rv_device_count": (ctypes.c_int, [ctypes.POINTER(ctypes.c_int)]),
    "rv_set_device": (ctypes.c_int, [ctypes.c_int]),
    "rv_host_alloc": (ctypes.c_int, [ctypes.c_int64, ctypes.POINTER(c_vp)]),
    "rv_host_free": (None, [c_vp]),
    "rv_trim": (ctypes.c_int, []),
    "rv_index_create": (ctypes.c_int, [ctypes.POINTER(c_vp), c_vp]),
    "rv_index_free": (None, [c_vp]),
    "rv_build": (ctypes.c_int, [c_vp, c_vp, ctypes.c_int64, c_vp, ctypes.c_int32, ctypes.c_int32]),
    "rv_build_device": (ctypes.c_int, [c_vp, c_vp, ctypes.c_int64, c_vp, ctypes.c_int32, ctypes.c_int32]),
    "rv_build_cached": (ctypes.c_int, [c_vp, c_vp, ctypes.c_int64, c_vp, ctypes.c_int32, ctypes.c_int32, c_vp, c_vp]),
    "rv_get_times": (ctypes.c_int, [c_vp, ctypes.POINTER(Times)]),
    "rv_index_n": (ctypes.c_int64, [c_vp]),
    "rv_profile": (ctypes.c_int, [c_vp, ctypes.c_int32]),
    "rv_get_profile": (ctypes.c_int, [c_vp, ctypes.POINTER(KernelProfile)]),
    "rv_get_sa": (ctypes.c_int, [c_vp, c_vp, ctypes.c_int32]),
    "rv_get_sai": (ctypes.c_int, [c_vp, c_vp, ctypes.c_int32]),
    "rv_get_lcp": (ctypes.c_int, [c_vp, c_vp, ctypes.c_int32]),
    "rv_get_so": (ctypes.c_int, [c_vp, c_vp]),
    "rv_get_text": (ctypes.c_int, [c_vp, c_vp]),
    "rv_put_text": (ctypes.c_int, [c_vp, ctypes.c_int64, c_vp, ctypes.c_int64]),
    "rv_device_arrays": (ctypes.c_int, [c_vp] + [ctypes.POINTER(c_vp)] * 5),
    "rv_mums_pair_count": (ctypes.c_int, [c_vp, ctypes.c_int32, ctypes.c_int32, c_i64p]),
    "rv_mums_pair_fetch": (ctypes.c_int, [c_vp, c_vp, ctypes.c_int64]),
    "rv_mums_multi_count": (ctypes.c_int, [c_vp, ctypes.c_int32, ctypes.c_int32, c_i64p, c_i64p]),
    "rv_mems_multi_count": (ctypes.c_int, [c_vp, ctypes.c_int32, ctypes.c_int32, c_i64p, c_i64p]),
    "rv_mums_multi_fetch": (ctypes.c_int, [c_vp, c_vp, ctypes.c_int64, c_vp, ctypes.c_int64]),
    "rv_result_device": (ctypes.c_int, [c_vp, ctypes.POINTER(c_vp), c_i64p, ctypes.POINTER(c_vp), c_i64p]),
    "rv_sub_root": (ctypes.c_int, [c_vp, ctypes.POINTER(c_vp)]),
    "rv_sub_n": (ctypes.c_int64, [c_vp]),
    "rv_sub_free": (None, [c_vp]),
    "rv_sub_get": (ctypes.c_int, [c_vp, ctypes.c_int32, c_vp]),
    "rv_sub_mums_pair": (ctypes.c_int, [c_vp, ctypes.c_int32, c_i64p]),
    "rv_sub_mums_multi": (ctypes.c_int, [c_vp, ctypes.c_int32, ctypes.c_int32, c_i64p, c_i64p]),
    "rv_sub_split": (ctypes.c_int, [c_vp, c_vp, ctypes.c_int32, c_vp, ctypes.c_int32, c_vp, ctypes.c_int32, c_vp, ctypes.c_int32,
                                    ctypes.c_int64, c_vp, ctypes.c_int32, ctypes.POINTER(c_vp)]),
    "rv_sub_step": (ctypes.c_int, [c_vp, c_vp, ctypes.c_int32, c_vp, ctypes.c_int32, c_vp, ctypes.c_int32, c_vp, ctypes.c_int32,
                                   ctypes.c_int64, c_vp, ctypes.c_int32, c_vp, ctypes.c_int32, ctypes.c_int32, ctypes.POINTER(c_vp)]),
    "rv_rec_stats": (ctypes.c_int, [c_vp, c_i64p, ctypes.POINTER(ctypes.c_double)]),
    "rv_rec_launches": (ctypes.c_int, [c_vp, c_i64p]),
    "rv_sub_extract": (ctypes.c_int, [c_vp, c_vp, ctypes.c_int32]),
    "rv_mums_tiny_batch": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, c_vp, c_vp]),
    "rv_chain_batch": (ctypes.c_int, [c_vp, ctypes.c_int32, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, ctypes.c_int64, ctypes.c_int32, c_vp, c_vp]),
    "rv_sub_step_batch": (ctypes.c_int, [c_vp, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32]),
    "rv_sub_step_batch_begin": (ctypes.c_int, [c_vp, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.POINTER(c_vp)]),
    "rv_sub_step_batch_end": (ctypes.c_int, [c_vp]),
    "rv_sub_fetch": (ctypes.c_int, [c_vp, c_vp, ctypes.c_int64, c_vp, ctypes.c_int64]),
    "rv_result_pack_device": (ctypes.c_int, [c_vp, c_vp, ctypes.c_int64]),
    "rv_peer_alloc": (ctypes.c_int, [ctypes.c_int64, ctypes.POINTER(c_vp), c_vp]),
    "rv_peer_open": (ctypes.c_int, [c_vp, ctypes.POINTER(c_vp)]),
    "rv_peer_close": (ctypes.c_int, [c_vp]),
    "rv_peer_free": (ctypes.c_int, [c_vp]),
    "rv_peer_read": (ctypes.c_int, [c_vp, c_vp, ctypes.c_int64]),
    "rv_sweep_pair_device": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                                            ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, c_i64p]),
    "rv_sweep_multi_device": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, ctypes.c_int64, ctypes.c_int64, ctypes.c_int32,
                                             ctypes.c_int32, ctypes.c_int32, c_i64p, c_i64p]),
}


class NativeError(RuntimeError):
    """A C-ABI call returned a negative status (message from rv_last_error)."""

    def __init__(self, code, msg):
        RuntimeError.__init__(self, "%s (rv_status %d)" % (msg, code))
        self.code = code
        self.msg = msg


def bind(path):
    """Load a library exporting the C-ABI and set every prototype."""
    L = ctypes.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(L, name)  # AttributeError if the library lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    return L


_lib = None


def lib():
    """The product library. Raises if it is not built -- never falls back to a CPU path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("reveal_b200: %s is missing; build the CUDA library first "
                              "(python -c 'import __graft_entry__ as g; g.build()'). There is no CPU fallback." % LIB_PATH)
        _lib = bind(LIB_PATH)
    return _lib


def check(L, status):
    if status != 0:
        raise NativeError(status, L.rv_last_error().decode("utf-8", "replace"))
