"""ctypes binding of libngscov_cuda.so (include/ngscov.h).  Fails loudly when the library is missing:
the product path never falls back to anything else."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libngscov_cuda.so")


class NgscovError(RuntimeError):
    def __init__(self, code, what, detail=""):
        self.code = code
        super().__init__(f"{what}: ngscov error {code} ({detail})")


class Cfg(C.Structure):
    _fields_ = [("mode", C.c_int32), ("window", C.c_uint32), ("emit_wig", C.c_int32), ("n_ref", C.c_int32),
                ("ref_len", C.POINTER(C.c_uint32)), ("ref_name", C.POINTER(C.c_char_p)),
                ("device", C.c_int32), ("reserved0", C.c_int32),
                ("slice_beg", C.c_uint64), ("slice_end", C.c_uint64), ("staging_bytes", C.c_uint64),
                ("flags", C.c_uint32), ("reserved1", C.c_uint32)]


class Run(C.Structure):
    _fields_ = [("contig", C.c_int32), ("start", C.c_uint32), ("end", C.c_uint32), ("depth", C.c_int32)]


class Halo(C.Structure):
    _fields_ = [("n_runs", C.c_uint64), ("first", Run), ("last", Run)]


class Stats(C.Structure):
    _fields_ = [("records", C.c_uint64), ("kept", C.c_uint64), ("m_ops", C.c_uint64), ("cigar_words", C.c_uint64),
                ("stream_bytes", C.c_uint64), ("boundaries", C.c_uint64), ("runs", C.c_uint64), ("positions", C.c_uint64),
                ("out_bytes", C.c_uint64 * 4), ("spec_subchunks", C.c_uint64), ("spec_repaired", C.c_uint64),
                ("ms_zero", C.c_float), ("ms_ingest", C.c_float), ("ms_scan", C.c_float), ("ms_format", C.c_float),
                ("ms_h2d", C.c_float), ("ms_d2h", C.c_float), ("launches", C.c_uint32), ("ms_inflate", C.c_float),
                ("ms_decode", C.c_float), ("reserved2", C.c_float)]


# every symbol include/ngscov.h declares: (restype, argtypes)
_H = C.c_void_p
SYMBOLS = {
    "ngscov_abi_version": (C.c_int, []),
    "ngscov_create": (C.c_int, [C.POINTER(_H), C.POINTER(Cfg)]),
    "ngscov_destroy": (None, [_H]),
    "ngscov_strerror": (C.c_char_p, [C.c_int]),
    "ngscov_last_cuda_error": (C.c_char_p, [_H]),
    "ngscov_staging": (C.c_int, [_H, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]),
    "ngscov_push": (C.c_int, [_H, C.c_int, C.c_size_t, C.c_int]),
    "ngscov_staging_wait": (C.c_int, [_H, C.c_int]),
    "ngscov_push_host": (C.c_int, [_H, C.c_void_p, C.c_size_t, C.c_int]),
    "ngscov_push_device": (C.c_int, [_H, C.c_void_p, C.c_size_t, C.c_int]),
    "ngscov_finish_runs": (C.c_int, [_H]),
    "ngscov_slice_done": (C.c_int, [_H, C.POINTER(C.c_int)]),
    "ngscov_contigs_done": (C.c_int, [_H, C.POINTER(C.c_uint32)]),
    "ngscov_set_region": (C.c_int, [_H, C.c_int, C.c_int, C.c_int]),
    "ngscov_push_bgzf": (C.c_int, [_H, C.c_void_p, C.c_size_t, C.c_size_t, C.c_int]),
    "ngscov_halo_get": (C.c_int, [_H, C.POINTER(Halo)]),
    "ngscov_halo_put": (C.c_int, [_H, C.c_int, C.c_uint32]),
    "ngscov_halo_device": (C.c_int, [_H, C.POINTER(C.c_void_p)]),
    "ngscov_stitch_device": (C.c_int, [_H, C.c_void_p, C.c_int, C.c_int]),
    "ngscov_bins_device": (C.c_int, [_H, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]),
    "ngscov_bins_count": (C.c_int, [_H, C.POINTER(C.c_uint64)]),
    "ngscov_bins_get": (C.c_int, [_H, C.POINTER(C.c_uint64), C.c_uint64]),
    "ngscov_bins_put": (C.c_int, [_H, C.POINTER(C.c_uint64), C.c_uint64]),
    "ngscov_bins_apply_run": (C.c_int, [_H, C.POINTER(Run), C.c_int, C.POINTER(C.c_uint64), C.c_uint64]),
    "ngscov_format": (C.c_int, [_H, C.c_uint32]),
    "ngscov_finish": (C.c_int, [_H]),
    "ngscov_output": (C.c_int, [_H, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]),
    "ngscov_output_device": (C.c_int, [_H, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]),
    "ngscov_runs_get": (C.c_int, [_H, C.POINTER(Run), C.c_uint64, C.POINTER(C.c_uint64)]),
    "ngscov_stats": (C.c_int, [_H, C.POINTER(Stats)]),
    "ngscov_reset": (C.c_int, [_H]),
    "ngscov_sync": (C.c_int, [_H]),
    "ngscov_stream": (C.c_void_p, [_H]),
}

_lib = None


def load_library(path=None):
    """Load libngscov_cuda.so and bind every symbol of include/ngscov.h.  Raises if it is not built."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or os.environ.get("NGSCOV_LIB") or LIB_PATH
    if not os.path.exists(p):
        raise ImportError(f"{p} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "or `make -C highperformancengs_b200/csrc` (there is no CPU fallback)")
    lib = C.CDLL(p)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)          # AttributeError if the header and the library disagree
        fn.restype = res
        fn.argtypes = args
    if path is None:
        _lib = lib
    return lib
