"""ctypes wrapper of the CPU oracle (oracle/ngscov_oracle.c).  TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SO = os.path.join(ROOT, "oracle", "_build", "libngscov_oracle.so")


class Result(C.Structure):
    _fields_ = [("bedgraph", C.c_void_p), ("bedgraph_len", C.c_size_t), ("depth", C.c_void_p), ("depth_len", C.c_size_t),
                ("wig", C.c_void_p), ("wig_len", C.c_size_t), ("chromsize", C.c_void_p), ("chromsize_len", C.c_size_t),
                ("n_records", C.c_uint64), ("n_kept", C.c_uint64), ("n_mops", C.c_uint64), ("cigar_words", C.c_uint64)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        l = C.CDLL(_SO)
        l.ngscov_oracle_run.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.POINTER(C.c_uint32), C.POINTER(C.c_char_p),
                                        C.c_int, C.c_uint32, C.c_int, C.POINTER(Result)]
        l.ngscov_oracle_free.argtypes = [C.POINTER(Result)]
        l.ngscov_oracle_slide.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.POINTER(C.c_uint32), C.POINTER(C.c_char_p), C.c_int,
                                          C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.POINTER(C.c_uint64)]
        l.ngscov_oracle_slide_region.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.POINTER(C.c_uint32), C.POINTER(C.c_char_p), C.c_int,
                                                 C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.POINTER(C.c_uint64)]
        l.ngscov_oracle_slide_free.argtypes = [C.c_void_p]
        _lib = l
    return _lib


def run(records, names, lens, mode, window, emit_wig=True):
    """-> dict(bedgraph, depth, wig, chromsize: bytes; n_records, n_kept, n_mops, cigar_words)."""
    rec = np.ascontiguousarray(records, dtype=np.uint8)
    n = len(names)
    nm = (C.c_char_p * max(n, 1))(*[x if isinstance(x, bytes) else x.encode() for x in names])
    ln = (C.c_uint32 * max(n, 1))(*[int(x) for x in lens])
    r = Result()
    rc = lib().ngscov_oracle_run(rec.ctypes.data, rec.size, n, ln, nm, mode, window, 1 if emit_wig else 0, C.byref(r))
    if rc:
        lib().ngscov_oracle_free(C.byref(r))
        raise RuntimeError(f"oracle failed: {rc}")
    out = {k: C.string_at(getattr(r, k), getattr(r, k + "_len")) if getattr(r, k + "_len") else b""
           for k in ("bedgraph", "depth", "wig", "chromsize")}
    out.update(n_records=r.n_records, n_kept=r.n_kept, n_mops=r.n_mops, cigar_words=r.cigar_words)
    lib().ngscov_oracle_free(C.byref(r))
    return out


def slide(records, names, lens, window, region=None):
    """bam_sliding_count restated (oracle/ngscov_slide_oracle.c) -> dict(txt: bytes of <outfile>.txt, n_count).
    region = (refID, beg, end) as bam_parse_region yields them: the `-r` mode."""
    rec = np.ascontiguousarray(records, dtype=np.uint8)
    n = len(names)
    nm = (C.c_char_p * max(n, 1))(*[x if isinstance(x, bytes) else x.encode() for x in names])
    ln = (C.c_uint32 * max(n, 1))(*[int(x) for x in lens])
    p, m, cnt = C.c_void_p(), C.c_size_t(), C.c_uint64()
    rr = region if region is not None else (-1, 0, 0)
    rc = lib().ngscov_oracle_slide_region(rec.ctypes.data, rec.size, n, ln, nm, int(window), int(rr[0]), int(rr[1]), int(rr[2]),
                                          C.byref(p), C.byref(m), C.byref(cnt))
    if rc:
        raise RuntimeError(f"slide oracle failed: {rc}")
    out = C.string_at(p, m.value) if m.value else b""
    lib().ngscov_oracle_slide_free(p)
    return {"txt": out, "n_count": cnt.value}
