"""ctypes binding of tools/libbamgen.so -- the synthetic BAM generator (inputs only)."""
import ctypes as C
import os
import subprocess
import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SO = os.path.join(_ROOT, "tools", "libbamgen.so")
_lib = None


def build(force=False):
    src = os.path.join(_ROOT, "tools", "bamgen.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O3", "-fopenmp", "-shared", "-fPIC", src, "-o", _SO, "-lz", "-lm"])
    exe = os.path.join(_ROOT, "tools", "bamgen")
    if force or not os.path.exists(exe) or os.path.getmtime(exe) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O3", "-fopenmp", "-DBAMGEN_MAIN", src, "-o", exe, "-lz", "-lm"])


def lib():
    global _lib
    if _lib is None:
        build()
        l = C.CDLL(_SO)
        l.bamgen_create.restype = C.c_void_p
        l.bamgen_create.argtypes = [C.c_int, C.c_uint64, C.c_double, C.c_int]
        l.bamgen_destroy.argtypes = [C.c_void_p]
        l.bamgen_n_ref.argtypes = [C.c_void_p]
        l.bamgen_ref_name.restype = C.c_char_p
        l.bamgen_ref_name.argtypes = [C.c_void_p, C.c_int]
        l.bamgen_ref_len.restype = C.c_uint32
        l.bamgen_ref_len.argtypes = [C.c_void_p, C.c_int]
        l.bamgen_n_reads.restype = C.c_uint64
        l.bamgen_n_reads.argtypes = [C.c_void_p]
        l.bamgen_contig_reads.restype = C.c_uint64
        l.bamgen_contig_reads.argtypes = [C.c_void_p, C.c_int]
        l.bamgen_records_size.restype = C.c_uint64
        l.bamgen_records_size.argtypes = [C.c_void_p, C.c_int, C.c_int]
        l.bamgen_records.restype = C.c_uint64
        l.bamgen_records.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_uint64]
        l.bamgen_tile_offsets.restype = C.c_uint64
        l.bamgen_tile_offsets.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_uint64]
        l.bamgen_write_bam.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
        _lib = l
    return _lib


class Generator:
    """One of the five synthetic configurations (1..5), optionally scaled down."""

    def __init__(self, config, scale=1.0, seed=0, seq=None):
        self.l = lib()
        if seq is None:
            seq = (config == 1)
        self.g = self.l.bamgen_create(config, seed, float(scale), 1 if seq else 0)
        if not self.g:
            raise ValueError("bad generator config")
        self.config, self.scale, self.seq = config, scale, bool(seq)
        self.n_ref = self.l.bamgen_n_ref(self.g)
        self.names = [self.l.bamgen_ref_name(self.g, i) for i in range(self.n_ref)]
        self.lens = [self.l.bamgen_ref_len(self.g, i) for i in range(self.n_ref)]
        self.n_reads = self.l.bamgen_n_reads(self.g)

    def __del__(self):
        if getattr(self, "g", None):
            self.l.bamgen_destroy(self.g)
            self.g = None

    def records_size(self, c0=0, c1=None):
        return self.l.bamgen_records_size(self.g, c0, self.n_ref + 1 if c1 is None else c1)

    def records_into(self, ptr, cap, c0=0, c1=None):
        """Generate the record stream of contigs [c0,c1) (n_ref = unmapped tail) at address ptr."""
        n = self.l.bamgen_records(self.g, c0, self.n_ref + 1 if c1 is None else c1, C.c_void_p(ptr), cap)
        if n == 0 and self.records_size(c0, c1):
            raise MemoryError("generator buffer too small")
        return n

    def records(self, c0=0, c1=None):
        n = self.records_size(c0, c1)
        buf = np.empty(n, dtype=np.uint8)
        self.records_into(buf.ctypes.data, n, c0, c1)
        return buf

    def tile_offsets(self, ci, c0=0):
        """Stream offset (within contigs [c0,..)) of the first record STARTING in each 16 kbp tile of contig ci."""
        cap = self.lens[ci] // 16384 + 2
        off = np.zeros(cap, dtype=np.uint64)
        n = self.l.bamgen_tile_offsets(self.g, c0, ci, off.ctypes.data_as(C.c_void_p), cap)
        return off[:n]

    def contig_reads(self, ci):
        return self.l.bamgen_contig_reads(self.g, ci)

    def write_bam(self, path, level=1):
        rc = self.l.bamgen_write_bam(self.g, path.encode(), level)
        if rc:
            raise IOError(f"bamgen_write_bam failed: {rc}")
