"""Host-side mirror of the depth engine's C ABI (include/ngscov.h).

``CoverageEngine`` is what the reference's ``DataBuf`` + two hash tables + ``Bins`` array were
(reference bam2depth.c:15-19,326-336): create it for one BAM header, push the inflated record stream,
finish, read the four text outputs.  All work happens in libngscov_cuda on one GPU.
"""
import ctypes as C
import numpy as np
from . import _lib as L

MODE_DEPTH, MODE_WIG = 0, 1
OUT_BEDGRAPH, OUT_DEPTH, OUT_WIG, OUT_CHROMSIZE = 0, 1, 2, 3
PUSH_LAST, PUSH_TRUNCATED_OK = 1, 2


class DeviceArray:
    """a device pointer the library owns, wrapped for torch.as_tensor / cupy (zero copy, __cuda_array_interface__)"""

    def __init__(self, ptr, n, typestr="<i8"):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": typestr, "data": (int(ptr), False), "version": 2}


class CoverageEngine:
    def __init__(self, ref_names, ref_lens, mode=MODE_DEPTH, window=20000, emit_wig=False, device=0,
                 slice_beg=0, slice_end=0, staging_bytes=0):
        self.lib = L.load_library()
        self.ref_names = [n if isinstance(n, bytes) else n.encode() for n in ref_names]
        self.ref_lens = [int(x) for x in ref_lens]
        n = len(self.ref_names)
        self._names = (C.c_char_p * max(n, 1))(*self.ref_names)
        self._lens = (C.c_uint32 * max(n, 1))(*self.ref_lens)
        cfg = L.Cfg(mode=mode, window=window, emit_wig=1 if emit_wig else 0, n_ref=n,
                    ref_len=C.cast(self._lens, C.POINTER(C.c_uint32)), ref_name=C.cast(self._names, C.POINTER(C.c_char_p)),
                    device=device, slice_beg=slice_beg, slice_end=slice_end, staging_bytes=staging_bytes)
        self.mode, self.window, self.emit_wig = mode, window, emit_wig
        self.h = C.c_void_p()
        self._check(self.lib.ngscov_create(C.byref(self.h), C.byref(cfg)), "ngscov_create")
        self._staging = {}

    # -- helpers
    def _check(self, rc, what):
        if rc != 0:
            detail = self.lib.ngscov_strerror(rc).decode()
            if self.h:
                detail += "; " + self.lib.ngscov_last_cuda_error(self.h).decode()
            raise L.NgscovError(rc, what, detail)

    def close(self):
        if getattr(self, "h", None):
            self.lib.ngscov_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- input
    def staging(self, idx):
        """numpy uint8 view of pinned staging buffer idx (0/1)."""
        if idx not in self._staging:
            p, cap = C.c_void_p(), C.c_size_t()
            self._check(self.lib.ngscov_staging(self.h, idx, C.byref(p), C.byref(cap)), "ngscov_staging")
            self._staging[idx] = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(cap.value,))
        return self._staging[idx]

    def push(self, idx, nbytes, last=False, truncated_ok=False):
        flags = (PUSH_LAST if last else 0) | (PUSH_TRUNCATED_OK if truncated_ok else 0)
        self._check(self.lib.ngscov_push(self.h, idx, nbytes, flags), "ngscov_push")

    def staging_wait(self, idx):
        self._check(self.lib.ngscov_staging_wait(self.h, idx), "ngscov_staging_wait")

    def push_host(self, records, last=True, truncated_ok=False):
        """Stream a host byte buffer (bytes / numpy uint8) through the two staging buffers."""
        buf = np.frombuffer(records, dtype=np.uint8) if not isinstance(records, np.ndarray) else records
        n, off, idx = buf.size, 0, 0
        cap = self.staging(0).size
        self.staging(1)
        if n == 0:
            self.push(0, 0, last=last, truncated_ok=truncated_ok)
            return
        while off < n:
            m = min(cap, n - off)
            self.staging_wait(idx)
            self.staging(idx)[:m] = buf[off:off + m]
            off += m
            self.push(idx, m, last=last and off >= n, truncated_ok=truncated_ok)
            idx ^= 1

    def push_host_ptr(self, host_ptr, nbytes, last=True, truncated_ok=False):
        """Stream nbytes at host address host_ptr (pinned memory preferred) -- ngscov_push_host."""
        flags = (PUSH_LAST if last else 0) | (PUSH_TRUNCATED_OK if truncated_ok else 0)
        self._check(self.lib.ngscov_push_host(self.h, C.c_void_p(host_ptr), nbytes, flags), "ngscov_push_host")

    def push_bgzf(self, blocks, skip=0, last=True, truncated_ok=False):
        """BGZF blocks as they lie in the file (bytes / uint8 array): inflated on the device -- ngscov_push_bgzf."""
        a = np.ascontiguousarray(np.frombuffer(blocks, dtype=np.uint8) if isinstance(blocks, (bytes, bytearray, memoryview)) else blocks)
        flags = (PUSH_LAST if last else 0) | (PUSH_TRUNCATED_OK if truncated_ok else 0)
        self._check(self.lib.ngscov_push_bgzf(self.h, C.c_void_p(a.ctypes.data), a.size, skip, flags), "ngscov_push_bgzf")

    def push_bgzf_ptr(self, host_ptr, nbytes, skip=0, last=True, truncated_ok=False):
        flags = (PUSH_LAST if last else 0) | (PUSH_TRUNCATED_OK if truncated_ok else 0)
        self._check(self.lib.ngscov_push_bgzf(self.h, C.c_void_p(host_ptr), nbytes, skip, flags), "ngscov_push_bgzf")

    def push_device(self, dev_ptr, nbytes, last=True):
        self._check(self.lib.ngscov_push_device(self.h, C.c_void_p(dev_ptr), nbytes, PUSH_LAST if last else 0), "ngscov_push_device")

    # -- compute
    def finish(self):
        self._check(self.lib.ngscov_finish(self.h), "ngscov_finish")

    def finish_runs(self):
        self._check(self.lib.ngscov_finish_runs(self.h), "ngscov_finish_runs")

    def format(self, mask):
        self._check(self.lib.ngscov_format(self.h, mask), "ngscov_format")

    def default_mask(self):
        if self.mode == MODE_DEPTH:
            m = (1 << OUT_BEDGRAPH) | (1 << OUT_DEPTH)
            if self.emit_wig:
                m |= (1 << OUT_WIG) | (1 << OUT_CHROMSIZE)
            return m
        return (1 << OUT_WIG) | (1 << OUT_CHROMSIZE)

    def set_region(self, ref, beg, end):
        """slide mode: count only what bam_fetch(ref, beg, end) hands out (`-r` of bam_sliding_count); ref < 0 clears"""
        self._check(self.lib.ngscov_set_region(self.h, int(ref), int(beg), int(end)), "ngscov_set_region")

    def reset(self):
        self._check(self.lib.ngscov_reset(self.h), "ngscov_reset")

    def sync(self):
        self._check(self.lib.ngscov_sync(self.h), "ngscov_sync")

    @property
    def stream(self):
        return self.lib.ngscov_stream(self.h)

    # -- output
    def output(self, which):
        p, n = C.c_void_p(), C.c_size_t()
        self._check(self.lib.ngscov_output(self.h, which, C.byref(p), C.byref(n)), "ngscov_output")
        if n.value >= 1 << 31:                 # ctypes.string_at takes a C int
            return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(n.value,)).tobytes()
        return C.string_at(p, n.value) if n.value else b""

    def output_view(self, which):
        """zero-copy numpy view of the pinned host text (valid until reset/close)."""
        p, n = C.c_void_p(), C.c_size_t()
        self._check(self.lib.ngscov_output(self.h, which, C.byref(p), C.byref(n)), "ngscov_output")
        if not n.value:
            return np.zeros(0, dtype=np.uint8)
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(n.value,))

    def output_device(self, which):
        p, n = C.c_void_p(), C.c_size_t()
        self._check(self.lib.ngscov_output_device(self.h, which, C.byref(p), C.byref(n)), "ngscov_output_device")
        return p.value, n.value

    def runs(self):
        n = C.c_uint64()
        self._check(self.lib.ngscov_runs_get(self.h, None, 0, C.byref(n)), "ngscov_runs_get")
        arr = (L.Run * max(1, n.value))()
        self._check(self.lib.ngscov_runs_get(self.h, arr, n.value, C.byref(n)), "ngscov_runs_get")
        return [(r.contig, r.start, r.end, r.depth) for r in arr[:n.value]]

    RUN_DTYPE = np.dtype([("contig", "<i4"), ("start", "<u4"), ("end", "<u4"), ("depth", "<i4")])

    def runs_array(self):
        """the binary run records as a numpy structured array (contig, start, end, depth) -- ngscov_runs_get without
        building Python objects (tens of millions of runs)"""
        n = C.c_uint64()
        self._check(self.lib.ngscov_runs_get(self.h, None, 0, C.byref(n)), "ngscov_runs_get")
        out = np.zeros(max(1, n.value), dtype=self.RUN_DTYPE)
        self._check(self.lib.ngscov_runs_get(self.h, C.cast(out.ctypes.data, C.POINTER(L.Run)), n.value, C.byref(n)), "ngscov_runs_get")
        return out[:n.value]

    def slice_done(self):
        """region shard: True once a record at or beyond the end of the slice has been decoded (synchronises)"""
        d = C.c_int(0)
        self._check(self.lib.ngscov_slice_done(self.h, C.byref(d)), "ngscov_slice_done")
        return bool(d.value)

    def halo(self):
        h = L.Halo()
        self._check(self.lib.ngscov_halo_get(self.h, C.byref(h)), "ngscov_halo_get")
        f, l = h.first, h.last
        return {"n_runs": h.n_runs, "first": (f.contig, f.start, f.end, f.depth), "last": (l.contig, l.start, l.end, l.depth)}

    def halo_put(self, drop_first, new_last_end):
        self._check(self.lib.ngscov_halo_put(self.h, 1 if drop_first else 0, int(new_last_end)), "ngscov_halo_put")

    # -- device-side stitch of region shards (no host round trip)
    def halo_device(self):
        """device address of the shard's halo record (9 x int64), valid after finish_runs"""
        p = C.c_void_p()
        self._check(self.lib.ngscov_halo_device(self.h, C.byref(p)), "ngscov_halo_device")
        return p.value

    def stitch_device(self, dev_all_halos, world, rank):
        self._check(self.lib.ngscov_stitch_device(self.h, C.c_void_p(dev_all_halos), world, rank), "ngscov_stitch_device")

    def bins_device(self):
        p, n = C.c_void_p(), C.c_uint64()
        self._check(self.lib.ngscov_bins_device(self.h, C.byref(p), C.byref(n)), "ngscov_bins_device")
        return p.value, n.value

    def bins(self):
        n = C.c_uint64()
        self._check(self.lib.ngscov_bins_count(self.h, C.byref(n)), "ngscov_bins_count")
        out = np.zeros(n.value, dtype=np.uint64)
        self._check(self.lib.ngscov_bins_get(self.h, out.ctypes.data_as(C.POINTER(C.c_uint64)), n.value), "ngscov_bins_get")
        return out

    def bins_put(self, bins):
        b = np.ascontiguousarray(bins, dtype=np.uint64)
        self._check(self.lib.ngscov_bins_put(self.h, b.ctypes.data_as(C.POINTER(C.c_uint64)), b.size), "ngscov_bins_put")

    def bins_apply_run(self, run, sign, bins):
        r = L.Run(contig=run[0], start=run[1], end=run[2], depth=run[3])
        self._check(self.lib.ngscov_bins_apply_run(self.h, C.byref(r), sign, bins.ctypes.data_as(C.POINTER(C.c_uint64)), bins.size), "ngscov_bins_apply_run")

    def stats(self):
        s = L.Stats()
        self._check(self.lib.ngscov_stats(self.h, C.byref(s)), "ngscov_stats")
        d = {k: getattr(s, k) for k, _ in L.Stats._fields_ if k not in ("out_bytes", "reserved", "reserved2")}
        d["out_bytes"] = list(s.out_bytes)
        return d
