"""Host-side BAM container helpers for the Python entry points, tests and the benchmark: BGZF inflate
(zlib), BAM header parse (samtools-0.1.19 bam.c:81-121 layout) and a BAI reader (bam_index.c:326-387).
The production host driver is the C program in host/; this mirrors it for Python callers."""
import os
import struct
import zlib
import numpy as np


def bgzf_blocks(data):
    """Yield (file_offset, block_len) of every BGZF block in `data` (bytes-like)."""
    off, n = 0, len(data)
    while off + 18 <= n:
        if data[off:off + 4] != b"\x1f\x8b\x08\x04" or data[off + 12:off + 14] != b"BC":
            raise ValueError(f"not a BGZF block at offset {off}")
        bsize = struct.unpack_from("<H", data, off + 16)[0] + 1
        yield off, bsize
        off += bsize


def bgzf_inflate(data):
    """Inflate a whole BGZF file held in memory -> bytes."""
    out = []
    for off, bsize in bgzf_blocks(data):
        out.append(zlib.decompress(data[off + 18:off + bsize - 8], -15))
    return b"".join(out)


def parse_header(inflated):
    """-> (names, lens, header_bytes) from the inflated stream."""
    if inflated[:4] != b"BAM\x01":
        raise ValueError("not a BAM stream")
    l_text = struct.unpack_from("<i", inflated, 4)[0]
    p = 8 + l_text
    n_ref = struct.unpack_from("<i", inflated, p)[0]
    p += 4
    names, lens = [], []
    for _ in range(n_ref):
        l_name = struct.unpack_from("<i", inflated, p)[0]
        p += 4
        names.append(bytes(inflated[p:p + l_name - 1]))
        p += l_name
        lens.append(struct.unpack_from("<I", inflated, p)[0])
        p += 4
    return names, lens, p


def load_bam(path):
    """-> (names, lens, records uint8 array) with the whole file inflated in memory."""
    with open(path, "rb") as f:
        raw = f.read()
    inf = bgzf_inflate(raw)
    names, lens, p = parse_header(inf)
    return names, lens, np.frombuffer(inf, dtype=np.uint8)[p:]


def find_bai(bam_path):
    """The reference requires an index next to the BAM (bam2depth.c:112-119; samtools looks for
    a.bam.bai then a.bai, bam_index.c:389-419)."""
    for cand in (bam_path + ".bai", (bam_path[:-4] if bam_path.endswith(".bam") else bam_path) + ".bai"):
        if os.path.exists(cand):
            return cand
    return None


def read_bai(path):
    """-> list per contig of dict(bins={bin: [(beg,end),..]}, linear=[voffset,..]), n_no_coor."""
    with open(path, "rb") as f:
        d = f.read()
    if d[:4] != b"BAI\x01":
        raise ValueError("not a BAI file")
    p = 4
    n_ref = struct.unpack_from("<i", d, p)[0]
    p += 4
    refs = []
    for _ in range(n_ref):
        n_bin = struct.unpack_from("<i", d, p)[0]
        p += 4
        bins = {}
        for _ in range(n_bin):
            b, n_chunk = struct.unpack_from("<Ii", d, p)
            p += 8
            ch = [struct.unpack_from("<QQ", d, p + 16 * k) for k in range(n_chunk)]
            p += 16 * n_chunk
            bins[b] = ch
        n_intv = struct.unpack_from("<i", d, p)[0]
        p += 4
        lin = list(struct.unpack_from(f"<{n_intv}Q", d, p)) if n_intv else []
        p += 8 * n_intv
        refs.append({"bins": bins, "linear": lin})
    n_no_coor = struct.unpack_from("<Q", d, p)[0] if p + 8 <= len(d) else 0
    return refs, n_no_coor
