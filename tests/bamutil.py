"""Test helpers: build BAM records / files from small read specs (pure Python).  TEST INFRASTRUCTURE."""
import os
import re
import struct
import subprocess
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CIGAR_OPS = "MIDNSHP=XB"
_REF_CONSUME = set("MDN=X")
_QRY_CONSUME = set("MIS=X")


def reg2bin(beg, end):
    end -= 1
    if beg >> 14 == end >> 14:
        return 4681 + (beg >> 14)
    if beg >> 17 == end >> 17:
        return 585 + (beg >> 17)
    if beg >> 20 == end >> 20:
        return 73 + (beg >> 20)
    if beg >> 23 == end >> 23:
        return 9 + (beg >> 23)
    if beg >> 26 == end >> 26:
        return 1 + (beg >> 26)
    return 0


def parse_cigar(s):
    if s in ("*", ""):
        return []
    return [(int(n), op) for n, op in re.findall(r"(\d+)([MIDNSHP=XB])", s)]


def record(name, refid, pos, flag, cigar, seq=True, mapq=60, aux=b"", rng=None):
    """One BAM record (with the 4-byte block_size prefix).  pos is 0-based.  rng: numpy RandomState for random SEQ codes
    (default: the constant 0x12 pattern, enough for the coverage tools, which never read SEQ)."""
    ops = parse_cigar(cigar)
    qlen = sum(n for n, op in ops if op in _QRY_CONSUME)
    rlen = sum(n for n, op in ops if op in _REF_CONSUME)
    l_seq = qlen if seq else 0
    qn = name.encode() + b"\0"
    if refid < 0:
        b = 4680
    else:
        b = reg2bin(pos, pos + (rlen if rlen else 1))
    core = struct.pack("<iiBBHHHiiii", refid, pos, len(qn), mapq if refid >= 0 else 0, b, len(ops), flag, l_seq, -1, -1, 0)
    cig = b"".join(struct.pack("<I", n << 4 | CIGAR_OPS.index(op)) for n, op in ops)
    if rng is None:
        sq = bytes([0x12] * ((l_seq + 1) // 2)) + bytes([30] * l_seq)
    else:
        codes = rng.choice(np.array([1, 2, 4, 8, 15, 2, 4], dtype=np.uint8), size=l_seq + (l_seq & 1))      # A C G T N, GC-rich
        if l_seq & 1:
            codes[-1] = 0
        sq = bytes((codes[0::2] << 4 | codes[1::2]).astype(np.uint8).tolist()) + bytes([30] * l_seq)
    body = core + qn + cig + sq + aux
    return struct.pack("<i", len(body)) + body


def records(reads, seq=True, seq_seed=None):
    """reads: iterable of (refid, pos, flag, cigar) -> uint8 array of the record stream."""
    rng = np.random.RandomState(seq_seed) if seq_seed is not None else None
    out = [record("r%d" % i, r[0], r[1], r[2], r[3], seq=seq, rng=rng) for i, r in enumerate(reads)]
    return np.frombuffer(b"".join(out), dtype=np.uint8)


def header_bytes(names, lens):
    text = "@HD\tVN:1.0\tSO:coordinate\n" + "".join("@SQ\tSN:%s\tLN:%d\n" % (n.decode() if isinstance(n, bytes) else n, l) for n, l in zip(names, lens))
    out = b"BAM\1" + struct.pack("<i", len(text)) + text.encode() + struct.pack("<i", len(names))
    for n, l in zip(names, lens):
        nb = (n if isinstance(n, bytes) else n.encode()) + b"\0"
        out += struct.pack("<i", len(nb)) + nb + struct.pack("<I", l)
    return out


def bgzf_compress(data, level=6, block=0xff00):
    out = []
    for i in range(0, len(data), block):
        chunk = bytes(data[i:i + block])
        co = zlib.compressobj(level, zlib.DEFLATED, -15)
        c = co.compress(chunk) + co.flush()
        out.append(b"\x1f\x8b\x08\x04\0\0\0\0\0\xff\x06\0BC\x02\0" + struct.pack("<H", len(c) + 25) + c +
                   struct.pack("<II", zlib.crc32(chunk) & 0xffffffff, len(chunk)))
    out.append(bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000"))
    return b"".join(out)


def write_bam(path, names, lens, recs, index=True):
    """Write a BGZF BAM; index it with the reference's own samtools indexer when oracle/_ref exists."""
    with open(path, "wb") as f:
        f.write(bgzf_compress(header_bytes(names, lens) + bytes(recs)))
    if index:
        tool = os.path.join(ROOT, "oracle", "_ref", "sam2bam")
        if os.path.exists(tool):
            subprocess.check_call([tool, "--index-only", path], stderr=subprocess.DEVNULL)
        else:
            open(path + ".bai", "wb").write(b"BAI\1" + struct.pack("<i", len(names)) + b"".join(struct.pack("<ii", 0, 0) for _ in names) + struct.pack("<Q", 0))


# ---- small hand-written cases --------------------------------------------------------------------
def toy_case():
    """samtools examples/toy.sam (12 reads, M/I/D/N/S/H/P ops), re-typed from the SAM spec example."""
    names, lens = [b"ref", b"ref2"], [45, 40]
    reads = [(0, 6, 163, "8M4I4M1D3M"), (0, 8, 0, "1S2I6M1P1I1P1I4M2I"), (0, 8, 0, "5H6M"), (0, 12, 0, "6M14N1I5M"),
             (0, 28, 16, "6H5M"), (0, 36, 83, "9M"), (1, 0, 0, "20M"), (1, 1, 0, "21M"), (1, 5, 0, "9M4I13M"),
             (1, 9, 0, "25M"), (1, 11, 0, "24M"), (1, 13, 0, "23M")]
    return names, lens, reads


def quirk_case():
    """SURVEY.md 8(c) quirk SAM: flag filters, =/X no-advance, clips, empty contig, read at contig end."""
    names, lens = [b"cA", b"cEmpty", b"cB"], [100, 60, 50]
    reads = [(0, 0, 0, "10M"), (0, 10, 0, "10M"), (0, 10, 0, "5=5X"), (0, 30, 256, "10M"), (0, 30, 1024, "10M"),
             (0, 30, 512, "10M"), (0, 30, 4, "10M"), (0, 40, 16, "5S5M5=5M5N5M2D3M4H"), (0, 90, 2048, "10M"),
             (2, 44, 0, "6M")]
    return names, lens, reads


def random_case(seed, n_reads=3000, contigs=((b"c1", 30000), (b"c2", 1000), (b"c3", 70000)), unmapped_tail=20):
    """Random flags/CIGARs incl. reads at position 0 and ending exactly at the contig end."""
    rng = np.random.RandomState(seed)
    cigs = ["150M", "70M2D80M", "5S145M", "60M3I87M", "50M1000N100M", "20M5=5X20M", "10H140M", "1M", "30M2N30M2N30M",
            "3S20M1I20M1D20M3S", "100M", "75M75S", "2M2D2M2D2M"]
    flags = [0, 16, 99, 147, 4, 256, 512, 1024, 2048, 0, 0, 16]
    names = [c[0] for c in contigs]
    lens = [c[1] for c in contigs]
    reads = []
    for ci, (nm, L) in enumerate(contigs):
        k = max(1, int(n_reads * L / sum(lens)))
        rs = []
        for _ in range(k):
            cg = cigs[rng.randint(len(cigs))]
            rl = sum(n for n, op in parse_cigar(cg) if op in _REF_CONSUME)
            if rl > L:
                cg, rl = "1M", 1
            u = rng.rand()
            pos = 0 if u < 0.02 else (L - rl if u < 0.04 else rng.randint(0, L - rl + 1))
            rs.append((ci, int(pos), flags[rng.randint(len(flags))], cg))
        rs.sort(key=lambda r: r[1])
        reads += rs
    reads += [(-1, -1, 4, "*")] * unmapped_tail
    return names, lens, reads
