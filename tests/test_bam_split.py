"""host/bamSplitChr against the unmodified reference tool (oracle/_ref/bamSplitChr, where
it was built): the same files, byte for byte, at every compression setting -- the BGZF blocking of samtools-0.1.19 is
restated exactly and both sides link this image's zlib.  Without the reference binary (GPU box) the inflated content of
each output is checked against the input's records instead."""
import gzip
import os
import struct
import subprocess

import numpy as np
import pytest

from tests import bamutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "highperformancengs_b200", "host", "bamSplitChr")
REF = os.path.join(ROOT, "oracle", "_ref", "bamSplitChr")


@pytest.fixture(scope="module")
def exe():
    if not os.path.exists(EXE):
        subprocess.check_call(["gcc", os.path.join(ROOT, "highperformancengs_b200", "host", "ngscov_split_cli.c"), "-o", EXE, "-O2", "-lz", "-lpthread"])
    return EXE


def split_both(exe, tmp_path, bam, extra=()):
    outs = []
    for k, tool in enumerate((exe, REF)):
        if k == 1 and not os.path.exists(REF):
            break
        d = tmp_path / ("o%d" % k)
        d.mkdir()
        p = subprocess.run([tool, "-o", str(d / "x"), *extra, bam], stderr=subprocess.PIPE)
        assert p.returncode == 0, p.stderr
        log = [l.split(b" at ")[0] for l in p.stderr.splitlines() if l.startswith(b"chr:")]
        outs.append(({f: open(str(d / f), "rb").read() for f in sorted(os.listdir(str(d)))}, log))
    return outs


def inflate(b):
    return gzip.decompress(b)            # BGZF = concatenated gzip members


def check_content(files, names, lens, reads, recs):
    """each output = header + the records of its contig, in order, and the standard empty last block"""
    hdr = bamutil.header_bytes(names, lens)
    raw = bytes(recs)
    per = {i: [] for i in range(len(names))}
    o = 0
    for r in reads:
        (bs,) = struct.unpack_from("<i", raw, o)
        if r[0] >= 0:
            per[r[0]].append(raw[o:o + 4 + bs])
        o += 4 + bs
    assert len(files) == len(names)
    for i, nm in enumerate(names):
        b = files["x_%s.bam" % nm.decode()]
        assert b.endswith(bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000"))
        assert inflate(b) == hdr + b"".join(per[i])


@pytest.mark.parametrize("extra", [(), ("-1", "x"), ("-u", "x")])
def test_split_matches_reference(exe, tmp_path, extra):
    names, lens, reads = bamutil.random_case(5, n_reads=40000, contigs=((b"c1", 300000), (b"cEmpty", 5000), (b"c3", 70000), (b"cLast", 100)))
    reads = [r for r in reads if r[0] not in (1, 3)]             # contigs without reads, one of them the last
    recs = bamutil.records(reads, seq_seed=7)
    bam = str(tmp_path / "t.bam")
    bamutil.write_bam(bam, names, lens, recs)
    outs = split_both(exe, tmp_path, bam, extra)
    check_content(outs[0][0], names, lens, reads, recs)
    if len(outs) == 2:
        assert outs[0][0].keys() == outs[1][0].keys()
        for f in outs[0][0]:
            assert outs[0][0][f] == outs[1][0][f], f
        assert outs[0][1] == outs[1][1]                          # the per-contig read counts on stderr


def test_long_records_and_a_large_header(exe, tmp_path):
    """records longer than one BGZF block (cut every 0xff00 bytes inside the record) and a header of several blocks"""
    rng = np.random.RandomState(3)
    names = [b"contig_with_a_rather_long_name_%05d" % i for i in range(3000)]
    lens = [400000] * len(names)
    reads = []
    for ci in (0, 7, 2999):
        pos = 0
        for k in range(60):
            n = int(rng.choice([100, 30000, 65000, 65244, 65245, 70000, 140000]))   # 65244: 4 + block_size == 0xff00 exactly
            reads.append((ci, pos, 0, "%dM" % n))
            pos += int(rng.randint(0, 50))
    recs = b""
    for i, r in enumerate(reads):
        n = int(r[3][:-1])
        # fixed-size name so that the sizes above land on the block boundary cases: 36 + 3 + 4 + (n+1)//2 + n
        recs += bamutil.record("rr", r[0], r[1], r[2], r[3], seq=True, rng=rng)
    recs = np.frombuffer(recs, dtype=np.uint8)
    bam = str(tmp_path / "t.bam")
    bamutil.write_bam(bam, names, lens, recs)
    outs = split_both(exe, tmp_path, bam)
    check_content(outs[0][0], names, lens, reads, recs)
    if len(outs) == 2:
        assert outs[0][0] == outs[1][0]


def test_generated_config_matches_reference(exe, tmp_path):
    if not os.path.exists(REF):
        pytest.skip("oracle/_ref/bamSplitChr not built (needs /root/reference)")
    from highperformancengs_b200.gen import Generator
    bam = str(tmp_path / "g.bam")
    Generator(2, scale=0.0005).write_bam(bam)
    outs = split_both(exe, tmp_path, bam)
    assert len(outs[0][0]) == 24 and outs[0][0] == outs[1][0] and outs[0][1] == outs[1][1]


def test_default_prefix_and_missing_index(exe, tmp_path):
    names, lens, reads = bamutil.random_case(6, n_reads=500)
    recs = bamutil.records(reads)
    bam = str(tmp_path / "t.bam")
    bamutil.write_bam(bam, names, lens, recs)
    subprocess.check_call([exe, bam], stderr=subprocess.DEVNULL)
    for nm in names:
        assert os.path.exists("%s_%s.bam" % (bam, nm.decode()))   # bamSplitChr.c:120: the input name is the prefix
    os.remove(bam + ".bai")
    p = subprocess.run([exe, bam], stderr=subprocess.PIPE)
    assert p.returncode == 1 and b"BAM indexing file is not available" in p.stderr   # bamSplitChr.c:63-70


def test_records_with_a_contig_but_no_position_are_left_out(exe, tmp_path):
    """refID >= 0 with POS -1 (placed-unmapped mates written that way by some aligners): bam_fetch never hands them out
    (is_overlap compares the position as uint32), so they are in none of the per-contig files"""
    sam2bam = os.path.join(ROOT, "oracle", "_ref", "sam2bam")
    names, lens, reads = bamutil.random_case(9, n_reads=2000)
    extra = [(0, -1, 4, "*"), (0, -1, 4, "*"), (2, -1, 69, "*")]
    reads = sorted([r for r in reads if r[0] >= 0] + extra, key=lambda r: (r[0], r[1])) + [r for r in reads if r[0] < 0]
    recs = bamutil.records(reads, seq_seed=1)
    bam = str(tmp_path / "t.bam")
    bamutil.write_bam(bam, names, lens, recs, index=False)
    if os.path.exists(sam2bam):
        subprocess.check_call([sam2bam, "--index-only", bam])
    else:
        open(bam + ".bai", "wb").write(b"BAI\1")          # the splitter only insists that the file exists
    outs = split_both(exe, tmp_path, bam)
    hdr = bamutil.header_bytes(names, lens)
    raw, o, per = bytes(recs), 0, {i: b"" for i in range(len(names))}
    for r in reads:
        (bs,) = struct.unpack_from("<i", raw, o)
        if r[0] >= 0 and r[1] >= 0:
            per[r[0]] += raw[o:o + 4 + bs]
        o += 4 + bs
    for i, nm in enumerate(names):
        assert inflate(outs[0][0]["x_%s.bam" % nm.decode()]) == hdr + per[i]
    if len(outs) == 2:
        assert outs[0][0] == outs[1][0] and outs[0][1] == outs[1][1]
