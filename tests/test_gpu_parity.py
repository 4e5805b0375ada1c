"""GPU parity tests: libngscov_cuda (through the C ABI) against the CPU oracle, bit-exact."""
import os

import numpy as np
import pytest

from tests import bamutil

pytestmark = pytest.mark.gpu

MODES = [(0, "depth"), (1, "wig")]


def run_engine(names, lens, recs, mode, window, emit_wig=True, staging_bytes=0, chunk=None, device_resident=False):
    from highperformancengs_b200 import CoverageEngine
    eng = CoverageEngine(names, lens, mode=mode, window=window, emit_wig=emit_wig, staging_bytes=staging_bytes)
    try:
        if device_resident:
            import torch
            t = torch.from_numpy(np.ascontiguousarray(recs)).cuda()
            torch.cuda.synchronize()
            eng.push_device(t.data_ptr(), t.numel(), last=True)
        else:
            eng.push_host(recs, last=True)
        eng.finish()
        out = {"bedgraph": eng.output(0) if mode == 0 else b"", "depth": eng.output(1) if mode == 0 else b"",
               "wig": eng.output(2) if (mode == 1 or emit_wig) else b"", "chromsize": eng.output(3) if (mode == 1 or emit_wig) else b""}
        out["stats"] = eng.stats()
        return out
    finally:
        eng.close()


def check(got, exp, mode, what):
    for k in (("bedgraph", "depth", "wig", "chromsize") if mode == 0 else ("wig", "chromsize")):
        assert got[k] == exp[k], f"{what}: {k} differs ({len(got[k])} vs {len(exp[k])} bytes)"
    assert got["stats"]["records"] == exp["n_records"], what
    assert got["stats"]["kept"] == exp["n_kept"], what


@pytest.mark.parametrize("mode,mname", MODES)
@pytest.mark.parametrize("case", ["toy", "quirk"])
@pytest.mark.parametrize("window", [1, 7, 10, 25, 20000])
def test_small_cases(oracle, mode, mname, case, window):
    names, lens, reads = bamutil.toy_case() if case == "toy" else bamutil.quirk_case()
    recs = bamutil.records(reads)
    exp = oracle.run(recs, names, lens, mode, window)
    got = run_engine(names, lens, recs, mode, window)
    check(got, exp, mode, f"{case} {mname} w={window}")


@pytest.mark.parametrize("mode,mname", MODES)
@pytest.mark.parametrize("seed", [1, 2, 3])
@pytest.mark.parametrize("window", [150, 999, 1000, 20000])
def test_random_cases(oracle, mode, mname, seed, window):
    names, lens, reads = bamutil.random_case(seed)
    recs = bamutil.records(reads, seq=(seed != 2))
    exp = oracle.run(recs, names, lens, mode, window)
    got = run_engine(names, lens, recs, mode, window)
    check(got, exp, mode, f"random{seed} {mname} w={window}")


@pytest.mark.parametrize("staging", [4096, 70000])
@pytest.mark.parametrize("sub", ["64", "512"])
@pytest.mark.parametrize("variant", ["fused", "split", "tiles", "tiles2"])
def test_straddling_pushes(oracle, monkeypatch, staging, sub, variant):
    """records cut by push boundaries (carry path) and tiny sub-chunks (speculation + repair path), for
    every ingest organisation: fused walk+decode, K1 frame / K2 decode split, shared-memory staged tiles"""
    monkeypatch.setenv("NGSCOV_SUBCHUNK", sub)
    if variant == "fused":
        monkeypatch.setenv("NGSCOV_FUSED", "1")
    if variant in ("tiles", "tiles2"):
        monkeypatch.setenv("NGSCOV_INGEST", variant)
    names, lens, reads = bamutil.random_case(7, n_reads=4000)
    recs = bamutil.records(reads)
    exp = oracle.run(recs, names, lens, 0, 500)
    got = run_engine(names, lens, recs, 0, 500, staging_bytes=staging)
    check(got, exp, 0, f"straddle staging={staging} sub={sub}")


def test_empty_input(oracle):
    names, lens = [b"chrA", b"chrB"], [1000, 20000]
    recs = np.zeros(0, dtype=np.uint8)
    for mode, _ in MODES:
        exp = oracle.run(recs, names, lens, mode, 300)
        got = run_engine(names, lens, recs, mode, 300)
        check(got, exp, mode, "empty")


@pytest.mark.parametrize("config,scale,window", [(1, 0.05, 20000), (2, 0.0005, 20000), (4, 0.01, 100), (4, 0.01, 20000), (5, 0.002, 20000)])
@pytest.mark.parametrize("resident", [False, True])
@pytest.mark.parametrize("variant", ["fused", "split", "tiles2"])
def test_generator_configs(oracle, monkeypatch, config, scale, window, resident, variant):
    if variant == "fused":
        monkeypatch.setenv("NGSCOV_FUSED", "1")
    if variant == "tiles2":
        monkeypatch.setenv("NGSCOV_INGEST", "tiles2")
    from highperformancengs_b200.gen import Generator
    g = Generator(config, scale=scale)
    recs = g.records()
    mode = 1 if config == 4 else 0
    exp = oracle.run(recs, g.names, g.lens, mode, window)
    got = run_engine(g.names, g.lens, recs, mode, window, device_resident=resident)
    check(got, exp, mode, f"config{config} scale={scale} w={window} resident={resident}")
    assert got["stats"]["records"] == g.n_reads


def test_truncated_stream_is_an_error(oracle):
    from highperformancengs_b200 import NgscovError
    names, lens, reads = bamutil.toy_case()
    recs = bamutil.records(reads)
    with pytest.raises(NgscovError) as e:
        run_engine(names, lens, recs[:-7], 0, 10)
    assert e.value.code == -4


def test_overhang_beyond_pad_is_range_error(oracle):
    from highperformancengs_b200 import NgscovError
    names, lens = [b"c"], [1000]
    recs = bamutil.records([(0, 900, 0, "100M200000N50M")], seq=False)
    with pytest.raises(NgscovError) as e:
        run_engine(names, lens, recs, 0, 100)
    assert e.value.code == -5


def test_small_overhang_matches_oracle(oracle):
    """a read hanging over the contig end by less than a window is defined in the reference: bedGraph
    coordinates run past the end, bins stop at the end"""
    names, lens = [b"c"], [1000]
    recs = bamutil.records([(0, 100, 0, "50M"), (0, 950, 0, "80M"), (0, 960, 16, "60M")], seq=False)
    for mode, _ in MODES:
        exp = oracle.run(recs, names, lens, mode, 300)
        got = run_engine(names, lens, recs, mode, 300)
        check(got, exp, mode, "overhang")


def test_runs_get_matches_bedgraph(oracle):
    """the binary run records (single push, no streamed finalize) are the bedGraph lines"""
    from highperformancengs_b200 import CoverageEngine
    names, lens, reads = bamutil.random_case(9, n_reads=2000)
    recs = bamutil.records(reads, seq=False)
    exp = oracle.run(recs, names, lens, 0, 500)
    with CoverageEngine(names, lens, mode=0, window=500) as eng:
        eng.push_host(recs, last=True)
        eng.finish()
        runs = eng.runs()
    text = "".join("%s\t%d\t%d\t%d\n" % (names[c].decode(), s, e, d) for c, s, e, d in runs)
    assert text.encode() == exp["bedgraph"]


@pytest.mark.parametrize("staging", [8192, 40000])
def test_streamed_per_contig_finalize(oracle, staging):
    """many contigs + small pushes: contigs are scanned / formatted / copied back while later ones are still
    being pushed (the streamed finalize path); result must not change"""
    contigs = tuple((("ctg%d" % i).encode(), 3000 + 700 * i) for i in range(12))
    names, lens, reads = bamutil.random_case(13, n_reads=5000, contigs=contigs)
    recs = bamutil.records(reads)
    for mode in (0, 1):
        exp = oracle.run(recs, names, lens, mode, 250)
        got = run_engine(names, lens, recs, mode, 250, staging_bytes=staging)
        check(got, exp, mode, f"streamed staging={staging} mode={mode}")


@pytest.mark.parametrize("config,scale,mode,window", [(5, 0.1, 0, 20000), (4, 0.1, 1, 20000), (2, 0.004, 0, 20000)])
def test_generator_configs_larger(oracle, config, scale, mode, window):
    """a few million reads per case: extreme pile-ups (config 5: depth > 10^4 on 400 amplicons), spliced RNA-seq
    (config 4) and WGS (config 2), streamed from host memory in 64 MiB pushes"""
    from highperformancengs_b200.gen import Generator
    g = Generator(config, scale=scale)
    recs = g.records()
    exp = oracle.run(recs, g.names, g.lens, mode, window)
    got = run_engine(g.names, g.lens, recs, mode, window, staging_bytes=64 << 20)
    check(got, exp, mode, f"config{config} scale={scale}")
    assert got["stats"]["records"] == g.n_reads


@pytest.mark.parametrize("fmt", ["fast", "generic"])
def test_nine_digit_positions_and_name_lengths(oracle, monkeypatch, fmt):
    """bedGraph formatting across the 8 -> 9 digit boundary of positions, with "name\\t" of exactly eight bytes (the
    register-assembled fast path), nine bytes and longer (the generic producer path), for both K5 variants"""
    if fmt == "generic":
        monkeypatch.setenv("NGSCOV_FMT", "generic")
    names = [b"chr1", b"sevench", b"eightchr", b"a_rather_long_contig_name_42"]
    lens = [100001000, 5000, 5000, 3000]
    reads = []
    for p in list(range(99999700, 100000400, 7)) + [0, 1, 9, 10, 99, 100, 999, 1000, 99999, 100000, 9999999, 10000000]:
        reads.append((0, p, 0, "150M"))
    reads.sort(key=lambda r: r[1])
    for c in (1, 2, 3):
        reads += [(c, p, 16 if p % 3 else 0, "37M5D20M") for p in range(0, 2500, 11)]
    recs = bamutil.records(reads, seq=False)
    exp = oracle.run(recs, names, lens, 0, 20000)
    got = run_engine(names, lens, recs, 0, 20000)
    check(got, exp, 0, f"nine digits fmt={fmt}")
    assert b"\t100000" in got["bedgraph"] and b"\t99999" in got["bedgraph"]


@pytest.mark.parametrize("decoders", [1, 2])
@pytest.mark.parametrize("staging", [0, 300000])
@pytest.mark.parametrize("case", ["random", "c2", "c4", "c5", "stored"])
def test_device_bgzf_inflate(oracle, tmp_path, monkeypatch, case, staging, decoders):
    """ngscov_push_bgzf: the BAM file's compressed blocks are inflated on the device (one warp per block) and feed the
    same pipeline; outputs must equal the oracle's on the host-inflated stream.  `stored` = level-0 BGZF (stored
    deflate blocks), `random` = zlib level 6 from Python (dynamic blocks), generator files = level 1.  decoders = 2:
    the opt-in kernel that runs two decoders per warp as a warp-wide state machine (NGSCOV_INF_S=2)."""
    monkeypatch.setenv("NGSCOV_INF_S", str(decoders))
    from highperformancengs_b200 import CoverageEngine, bamio
    from highperformancengs_b200.gen import Generator
    bam = str(tmp_path / "x.bam")
    mode, window = 0, 500
    if case in ("random", "stored"):
        names, lens, reads = bamutil.random_case(31, n_reads=6000)
        recs = bamutil.records(reads)
        if case == "stored":
            import zlib, struct
            raw = bamutil.header_bytes(names, lens) + bytes(recs)
            out = b""
            for o in range(0, len(raw), 60000):
                d = raw[o:o + 60000]
                c = zlib.compressobj(0, zlib.DEFLATED, -15)
                cd = c.compress(d) + c.flush()
                out += struct.pack("<BBBBIBBHBBHH", 31, 139, 8, 4, 0, 0, 255, 6, 66, 67, 2, len(cd) + 25) + cd + struct.pack("<II", zlib.crc32(d), len(d))
            open(bam, "wb").write(out)
        else:
            bamutil.write_bam(bam, names, lens, recs, index=False)
    else:
        cfgn, scale = {"c2": (2, 0.004), "c4": (4, 0.02), "c5": (5, 0.004)}[case]
        g = Generator(cfgn, scale=scale)
        g.write_bam(bam)
        names, lens, recs = g.names, g.lens, g.records()
        mode, window = (1, 100) if cfgn == 4 else (0, 20000)
    data = open(bam, "rb").read()
    inf = bamio.bgzf_inflate(data)
    _, _, hdr_len = bamio.parse_header(inf)
    assert bytes(inf[hdr_len:]) == bytes(recs)
    exp = oracle.run(recs, names, lens, mode, window)
    with CoverageEngine(names, lens, mode=mode, window=window, emit_wig=True, staging_bytes=staging) as eng:
        eng.push_bgzf(data, skip=hdr_len, last=True)
        eng.finish()
        got = {"bedgraph": eng.output(0) if mode == 0 else b"", "depth": eng.output(1) if mode == 0 else b"", "wig": eng.output(2), "chromsize": eng.output(3)}
        got["stats"] = eng.stats()
    check(got, exp, mode, f"device inflate {case} staging={staging}")
    assert got["stats"]["ms_inflate"] > 0


@pytest.mark.parametrize("decoders", [1, 2])
def test_device_bgzf_inflate_rejects_corrupt_block(tmp_path, monkeypatch, decoders):
    from highperformancengs_b200 import CoverageEngine, NgscovError, bamio
    monkeypatch.setenv("NGSCOV_INF_S", str(decoders))
    names, lens, reads = bamutil.random_case(32, n_reads=3000)
    bam = str(tmp_path / "x.bam")
    bamutil.write_bam(bam, names, lens, bamutil.records(reads), index=False)
    data = bytearray(open(bam, "rb").read())
    _, _, hdr_len = bamio.parse_header(bamio.bgzf_inflate(bytes(data)))
    for k in range(40, 60):
        data[len(data) // 2 + k] ^= 0x5A          # damage the deflate stream of a block in the middle
    with pytest.raises(NgscovError) as e:
        with CoverageEngine(names, lens, mode=0, window=500) as eng:
            eng.push_bgzf(bytes(data), skip=hdr_len, last=True)
            eng.finish()
    assert e.value.code == -4


def test_large_scale_properties(tmp_path):
    """Size-independent properties at a size the oracle cannot reach (config 2, NGSCOV_TEST_SCALE of the 30x genome,
    default 0.1 = 62 M reads; 1.0 = the full benchmark input): the runs are sorted, disjoint, maximal, non-zero and
    inside their contigs; the two accumulators agree (sum of depth over runs == sum of the window bins, per window);
    the bedGraph text has exactly one line per run and exactly the length the runs' digit counts predict; its first and
    last lines are the first and last runs; and the same bytes come out when the input is the compressed BAM inflated
    on the device."""
    import zlib
    import torch
    from highperformancengs_b200 import CoverageEngine, bamio
    from highperformancengs_b200.gen import Generator
    scale = float(os.environ.get("NGSCOV_TEST_SCALE", "0.1"))
    W = 20000
    g = Generator(2, scale=scale, seq=False)
    recs = g.records()
    dev = torch.from_numpy(recs).cuda()
    torch.cuda.synchronize()
    lens = np.asarray(g.lens, dtype=np.int64)
    with CoverageEngine(g.names, g.lens, mode=0, window=W) as eng:
        eng.push_device(dev.data_ptr(), dev.numel(), last=True)
        eng.finish()
        runs = eng.runs_array()
        bins = eng.bins().astype(np.int64)
        bed = np.array(eng.output_view(0), copy=True)
        st = eng.stats()
    del dev
    c, s, e, d = runs["contig"].astype(np.int64), runs["start"].astype(np.int64), runs["end"].astype(np.int64), runs["depth"].astype(np.int64)
    assert st["records"] == g.n_reads and len(runs) == st["runs"] and len(runs) > 0
    # structure of the run list
    assert (e > s).all() and (d > 0).all() and (c >= 0).all() and (e <= lens[c]).all()
    same = c[1:] == c[:-1]
    assert (c[1:] >= c[:-1]).all() and (s[1:][same] >= e[:-1][same]).all()
    abut = same & (s[1:] == e[:-1])
    assert (d[1:][abut] != d[:-1][abut]).all(), "two abutting runs of equal depth: not maximal"
    # the window bins, recomputed from the runs (variant A: exact sum of depth over the bases of each window)
    n_win = lens // W + 1
    win0 = np.concatenate([[0], np.cumsum(n_win)])[:-1]
    k0, k1 = s // W, (e - 1) // W
    exp_bins = np.zeros(int(n_win.sum()), dtype=np.int64)
    one = k0 == k1
    np.add.at(exp_bins, win0[c[one]] + k0[one], (e[one] - s[one]) * d[one])
    for ci, si, ei, di, a, b in zip(c[~one], s[~one], e[~one], d[~one], k0[~one], k1[~one]):      # runs crossing a window edge: few
        for k in range(a, b + 1):
            lo, hi = max(si, k * W), min(ei, (k + 1) * W)
            exp_bins[win0[ci] + k] += (hi - lo) * di
    assert int(bins.sum()) == int(((e - s) * d).sum())
    assert (bins == exp_bins).all()
    # the text: one line per run, and exactly as long as the digit counts say
    p10 = 10 ** np.arange(1, 11, dtype=np.int64)
    dig = lambda v: np.searchsorted(p10, v, side="right") + 1
    name_len = np.asarray([len(n) for n in g.names], dtype=np.int64)
    assert int((name_len[c] + dig(s) + dig(e) + dig(d) + 4).sum()) == bed.size
    assert int((bed == 10).sum()) == len(runs)
    head = bytes(bed[:200000]).split(b"\n")[:1000]
    tail = bytes(bed[-200000:]).split(b"\n")[-1001:-1]
    fmt = lambda i: b"%s\t%d\t%d\t%d" % (g.names[c[i]], s[i], e[i], d[i])
    assert head == [fmt(i) for i in range(1000)]
    assert tail == [fmt(i) for i in range(len(runs) - 1000, len(runs))]
    # the same bytes from the BAM file's compressed blocks (device inflate)
    ref_sum = zlib.adler32(bed)
    bam = str(tmp_path / "big.bam")
    g.write_bam(bam)
    raw = np.fromfile(bam, dtype=np.uint8)
    inf, hdr_len = b"", None
    for off, bs in bamio.bgzf_blocks(memoryview(raw[:8 << 20])):
        inf += zlib.decompress(bytes(raw[off + 18:off + bs - 8]), -15)
        try:
            hdr_len = bamio.parse_header(inf)[2]
            break
        except Exception:
            continue
    with CoverageEngine(g.names, g.lens, mode=0, window=W) as eng:
        eng.push_bgzf(raw, skip=hdr_len, last=True)
        eng.finish()
        assert zlib.adler32(np.array(eng.output_view(0), copy=False)) == ref_sum


def test_unsorted_input_takes_the_red_path(oracle):
    """records that are not coordinate sorted: the owner tiles cannot find them by key; the library notices (sort keys are
    checked while the records are framed) and falls back to REDs + the classic scan -- same output as the oracle, which
    groups by contig"""
    import random
    names, lens, reads = bamutil.random_case(21, n_reads=3000)
    rnd = random.Random(3)
    rnd.shuffle(reads)
    recs = bamutil.records(reads, seq=False)
    for mode, _ in MODES:
        exp = oracle.run(recs, names, lens, mode, 400)
        got = run_engine(names, lens, recs, mode, 400)
        check(got, exp, mode, "unsorted")
        got = run_engine(names, lens, recs, mode, 400, staging_bytes=16384)     # detected in a later chunk: catch-up pass
        check(got, exp, mode, "unsorted, small pushes")


def test_long_and_spliced_reads_far_deltas(oracle):
    """reads longer than the owner's view (NGS_NEAR_H = 512), pairs with large gaps, many-op CIGARs: far deltas go through
    REDs + the far flags, CHAIN entries are decoded from the stream"""
    import random
    rnd = random.Random(11)
    names, lens = [b"chrL", b"chrM"], [400000, 90000]
    reads = []
    for c, L in enumerate(lens):
        p = 0
        while p < L - 30000:
            p += rnd.randrange(1, 400)
            k = rnd.randrange(6)
            if k == 0:
                cig = "%dM" % rnd.randrange(600, 9000)                       # far end
            elif k == 1:
                cig = "%dM%dN%dM" % (rnd.randrange(20, 200), rnd.randrange(1, 60), rnd.randrange(20, 200))     # PAIR entry
            elif k == 2:
                cig = "%dM%dN%dM" % (rnd.randrange(20, 200), rnd.randrange(500, 20000), rnd.randrange(20, 200))   # CHAIN
            elif k == 3:
                cig = "5S%dM2I%dM3D%dM%dN%dM4S" % tuple(rnd.randrange(10, 120) for _ in range(5))
            elif k == 4:
                cig = "%dM" % rnd.randrange(30, 300)
            else:
                cig = "%dM%dD%dM" % (rnd.randrange(10, 100), rnd.randrange(1, 30), rnd.randrange(10, 4090))
            reads.append((c, p, rnd.choice([0, 16, 0x400, 0x100]), cig))
    recs = bamutil.records(reads, seq=False)
    for mode, _ in MODES:
        for window in (300, 20000):
            exp = oracle.run(recs, names, lens, mode, window)
            got = run_engine(names, lens, recs, mode, window)
            check(got, exp, mode, f"far deltas mode={mode} w={window}")
            got = run_engine(names, lens, recs, mode, window, staging_bytes=32768)
            check(got, exp, mode, f"far deltas, streamed, mode={mode} w={window}")


def test_classic_pipeline_still_matches(oracle, monkeypatch):
    """NGSCOV_PIPELINE=classic: K1 framing + K2 REDs + scan + format kernels (the round-1 organisation), kept as the
    fallback for unsorted input and as a second implementation to compare with"""
    monkeypatch.setenv("NGSCOV_PIPELINE", "classic")
    names, lens, reads = bamutil.random_case(4)
    recs = bamutil.records(reads)
    for mode, _ in MODES:
        exp = oracle.run(recs, names, lens, mode, 500)
        got = run_engine(names, lens, recs, mode, 500)
        check(got, exp, mode, "classic")


def _dense_case(seed, L, cov, names=(b"chrD",)):
    """uniform random short reads at `cov`x coverage: ~2 * cov * 4096 / read_len run boundaries per owner tile"""
    rng = np.random.RandomState(seed)
    reads = []
    for c in range(len(names)):
        n = int(L * cov / 40)
        pos = np.sort(rng.randint(0, L - 60, size=n))
        ln = rng.randint(20, 60, size=n)
        reads += [(c, int(p), 0, "%dM" % int(l)) for p, l in zip(pos, ln)]
    return list(names), [L] * len(names), reads


def test_dense_tiles_take_several_text_batches(oracle):
    """more run boundaries in an owner tile than one text batch holds (1280): the length pre-pass and the batch loop of
    k_cover; a tile with ~3000 boundaries needs three batches"""
    names, lens, reads = _dense_case(5, 60000, 40.0)
    recs = bamutil.records(reads, seq=False)
    for mode, _ in MODES:
        exp = oracle.run(recs, names, lens, mode, 1000)
        got = run_engine(names, lens, recs, mode, 1000)
        check(got, exp, mode, "dense tiles")
        if mode == 0:
            assert exp["bedgraph"].count(b"\n") > 2 * 1280 * (60000 // 4096)


def test_lines_longer_than_the_staging_area(oracle):
    """a contig name of 250 characters: the lines of a batch do not fit the shared-memory staging area and go straight
    to global memory (byte-wise writer), also in tiles with several batches"""
    names, lens, reads = _dense_case(6, 30000, 30.0, names=(b"N" * 250, b"short"))
    recs = bamutil.records(reads, seq=False)
    exp = oracle.run(recs, names, lens, 0, 500)
    got = run_engine(names, lens, recs, 0, 500)
    check(got, exp, 0, "long contig name")


@pytest.mark.parametrize("staging", [16384, 100000])
def test_streamed_finalize_inside_contigs(oracle, monkeypatch, staging):
    """tile-granular streamed finalize: with small pushes and NGSCOV_STREAM_MIN_TILES=1 k_cover is launched over a few
    tiles at a time in the middle of contigs; depth carry and last boundary travel from launch to launch through the
    look-back word of the previous launch's last tile"""
    monkeypatch.setenv("NGSCOV_STREAM_MIN_TILES", "1")
    names, lens, reads = bamutil.random_case(23, n_reads=9000, contigs=((b"c1", 150000), (b"c2", 9000), (b"c3", 300000)))
    reads = [r for r in reads if r[0] < 0 or r[3] != "50M1000N100M"] + [(2, 1000, 0, "100M250000N50M")]      # plus one far delta across many tiles
    reads.sort(key=lambda r: (r[0] if r[0] >= 0 else 1 << 30, r[1]))
    recs = bamutil.records(reads, seq=False)
    for mode, _ in MODES:
        exp = oracle.run(recs, names, lens, mode, 777)
        got = run_engine(names, lens, recs, mode, 777, staging_bytes=staging)
        check(got, exp, mode, f"streamed inside contigs, staging={staging}")
    # long stretches without coverage between launches
    names, lens = [b"g1"], [2000000]
    reads = [(0, 100, 0, "150M"), (0, 120, 0, "50M"), (0, 900000, 0, "150M"), (0, 900050, 16, "150M"), (0, 1999000, 0, "100M")]
    reads = sorted(reads + [(0, 400000 + 37 * i, 0, "30M") for i in range(2000)], key=lambda r: r[1])
    recs = bamutil.records(reads, seq=False)
    exp = oracle.run(recs, names, lens, 0, 20000)
    got = run_engine(names, lens, recs, 0, 20000, staging_bytes=8192)
    check(got, exp, 0, "streamed with gaps")
