"""bam_sliding_count path (NGSCOV_MODE_SLIDE).
CPU: the oracle (oracle/ngscov_slide_oracle.c) against the committed outputs of the unmodified reference binary
(tests/golden/slide_outputs.json, made by tests/golden/make_slide_golden.py) and, where oracle/_ref exists, against the
binary itself.  GPU: the engine through the C ABI and the drop-in C driver against the same outputs."""
import hashlib
import json
import os
import subprocess

import numpy as np
import pytest

from tests import bamutil, slide_cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "slide_outputs.json")))
RGOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "slide_region_outputs.json")))
PARAMS = [(n, k, p, w) for n, k, p, ws in slide_cases.CASES for w in ws]
_BY_NAME = {c[0]: c for c in slide_cases.CASES}
RPARAMS = [(n, _BY_NAME[n][1], _BY_NAME[n][2], w, r, parsed) for n, w, r, parsed in slide_cases.REGION_CASES]
RIDS = [f"{p[0]}-{p[3]}-{p[4]}" for p in RPARAMS]
IDS = [f"{p[0]}-{p[3]}" for p in PARAMS]
_cache = {}


def _input(name, kind, params):
    if name not in _cache:
        _cache.clear()
        _cache[name] = slide_cases.case_input(kind, params)
    return _cache[name]


def check_gold(txt, n_count, key):
    g = GOLD[key]
    assert len(txt) == g["size"], f"{key}: size {len(txt)} != reference {g['size']}"
    assert hashlib.md5(txt).hexdigest() == g["md5"], f"{key}: differs from the reference output"
    if "text" in g:
        assert txt.decode() == g["text"]
    assert n_count == g["n_count"]


def test_gold_covers_every_case():
    assert set(GOLD) == {f"{n}|{w}" for n, k, p, w in PARAMS}


@pytest.mark.parametrize("name,kind,params,window", PARAMS, ids=IDS)
def test_oracle_matches_reference_outputs(oracle, name, kind, params, window):
    names, lens, recs, g = _input(name, kind, params)
    got = oracle.slide(recs, names, lens, window)
    check_gold(got["txt"], got["n_count"], f"{name}|{window}")


@pytest.mark.parametrize("name,kind,params,window,region,parsed", RPARAMS, ids=RIDS)
def test_oracle_region_mode_matches_reference_outputs(oracle, name, kind, params, window, region, parsed):
    """`-r`: only what bam_fetch hands out is counted (bam_sliding_count.c:393-409)"""
    names, lens, recs, g = _input(name, kind, params)
    got = oracle.slide(recs, names, lens, window, region=parsed)
    g = RGOLD[f"{name}|{window}|{region}"]
    assert (len(got["txt"]), hashlib.md5(got["txt"]).hexdigest(), got["n_count"]) == (g["size"], g["md5"], g["n_count"])


def test_region_gold_covers_every_case():
    assert set(RGOLD) == {f"{p[0]}|{p[3]}|{p[4]}" for p in RPARAMS}


def test_oracle_vs_reference_binary(oracle, tmp_path):
    ref = os.path.join(ROOT, "oracle", "_ref", "bam_sliding_count")
    if not os.path.exists(ref):
        pytest.skip("oracle/_ref/bam_sliding_count not built (needs /root/reference)")
    names, lens, reads = bamutil.random_case(12, n_reads=2500)
    recs = bamutil.records(reads, seq=True, seq_seed=5)
    bam = str(tmp_path / "t.bam")
    bamutil.write_bam(bam, names, lens, recs, index=False)
    for w in (500, 20000):
        subprocess.check_call([ref, "-o", str(tmp_path / "o"), "-w", str(w), bam], stderr=subprocess.DEVNULL, cwd=str(tmp_path))
        assert open(str(tmp_path / "o.txt"), "rb").read() == oracle.slide(recs, names, lens, w)["txt"]


EDGE_READS = [(0, 100, 0, "50M"), (0, 900, 0, "100M"), (0, 901, 0, "100M"), (0, 950, 16, "10M2000N40M"), (0, 990, 0, "5S10M"), (0, 999, 0, "*"),
              (0, 1000, 0, "*"), (0, 1000, 0, "20M"), (0, 1001, 4, "20M"), (0, 1200, 0, "70000M"), (0, 1499, 0, "30M"), (0, 1500, 0, "30M"),
              (0, 400, 0, "66000M"), (1, 1100, 0, "50M")]


def test_oracle_region_edges_vs_reference_binary(oracle, tmp_path):
    """reads that start before the region and reach into it, end exactly at its first base, have no CIGAR"""
    ref = os.path.join(ROOT, "oracle", "_ref", "bam_sliding_count")
    if not os.path.exists(ref):
        pytest.skip("oracle/_ref/bam_sliding_count not built (needs /root/reference)")
    names, lens = [b"a", b"b"], [200000, 5000]
    recs = bamutil.records(sorted(EDGE_READS, key=lambda r: (r[0], r[1])), seq=True, seq_seed=3)
    bam = str(tmp_path / "t.bam")
    bamutil.write_bam(bam, names, lens, recs, index=True)
    for region, parsed in [("a:1001-1500", (0, 1000, 1500)), ("a:1000-1001", (0, 999, 1001)), ("a", (0, 0, 1 << 29)), ("b:2001-3000", (1, 2000, 3000)),
                           ("a:951-951", (0, 950, 951)), ("a:70401", (0, 70400, 1 << 29))]:
        r = subprocess.run([ref, "-o", str(tmp_path / "o"), "-w", "100", "-r", region, bam], capture_output=True, text=True, cwd=str(tmp_path))
        assert r.returncode == 0
        exp = oracle.slide(recs, names, lens, 100, region=parsed)
        assert open(str(tmp_path / "o.txt"), "rb").read() == exp["txt"], region
        assert int([ln for ln in r.stderr.splitlines() if ln.strip().isdigit()][0]) == exp["n_count"], region


def run_engine(names, lens, recs, window, staging_bytes=0, region=None):
    from highperformancengs_b200 import CoverageEngine
    with CoverageEngine(names, lens, mode=2, window=window, staging_bytes=staging_bytes) as eng:
        if region is not None:
            eng.set_region(*region)
        eng.push_host(recs, last=True)
        eng.finish()
        return eng.output(4), eng.stats()


@pytest.mark.gpu
@pytest.mark.parametrize("name,kind,params,window", PARAMS, ids=IDS)
def test_engine_matches_reference_outputs(name, kind, params, window):
    names, lens, recs, g = _input(name, kind, params)
    txt, st = run_engine(names, lens, recs, window)
    check_gold(txt, st["kept"], f"{name}|{window}")


@pytest.mark.gpu
@pytest.mark.parametrize("name,kind,params,window,region,parsed", RPARAMS, ids=RIDS)
def test_engine_region_mode_matches_reference_outputs(name, kind, params, window, region, parsed):
    names, lens, recs, g = _input(name, kind, params)
    txt, st = run_engine(names, lens, recs, window, region=parsed)
    g = RGOLD[f"{name}|{window}|{region}"]
    assert (len(txt), hashlib.md5(txt).hexdigest(), st["kept"]) == (g["size"], g["md5"], g["n_count"])


@pytest.mark.gpu
def test_engine_region_mode_edges(oracle):
    """reads that start before the region and reach into it, end exactly at its first base, have no CIGAR, are too long
    for a table entry (sequential path), and set_region after a push"""
    from highperformancengs_b200 import CoverageEngine, NgscovError
    names, lens = [b"a", b"b"], [200000, 5000]
    recs = bamutil.records(sorted(EDGE_READS, key=lambda r: (r[0], r[1])), seq=True, seq_seed=3)
    for region in [(0, 1000, 1500), (0, 999, 1001), (0, 0, 1 << 29), (1, 0, 1 << 29), (1, 2000, 3000), (0, 950, 951), (0, 70400, 1 << 29)]:
        exp = oracle.slide(recs, names, lens, 100, region=region)
        txt, st = run_engine(names, lens, recs, 100, region=region)
        assert txt == exp["txt"] and st["kept"] == exp["n_count"], region
    with CoverageEngine(names, lens, mode=2, window=100) as eng:
        eng.push_host(recs, last=False)
        with pytest.raises(NgscovError):
            eng.set_region(0, 0, 10)
    with CoverageEngine(names, lens, mode=0, window=100) as eng:
        with pytest.raises(NgscovError):
            eng.set_region(0, 0, 10)


@pytest.mark.gpu
@pytest.mark.parametrize("staging", [8192, 70000])
def test_engine_streamed_pushes_and_long_reads(oracle, staging):
    """records cut by push boundaries (SEQ must be complete before a record is counted), reads too long for a table
    entry (l_qseq >= 65535: the sequential path), a read at POS -1, unmapped reads in between"""
    names, lens, reads = bamutil.random_case(17, n_reads=3000)
    reads = reads[:-20] + [(2, 100, 0, "70000M"), (2, 200, 16, "66000M5S")] + reads[-20:]
    reads.sort(key=lambda r: (r[0] if r[0] >= 0 else 1 << 30, r[1]))
    recs = bamutil.records(reads, seq=True, seq_seed=9)
    for w in (300, 20000):
        exp = oracle.slide(recs, names, lens, w)
        txt, st = run_engine(names, lens, recs, w, staging_bytes=staging if staging > 70000 else 0)
        assert txt == exp["txt"] and st["kept"] == exp["n_count"]
    exp = oracle.slide(recs, names, lens, 500)
    txt, st = run_engine(names, lens, recs, 500, staging_bytes=262144)
    assert txt == exp["txt"]


@pytest.mark.gpu
def test_c_driver_bam_sliding_count(tmp_path):
    exe = os.path.join(ROOT, "highperformancengs_b200", "host", "bam_sliding_count")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "highperformancengs_b200", "host")])
    name, kind, params, windows = slide_cases.CASES[2]
    names, lens, recs, g = slide_cases.case_input(kind, params)
    bam = str(tmp_path / "in.bam")
    bamutil.write_bam(bam, names, lens, recs, index=False)
    for w in windows:
        r = subprocess.run([exe, "-o", "o", "-w", str(w), bam], cwd=str(tmp_path), capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        n_count = int([ln for ln in r.stderr.splitlines() if ln.strip().isdigit()][0])
        check_gold(open(str(tmp_path / "o.txt"), "rb").read(), n_count, f"{name}|{w}")
    # -r: the reference insists on the index, prints the parsed region on stdout and counts what bam_fetch hands out
    r = subprocess.run([exe, "-o", "o", "-r", "chr1:1-100", bam], cwd=str(tmp_path), capture_output=True, text=True)
    assert r.returncode == 1 and "BAM indexing file is not available" in r.stderr
    bamutil.write_bam(bam, names, lens, recs, index=True)
    rn, rw, region, parsed = [c for c in slide_cases.REGION_CASES if c[0] == name][0]
    r = subprocess.run([exe, "-o", "o", "-w", str(rw), "-r", region, bam], cwd=str(tmp_path), capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    g = RGOLD[f"{name}|{rw}|{region}"]
    txt = open(str(tmp_path / "o.txt"), "rb").read()
    assert r.stdout == g["stdout"] and hashlib.md5(txt).hexdigest() == g["md5"]
    assert int([ln for ln in r.stderr.splitlines() if ln.strip().isdigit()][0]) == g["n_count"]
    r = subprocess.run([exe, "-o", "o", "-r", "nosuchcontig:5-9", bam], cwd=str(tmp_path), capture_output=True, text=True)
    assert r.returncode == 1 and "Invalid region" in r.stderr


def test_region_strings_parse_like_samtools(tmp_path):
    """parse_region of the C driver against what samtools-0.1.19 bam_parse_region returned for the same strings
    (table recorded from oracle/_ref/libbam.a; refID -1 = "Invalid region")"""
    table = [("c1", 0, 0, 1 << 29), ("c1:1000-5000", 0, 999, 5000), ("c1:2,001-2,500", 0, 2000, 2500), ("chr2:69000", 1, 68999, 1 << 29),
             (" chr2 : 10 - 20 ", 1, 9, 20), ("HLA:01-2", 2, 0, 1 << 29), ("HLA:01-2:5-9", 2, 4, 9), ("x:y", 3, 0, 1 << 29), ("x:y:7", 3, 6, 1 << 29),
             ("nope:1-2", -1, None, None), ("c1:", 0, 0, 1 << 29), ("c1:-5", 0, -5, 5), ("c1:5-", 0, 4, 0), ("c1:1-2-3", -1, None, None),
             ("c1:0-10", 0, 0, 10), ("c1:abc", -1, None, None)]
    src = tmp_path / "h.c"
    cli = os.path.join(ROOT, "highperformancengs_b200", "host", "ngscov_slide_cli.c")
    src.write_text('#include <stdio.h>\n#define main slide_main\n#include "%s"\n#undef main\n' % cli +
                   'int main(int argc, char **argv) { char *names[] = { "c1", "chr2", "HLA:01-2", "x:y" }; int b, e;\n'
                   '  for (int i = 1; i < argc; ++i) { int r = parse_region(names, 4, argv[i], &b, &e); printf("%d %d %d\\n", r, b, e); } return 0; }\n')
    exe = str(tmp_path / "h")
    lib = os.path.join(ROOT, "highperformancengs_b200")
    subprocess.check_call(["gcc", "-O1", "-w", str(src), "-o", exe, "-L" + lib, "-lngscov_cuda", "-Wl,-rpath," + lib, "-lz"])
    out = subprocess.run([exe] + [t[0] for t in table], capture_output=True, text=True, check=True).stdout.split("\n")
    for (q, r, b, e), line in zip(table, out):
        gr, gb, ge = (int(x) for x in line.split())
        assert gr == r, q
        if r >= 0:
            assert (gb, ge) == (b, e), q
