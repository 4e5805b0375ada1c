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
PARAMS = [(n, k, p, w) for n, k, p, ws in slide_cases.CASES for w in ws]
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


def run_engine(names, lens, recs, window, staging_bytes=0):
    from highperformancengs_b200 import CoverageEngine
    with CoverageEngine(names, lens, mode=2, window=window, staging_bytes=staging_bytes) as eng:
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
    r = subprocess.run([exe, "-o", "o", "-r", "chr1:1-100", bam], cwd=str(tmp_path), capture_output=True, text=True)
    assert r.returncode == 1 and "not supported" in r.stderr
