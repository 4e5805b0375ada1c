"""CPU: host-side logic -- the C-ABI library loads and exports every symbol of include/ngscov.h, the
kernels' record-level logic (csrc/ngscov_hd.h) reproduces the oracle through a sequential model of the
kernel orchestration, CLI argument handling mirrors the reference."""
import ctypes
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from highperformancengs_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "ngscov.h")).read()
    declared = set(re.findall(r"\b(ngscov_[a-z_0-9]+)\s*\(", hdr))
    declared -= {"ngscov_t"}
    assert declared, "no declarations parsed"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for sym in sorted(declared):
        assert hasattr(lib, sym), f"{sym} declared in include/ngscov.h but not exported"
    assert declared == set(_lib.SYMBOLS), "ctypes binding table and header disagree"
    bound = _lib.load_library()
    assert bound.ngscov_abi_version() == 1
    assert bound.ngscov_strerror(-4) == b"malformed or truncated BAM record stream"


def test_no_gpu_means_loud_failure(have_gpu):
    if have_gpu:
        pytest.skip("a GPU is present")
    from highperformancengs_b200 import CoverageEngine, NgscovError
    with pytest.raises(NgscovError) as e:
        CoverageEngine([b"c"], [100])
    assert e.value.code == -2          # NGSCOV_ECUDA: there is no CPU fallback


@pytest.fixture(scope="module")
def hd_check(oracle):
    out = os.path.join(ROOT, "tests", "_build")
    os.makedirs(out, exist_ok=True)
    exe = os.path.join(out, "hd_check")
    subprocess.check_call(["gcc", "-O2", "-c", os.path.join(ROOT, "oracle", "ngscov_oracle.c"), "-o", os.path.join(out, "oracle.o")])
    subprocess.check_call(["g++", "-O2", "-std=c++17", os.path.join(ROOT, "tests", "hd_check.cpp"), os.path.join(out, "oracle.o"), "-o", exe, "-lz", "-lm"])
    return exe


@pytest.mark.parametrize("case,mode,window,sub,chunk", [
    ("random", "depth", 150, 64, 1000), ("random", "wig", 999, 256, 70000), ("random", "depth", 20000, 1024, 10 ** 9),
    ("c2", "depth", 500, 512, 200000), ("c4", "wig", 100, 1024, 10 ** 9), ("c5", "depth", 1000, 2048, 3000000)])
def test_kernel_logic_model_matches_oracle(hd_check, tmp_path, case, mode, window, sub, chunk):
    from tests import bamutil
    bam = str(tmp_path / "x.bam")
    if case == "random":
        names, lens, reads = bamutil.random_case(21, n_reads=2500)
        bamutil.write_bam(bam, names, lens, bamutil.records(reads), index=False)
    else:
        from highperformancengs_b200.gen import Generator
        Generator({"c2": 2, "c4": 4, "c5": 5}[case], scale={"c2": 0.0003, "c4": 0.004, "c5": 0.001}[case]).write_bam(bam)
    r = subprocess.run([hd_check, bam, mode, str(window), str(sub), str(chunk)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


def test_register_line_assembly_known_answers(hd_check):
    """ngs_lw / ngs_num_conv (the K5 fast path) against snprintf: all digit counts, name lengths, alignments"""
    r = subprocess.run([hd_check, "--selftest-lines"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


@pytest.fixture(scope="module")
def inflate_check():
    out = os.path.join(ROOT, "tests", "_build")
    os.makedirs(out, exist_ok=True)
    exe = os.path.join(out, "inflate_check")
    subprocess.check_call(["g++", "-O2", "-std=c++17", os.path.join(ROOT, "tests", "inflate_check.cpp"), "-o", exe, "-lz"])
    return exe


def test_deflate_core_matches_zlib_on_synthetic_streams(inflate_check):
    """csrc/ngscov_inflate.h (what k_bgzf_inflate runs per warp) against zlib: stored / fixed / dynamic blocks, every
    level and strategy, long codes, overlapping copies, truncated and corrupt streams"""
    r = subprocess.run([inflate_check], capture_output=True, text=True)
    assert r.returncode == 0 and " ok" in r.stdout, r.stdout + r.stderr


@pytest.mark.parametrize("config,scale,seq", [(2, 0.001, False), (4, 0.01, True), (5, 0.002, True)])
def test_deflate_core_matches_zlib_on_bam_files(inflate_check, tmp_path, config, scale, seq):
    from highperformancengs_b200.gen import Generator
    bam = str(tmp_path / "x.bam")
    Generator(config, scale=scale, seq=seq).write_bam(bam)
    r = subprocess.run([inflate_check, bam], capture_output=True, text=True)
    assert r.returncode == 0 and " ok" in r.stdout, r.stdout + r.stderr


def test_python_cli_atoi_and_usage(capsys):
    from highperformancengs_b200 import cli
    assert cli._atoi("20000") == 20000 and cli._atoi(" 12abc") == 12 and cli._atoi("x") == 0 and cli._atoi("-5") == -5
    with pytest.raises(SystemExit) as e:
        cli.bam2depth(["bam2depth"])
    assert e.value.code == 1
    with pytest.raises(SystemExit) as e:
        cli.bam2wig(["bam2wig", "-h"])
    assert e.value.code == 1
    assert "Usage" in capsys.readouterr().err


def test_c_drivers_usage_and_missing_index(tmp_path):
    host = os.path.join(ROOT, "highperformancengs_b200", "host")
    for tool in ("bam2depth", "bam2wig"):
        exe = os.path.join(host, tool)
        if not os.path.exists(exe):
            pytest.skip("C drivers not built")
        r = subprocess.run([exe], capture_output=True, text=True)
        assert r.returncode == 1 and "Usage" in r.stderr            # reference: usage on stderr, exit(1)
        r = subprocess.run([exe, "-h"], capture_output=True, text=True)
        assert r.returncode == 1
        from tests import bamutil
        names, lens, reads = bamutil.toy_case()
        bam = str(tmp_path / "noidx.bam")
        bamutil.write_bam(bam, names, lens, bamutil.records(reads), index=False)
        os.remove(bam + ".bai") if os.path.exists(bam + ".bai") else None
        r = subprocess.run([exe, "-o", str(tmp_path / "o"), bam], capture_output=True, text=True, cwd=str(tmp_path))
        assert r.returncode == 1 and "BAM indexing file is not available" in r.stderr   # reference bam2depth.c:112-119
        r = subprocess.run([exe, "-o", str(tmp_path / "o"), str(tmp_path / "missing.bam")], capture_output=True, text=True, cwd=str(tmp_path))
        assert r.returncode == 1 and "Fail to open BAM file" in r.stderr


def test_file_hash_tool_matches_the_bench_definition(tmp_path):
    """tools/file_hash.py (CPU, used to pin device hashes against files the reference wrote) = the rule bench.py states:
    sum_j byte_j * ((j + 1) * 0x9E3779B97F4A7C15 mod 2^64) mod 2^64"""
    import sys
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from file_hash import file_hash
    b = np.random.RandomState(4).randint(0, 256, size=300007, dtype=np.uint8)
    p = str(tmp_path / "x.bin")
    b.tofile(p)
    M, mask = 0x9E3779B97F4A7C15, (1 << 64) - 1
    exp = 0
    for j in range(0, b.size, 997):               # a strided sample of terms in Python integers, the rest by linearity
        exp = (exp + int(b[j]) * (((j + 1) * M) & mask)) & mask
    sub = np.zeros_like(b)
    sub[::997] = b[::997]
    q = str(tmp_path / "y.bin")
    sub.tofile(q)
    assert file_hash(q, step=4096)["hash"] == "%016x" % exp
    whole = file_hash(p, step=65536)
    assert whole["bytes"] == b.size and whole["hash"] == file_hash(p, step=1 << 20)["hash"]
    gold = os.path.join(ROOT, "tests", "golden", "config2_full_reference.json")
    if os.path.exists(gold):
        import json
        g = json.load(open(gold))
        assert g["reads"] == 618271620 and len(g["bedgraph"]["hash"]) == 16 and g["bedgraph"]["bytes"] > 2e10


def test_full_size_gate_of_the_bench(tmp_path):
    """bench.py compares the device hash of the full-size config-2 output with the hash of the files the unmodified
    reference wrote (tests/golden/config2_full_reference.json).  The line saved from the B200 run must pass the gate, and
    a different hash must not."""
    import importlib.util
    import json
    spec = importlib.util.spec_from_file_location("ngscov_bench", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    gold = os.path.join(ROOT, "tests", "golden", "config2_full_reference.json")
    if not os.path.exists(gold):
        pytest.skip("golden not made yet")
    line = json.loads(open(os.path.join(ROOT, "profiles", "r2_bench_config2.json")).read().strip().splitlines()[-1])
    ph = line["parity"]["hash"]
    assert bench.full_size_gate(ph, line["config"]["reads"])["match"] is True
    bad = dict(ph, bedgraph="0" * 16)
    assert bench.full_size_gate(bad, line["config"]["reads"])["match"] is False
    assert bench.full_size_gate(ph, line["config"]["reads"], gold_path=str(tmp_path / "none.json")) is None
