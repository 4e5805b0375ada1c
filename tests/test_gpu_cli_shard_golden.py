"""GPU: (a) engine against the committed reference outputs (golden), (b) the C drop-in drivers end to end
on BAM files, (c) region sharding: slices processed separately and stitched equal the whole."""
import os
import subprocess

import numpy as np
import pytest

from tests import bamutil, cases
from tests.test_oracle_golden import check_against_gold, PARAMS, _input

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("name,kind,params,tool,window", PARAMS, ids=[f"{p[0]}-{p[3]}-{p[4]}" for p in PARAMS])
def test_engine_matches_reference_outputs(name, kind, params, tool, window):
    from highperformancengs_b200 import CoverageEngine
    names, lens, recs, g = _input(name, kind, params)
    mode = 0 if tool == "depth" else 1
    with CoverageEngine(names, lens, mode=mode, window=window, emit_wig=True) as eng:
        eng.push_host(recs, last=True)
        eng.finish()
        got = {k: eng.output(i) for i, k in enumerate(("bedgraph", "depth", "wig", "chromsize")) if mode == 0 or i >= 2}
    check_against_gold(got, f"{name}|{tool}|{window}")


@pytest.mark.parametrize("case,tool,window", [("random1", "depth", 150), ("random1", "wig", 150), ("c2_s0.0005", "depth", 20000),
                                              ("c4_s0.01", "wig", 100), ("toy", "depth", 10)])
def test_c_driver_writes_reference_identical_files(tmp_path, case, tool, window):
    exe = os.path.join(ROOT, "highperformancengs_b200", "host", "bam2depth" if tool == "depth" else "bam2wig")
    spec = [c for c in cases.CASES if c[0] == case][0]
    names, lens, recs, g = cases.case_input(spec[1], spec[2])
    bam = str(tmp_path / (case + ".bam"))
    if g is not None:
        g.write_bam(bam)
    else:
        bamutil.write_bam(bam, names, lens, recs)
    args = [exe, "-o", str(tmp_path / "o"), "-w", str(window)] + (["-W"] if tool == "depth" else []) + [bam]
    env = dict(os.environ, NGSCOV_STAGING_MB="1")           # small staging: records straddle pushes
    r = subprocess.run(args, cwd=str(tmp_path), capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stderr
    assert f"Converted {bam} to wig format at" in r.stderr and f"{names[0].decode()} at " in r.stderr
    files = {"wig": "o.1.wig", "chromsize": "o.1.chromSize.txt"}
    if tool == "depth":
        files.update(bedgraph=os.path.basename(bam) + ".1.bedGraph", depth="o.1.depth")
    got = {k: open(str(tmp_path / p), "rb").read() for k, p in files.items()}
    check_against_gold(got, f"{case}|{tool}|{window}")


@pytest.mark.parametrize("case,tool,window", [("random1", "depth", 150), ("random1", "wig", 150), ("c2_s0.0005", "depth", 20000),
                                              ("c4_s0.01", "wig", 100), ("c4_s0.01", "wig", 20000)])
@pytest.mark.parametrize("gpu_inflate", [0, 1])
@pytest.mark.parametrize("shards", [1, 2, 3, 7])
def test_c_driver_region_shards_from_bai(tmp_path, case, tool, window, shards, gpu_inflate):
    """NGSCOV_GPUS=N: every shard reads the BAM from the BAI linear-index offset of its slice, stops once the engine has
    seen a record behind the slice, and the stitched files equal the reference's (all shards on device 0 here)"""
    exe = os.path.join(ROOT, "highperformancengs_b200", "host", "bam2depth" if tool == "depth" else "bam2wig")
    spec = [c for c in cases.CASES if c[0] == case][0]
    names, lens, recs, g = cases.case_input(spec[1], spec[2])
    bam = str(tmp_path / (case + ".bam"))
    if g is not None:
        g.write_bam(bam)
    else:
        bamutil.write_bam(bam, names, lens, recs)
    args = [exe, "-o", str(tmp_path / "o"), "-w", str(window)] + (["-W"] if tool == "depth" else []) + [bam]
    env = dict(os.environ, NGSCOV_STAGING_MB="1", NGSCOV_GPUS=str(shards), NGSCOV_DEVICES=",".join(["0"] * shards), NGSCOV_VERBOSE="1",
               NGSCOV_GPU_INFLATE=str(gpu_inflate))
    r = subprocess.run(args, cwd=str(tmp_path), capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stderr
    assert shards == 1 or r.stderr.count("ngscov: shard ") >= 2, r.stderr
    if gpu_inflate and shards == 1:
        assert "device inflate 0.000 ms" not in r.stderr, r.stderr
    files = {"wig": "o.1.wig", "chromsize": "o.1.chromSize.txt"}
    if tool == "depth":
        files.update(bedgraph=os.path.basename(bam) + ".1.bedGraph", depth="o.1.depth")
    got = {k: open(str(tmp_path / p), "rb").read() for k, p in files.items()}
    check_against_gold(got, f"{case}|{tool}|{window}")


@pytest.mark.parametrize("gpu_inflate", ["0", "1"])
@pytest.mark.parametrize("case,tool,window", [("random1", "depth", 150), ("c4_s0.01", "wig", 100)])
def test_python_cli_mirror(tmp_path, monkeypatch, case, tool, window, gpu_inflate):
    """highperformancengs_b200.cli (ctypes mirror of the C drivers): same files as the reference, host or device inflate"""
    from highperformancengs_b200 import cli
    spec = [c for c in cases.CASES if c[0] == case][0]
    names, lens, recs, g = cases.case_input(spec[1], spec[2])
    bam = str(tmp_path / (case + ".bam"))
    if g is not None:
        g.write_bam(bam)
    else:
        bamutil.write_bam(bam, names, lens, recs)
    monkeypatch.chdir(tmp_path)
    monkeypatch.setenv("NGSCOV_GPU_INFLATE", gpu_inflate)
    argv = ["bam2depth" if tool == "depth" else "bam2wig", "-o", "o", "-w", str(window)] + (["-W"] if tool == "depth" else []) + [bam]
    assert (cli.bam2depth if tool == "depth" else cli.bam2wig)(argv) in (0, None)
    files = {"wig": "o.1.wig", "chromsize": "o.1.chromSize.txt"}
    if tool == "depth":
        files.update(bedgraph=os.path.basename(bam) + ".1.bedGraph", depth="o.1.depth")
    got = {k: open(str(tmp_path / p), "rb").read() for k, p in files.items()}
    check_against_gold(got, f"{case}|{tool}|{window}")


def _sharded(names, lens, recs, mode, window, n):
    """n engines on one GPU, each owning a genome slice but fed the whole stream (deltas are clipped)."""
    from highperformancengs_b200 import CoverageEngine, shard
    slices = shard.plan_slices(lens, n)
    engs = [CoverageEngine(names, lens, mode=mode, window=window, emit_wig=True, slice_beg=b, slice_end=e) for b, e in slices if e > b]
    try:
        for e in engs:
            e.push_host(recs, last=True)
            e.finish_runs()
        decisions, corr = shard.stitch([e.halo() for e in engs])
        bins = None
        for e, d in zip(engs, decisions):
            e.halo_put(*d)
            b = e.bins()
            bins = b if bins is None else bins + b
        for run, sign in corr:
            engs[0].bins_apply_run(run, sign, bins)
        engs[0].bins_put(bins)
        bed = b""
        for i, e in enumerate(engs):
            e.format(e.default_mask() if i == 0 else ((1 << 0) if mode == 0 else 0))
            if mode == 0:
                bed += e.output(0)
        out = {"bedgraph": bed, "depth": engs[0].output(1) if mode == 0 else b"", "wig": engs[0].output(2), "chromsize": engs[0].output(3)}
        return out
    finally:
        for e in engs:
            e.close()


@pytest.mark.parametrize("n", [2, 3, 8])
@pytest.mark.parametrize("mode", [0, 1])
def test_sharded_equals_whole(oracle, n, mode):
    names, lens, reads = bamutil.random_case(5, n_reads=6000, contigs=((b"c1", 150000), (b"c2", 40000), (b"c3", 300000)))
    recs = bamutil.records(reads, seq=False)
    for window in (1000, 16384, 777):
        exp = oracle.run(recs, names, lens, mode, window)
        got = _sharded(names, lens, recs, mode, window, n)
        for k in (("bedgraph", "depth", "wig", "chromsize") if mode == 0 else ("wig", "chromsize")):
            assert got[k] == exp[k], f"n={n} mode={mode} w={window}: {k} differs"


def test_sharded_low_coverage_long_runs(oracle):
    """constant-depth runs that span whole shards (chained merges)"""
    names, lens = [b"c1"], [200000]
    reads = [(0, 100, 0, "150000M"), (0, 120000, 0, "70000M"), (0, 199000, 0, "1000M")]
    recs = bamutil.records(reads, seq=False)
    for mode in (0, 1):
        exp = oracle.run(recs, names, lens, mode, 20000)
        got = _sharded(names, lens, recs, mode, 20000, 8)
        for k in (("bedgraph", "depth", "wig") if mode == 0 else ("wig",)):
            assert got[k] == exp[k], (mode, k)


def _sharded_device(names, lens, recs, mode, window, n, staging_bytes=0):
    """the same n engines, stitched on the DEVICE: halo records gathered with device-to-device copies (what the NCCL
    all_gather does between GPUs), ngscov_stitch_device on every shard, window bins summed in device memory"""
    import torch
    from highperformancengs_b200 import CoverageEngine, shard
    from highperformancengs_b200.engine import DeviceArray
    slices = [(b, e) for b, e in shard.plan_slices(lens, n) if e > b]
    engs = [CoverageEngine(names, lens, mode=mode, window=window, emit_wig=True, slice_beg=b, slice_end=e, staging_bytes=staging_bytes) for b, e in slices]
    try:
        for e in engs:
            e.push_host(recs, last=True)
            e.finish_runs()
            e.sync()
        world = len(engs)
        allh = torch.cat([torch.as_tensor(DeviceArray(e.halo_device(), 9), device="cuda").clone() for e in engs])
        torch.cuda.synchronize()
        for r, e in enumerate(engs):
            e.stitch_device(allh.data_ptr(), world, r)
            e.sync()
        bins = [torch.as_tensor(DeviceArray(*e.bins_device()), device="cuda") for e in engs]
        total = torch.stack(bins).sum(0)
        bins[0].copy_(total)
        torch.cuda.synchronize()
        bed = b""
        for i, e in enumerate(engs):
            e.format(e.default_mask() if i == 0 else ((1 << 0) if mode == 0 else 0))
            if mode == 0:
                bed += e.output(0)
        return {"bedgraph": bed, "depth": engs[0].output(1) if mode == 0 else b"", "wig": engs[0].output(2), "chromsize": engs[0].output(3)}
    finally:
        for e in engs:
            e.close()


@pytest.mark.parametrize("n", [2, 3, 8])
@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("staging", [0, 16384])
def test_sharded_device_stitch_equals_whole(oracle, n, mode, staging):
    """device-side stitch (ngscov_halo_device / ngscov_stitch_device / ngscov_bins_device), also with small pushes so
    that the shards' contigs are finalized while the stream is still arriving (streamed finalize of region shards)"""
    names, lens, reads = bamutil.random_case(5, n_reads=6000, contigs=((b"c1", 150000), (b"c2", 40000), (b"c3", 300000)))
    recs = bamutil.records(reads, seq=False)
    for window in (1000, 777):
        exp = oracle.run(recs, names, lens, mode, window)
        got = _sharded_device(names, lens, recs, mode, window, n, staging_bytes=staging)
        for k in (("bedgraph", "depth", "wig", "chromsize") if mode == 0 else ("wig", "chromsize")):
            assert got[k] == exp[k], f"n={n} mode={mode} w={window} staging={staging}: {k} differs"


def test_sharded_device_stitch_long_runs(oracle):
    """constant-depth runs that span whole shards (chained merges), stitched on the device"""
    names, lens = [b"c1"], [200000]
    reads = [(0, 100, 0, "150000M"), (0, 120000, 0, "70000M"), (0, 199000, 0, "1000M")]
    recs = bamutil.records(reads, seq=False)
    for mode in (0, 1):
        exp = oracle.run(recs, names, lens, mode, 20000)
        got = _sharded_device(names, lens, recs, mode, 20000, 8)
        for k in (("bedgraph", "depth", "wig") if mode == 0 else ("wig",)):
            assert got[k] == exp[k], (mode, k)


@pytest.mark.parametrize("tag,key", __import__("tests.test_oracle_golden", fromlist=["ST_PARAMS"]).ST_PARAMS)
def test_engine_on_samtools_examples(tag, key):
    """the files the reference itself ships (samtools examples/toy.sam, ex1.sam.gz) through the engine: md5 of every output
    equals that of the unmodified reference binaries"""
    from tests.test_oracle_golden import samtools_example, check_samtools_example
    from highperformancengs_b200 import CoverageEngine
    names, lens, recs = samtools_example(tag)
    tool, w = key.split("|")
    mode = 0 if tool == "depth" else 1
    with CoverageEngine(names, lens, mode=mode, window=int(w), emit_wig=True) as eng:
        eng.push_host(recs, last=True)
        eng.finish()
        got = {"wig": eng.output(2), "chromsize": eng.output(3)}
        if mode == 0:
            got.update(bedgraph=eng.output(0), depth=eng.output(1))
    check_samtools_example(got, tag, key)


@pytest.mark.parametrize("name,kind,params,runs", cases.FULL_CASES, ids=[c[0] for c in cases.FULL_CASES])
def test_full_size_configs_4_and_5(name, kind, params, runs):
    """BASELINE configs 4 (20 M spliced reads, bam2wig) and 5 (50 M reads piled on 200 kb, depth > 10^4) at FULL size:
    md5 of the engine's files against those the reference binary wrote for the same input (tests/golden)"""
    from highperformancengs_b200 import CoverageEngine
    names, lens, recs, g = cases.case_input(kind, params)
    for tool, w in runs:
        mode = 0 if tool == "depth" else 1
        with CoverageEngine(names, lens, mode=mode, window=w, emit_wig=True) as eng:
            eng.push_host(recs, last=True)
            eng.finish()
            got = {"wig": eng.output(2), "chromsize": eng.output(3)}
            if mode == 0:
                got.update(bedgraph=eng.output(0), depth=eng.output(1))
            st = eng.stats()
        assert st["records"] == g.n_reads
        check_against_gold(got, f"{name}|{tool}|{w}")
        if name == "c5_full":
            depths = np.fromiter((int(ln.rsplit(b"\t", 1)[1]) for ln in got["bedgraph"].splitlines()), dtype=np.int64)
            assert depths.max() > 10000, "config 5 is the depth > 10^4 pile-up workload"


def test_c_driver_index_free_mode(tmp_path):
    """NGSCOV_NO_INDEX=1: a sorted BAM without a .bai (the reference exits with "BAM indexing file is not available")"""
    exe = os.path.join(ROOT, "highperformancengs_b200", "host", "bam2depth")
    spec = [c for c in cases.CASES if c[0] == "random1"][0]
    names, lens, recs, g = cases.case_input(spec[1], spec[2])
    bam = str(tmp_path / "random1.bam")
    bamutil.write_bam(bam, names, lens, recs, index=False)
    r = subprocess.run([exe, "-o", "o", "-w", "150", "-W", bam], cwd=str(tmp_path), capture_output=True, text=True)
    assert r.returncode == 1 and "indexing file is not available" in r.stderr
    r = subprocess.run([exe, "-o", "o", "-w", "150", "-W", bam], cwd=str(tmp_path), capture_output=True, text=True, env=dict(os.environ, NGSCOV_NO_INDEX="1"))
    assert r.returncode == 0, r.stderr
    got = {"bedgraph": open(str(tmp_path / "random1.bam.1.bedGraph"), "rb").read(), "depth": open(str(tmp_path / "o.1.depth"), "rb").read(),
           "wig": open(str(tmp_path / "o.1.wig"), "rb").read(), "chromsize": open(str(tmp_path / "o.1.chromSize.txt"), "rb").read()}
    check_against_gold(got, "random1|depth|150")
