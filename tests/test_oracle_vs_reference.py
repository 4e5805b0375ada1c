"""CPU, only where the reference binaries exist (oracle/_ref, i.e. the build container): run the
unmodified reference tools and compare with the oracle byte for byte; also cross-check our BAM/BAI
writer against samtools' own indexer."""
import os
import subprocess

import pytest

from tests import bamutil, cases
from highperformancengs_b200 import bamio


def _ref_outputs(ref_bin, tool, window, bam, d):
    pfx = os.path.join(d, "o")
    if tool == "depth":
        subprocess.check_call([ref_bin["bam2depth"], "-o", pfx, "-w", str(window), "-W", bam], cwd=d, stderr=subprocess.DEVNULL)
        files = {"bedgraph": os.path.join(d, os.path.basename(bam) + ".1.bedGraph"), "depth": pfx + ".1.depth", "wig": pfx + ".1.wig",
                 "chromsize": pfx + ".1.chromSize.txt"}
    else:
        subprocess.check_call([ref_bin["bam2wig"], "-o", pfx, "-w", str(window), bam], cwd=d, stderr=subprocess.DEVNULL)
        files = {"wig": pfx + ".1.wig", "chromsize": pfx + ".1.chromSize.txt"}
    return {k: open(p, "rb").read() for k, p in files.items()}


@pytest.mark.parametrize("seed", [11, 12])
@pytest.mark.parametrize("tool,window", [("depth", 1), ("depth", 333), ("wig", 333), ("wig", 20000), ("wig", 2)])
def test_reference_binary_vs_oracle_random(ref_bin, oracle, tmp_path, seed, tool, window):
    if ref_bin is None:
        pytest.skip("reference binaries not built here (oracle/_ref)")
    names, lens, reads = bamutil.random_case(seed, n_reads=1500)
    recs = bamutil.records(reads)
    bam = str(tmp_path / "r.bam")
    bamutil.write_bam(bam, names, lens, recs)
    ref = _ref_outputs(ref_bin, tool, window, bam, str(tmp_path))
    got = oracle.run(recs, names, lens, 0 if tool == "depth" else 1, window)
    for k, v in ref.items():
        assert got[k] == v, f"{k} differs from the reference binary"


@pytest.mark.parametrize("config,scale", [(2, 0.0003), (4, 0.005)])
def test_generator_bam_and_bai_are_accepted_by_reference(ref_bin, oracle, tmp_path, config, scale):
    if ref_bin is None:
        pytest.skip("reference binaries not built here (oracle/_ref)")
    from highperformancengs_b200.gen import Generator
    g = Generator(config, scale=scale)
    bam = str(tmp_path / "g.bam")
    g.write_bam(bam)
    # the file round-trips through our own reader
    names, lens, recs = bamio.load_bam(bam)
    assert names == g.names and lens == g.lens and bytes(recs) == bytes(g.records())
    # our .bai carries the same linear index and the same chunks as samtools' indexer builds
    mine, nc_mine = bamio.read_bai(bam + ".bai")
    alt = str(tmp_path / "alt.bam")
    os.link(bam, alt)
    subprocess.check_call([ref_bin["sam2bam"], "--index-only", alt], stderr=subprocess.DEVNULL)
    theirs, nc_theirs = bamio.read_bai(alt + ".bai")
    assert nc_mine == nc_theirs
    for a, b in zip(mine, theirs):
        assert a["linear"] == b["linear"]
        assert {k: sorted(v) for k, v in a["bins"].items()} == {k: sorted(v) for k, v in b["bins"].items()}
    tool = "wig" if config == 4 else "depth"
    ref = _ref_outputs(ref_bin, tool, 1000, bam, str(tmp_path))
    got = oracle.run(recs, names, lens, 0 if tool == "depth" else 1, 1000)
    for k, v in ref.items():
        assert got[k] == v
