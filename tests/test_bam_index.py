"""host/bam_index: the BAI writer against samtools-0.1.19's own indexer (the reference's
vendored libbam, through oracle/_ref/sam2bam --index-only, where it was built) and against the index the generator writes."""
import os
import shutil
import subprocess

import pytest

from tests import bamutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "highperformancengs_b200", "host", "bam_index")
SAM2BAM = os.path.join(ROOT, "oracle", "_ref", "sam2bam")


@pytest.fixture(scope="module")
def exe():
    if not os.path.exists(EXE):
        subprocess.check_call(["gcc", os.path.join(ROOT, "highperformancengs_b200", "host", "ngscov_index_cli.c"), "-o", EXE, "-O2", "-lz", "-lpthread"])
    return EXE


def same_index(a, b):
    from highperformancengs_b200 import bamio
    (ra, na), (rb, nb) = bamio.read_bai(a), bamio.read_bai(b)
    assert na == nb and len(ra) == len(rb)
    for i, (x, y) in enumerate(zip(ra, rb)):
        assert x["bins"] == y["bins"], f"contig {i}: bins differ"
        assert x["linear"] == y["linear"], f"contig {i}: linear index differs"


@pytest.mark.parametrize("seed,n_reads", [(1, 3000), (2, 3000), (8, 40000)])
def test_against_samtools_indexer(exe, tmp_path, seed, n_reads):
    if not os.path.exists(SAM2BAM):
        pytest.skip("oracle/_ref/sam2bam not built (needs /root/reference)")
    names, lens, reads = bamutil.random_case(seed, n_reads=n_reads, contigs=((b"c1", 300000), (b"cEmpty", 5000), (b"c3", 70000)))
    reads = [r for r in reads if r[0] != 1]                      # a contig without reads
    recs = bamutil.records(reads)
    bam = str(tmp_path / "t.bam")
    bamutil.write_bam(bam, names, lens, recs, index=False)
    subprocess.check_call([exe, bam, str(tmp_path / "ours.bai")])
    subprocess.check_call([SAM2BAM, "--index-only", bam])
    same_index(str(tmp_path / "ours.bai"), bam + ".bai")


@pytest.mark.parametrize("config,scale", [(1, 0.05), (2, 0.0005), (4, 0.01)])
def test_against_generator_index(exe, tmp_path, config, scale):
    """the generator (tools/bamgen.c) restates the same indexer for its own files: byte-identical .bai"""
    from highperformancengs_b200.gen import Generator
    g = Generator(config, scale=scale)
    bam = str(tmp_path / "g.bam")
    g.write_bam(bam)
    subprocess.check_call([exe, bam, str(tmp_path / "ours.bai")])
    assert open(str(tmp_path / "ours.bai"), "rb").read() == open(bam + ".bai", "rb").read()


def test_reference_accepts_the_index(exe, tmp_path, ref_bin):
    """the unmodified reference bam2depth run with OUR index gives the outputs it gives with samtools' index"""
    if not ref_bin:
        pytest.skip("oracle/_ref not built")
    names, lens, reads = bamutil.random_case(3)
    recs = bamutil.records(reads)
    outs = []
    for k in range(2):
        d = tmp_path / str(k)
        d.mkdir()
        bam = str(d / "t.bam")
        bamutil.write_bam(bam, names, lens, recs, index=False)
        if k == 0:
            subprocess.check_call([exe, bam])
        else:
            subprocess.check_call([SAM2BAM, "--index-only", bam])
        subprocess.check_call([os.path.join(ROOT, "oracle", "_ref", "bam2depth"), "-o", "o", "-w", "500", "-W", bam], cwd=str(d), stderr=subprocess.DEVNULL)
        outs.append({f: open(str(d / f), "rb").read() for f in sorted(os.listdir(str(d))) if not f.startswith("t.bam.bai") and f != "t.bam"})
    assert outs[0] == outs[1] and len(outs[0]) == 4


def test_unsorted_input_is_refused(exe, tmp_path):
    names, lens, reads = bamutil.random_case(4, n_reads=500)
    reads = [r for r in reads if r[0] >= 0]
    reads = reads[len(reads) // 2:] + reads[:len(reads) // 2]
    bam = str(tmp_path / "u.bam")
    bamutil.write_bam(bam, names, lens, bamutil.records(reads), index=False)
    r = subprocess.run([exe, bam], capture_output=True, text=True)
    assert r.returncode == 1 and "not sorted" in r.stderr
