"""Derives tests/golden/samtools_examples.json from the example alignments the reference itself ships inside its vendored
samtools tarball (/root/reference/samtools-0.1.19.tar.bz2: examples/toy.sam, examples/ex1.sam.gz) -- the only
reference-held inputs for this path.  Only the fields the coverage path reads are kept
(contig, 0-based POS, FLAG, CIGAR, in file order) plus the contig table; the expected md5s are those of the UNMODIFIED
reference binaries (oracle/_ref) run on the real files here, and are checked against the values SURVEY.md recorded.

    python tests/golden/make_samtools_fixtures.py      (needs /root/reference and oracle/_ref: the build container)
"""
import gzip
import hashlib
import json
import os
import subprocess
import sys
import tarfile
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = os.path.join(ROOT, "oracle", "_ref")
TARBALL = "/root/reference/samtools-0.1.19.tar.bz2"
SURVEY_MD5 = {
    ("toy", "depth", 10, "bedgraph"): "4f4f60ece97bd03e9c84e2960c26def2",
    ("ex1", "depth", 100, "bedgraph"): "8a06af30086d31a290d43f2ac3ad2343",
    ("ex1", "depth", 100, "depth"): "0ae8d46b075ea0bac0fd01e1a61b3226",
    ("ex1", "depth", 100, "wig"): "84fce416017db6a96779e0a8f30b2007",
    ("ex1", "wig", 100, "wig"): "1385dda8cc49d86f0e1f104f509ee33d",
}


def parse_sam(text, contigs):
    idx = {n: i for i, (n, _) in enumerate(contigs)}
    reads = []
    for ln in text.splitlines():
        if not ln or ln[0] == "@":
            continue
        f = ln.split("\t")
        reads.append([idx.get(f[2], -1), int(f[3]) - 1, int(f[1]), f[5]])
    return reads


def main():
    from tests.golden.make_golden import run_reference
    out = {}
    with tempfile.TemporaryDirectory(prefix="ngscov_st_") as d:
        with tarfile.open(TARBALL) as t:
            for m in ("toy.sam", "ex1.sam.gz", "ex1.fa"):
                with open(os.path.join(d, m), "wb") as f:
                    f.write(t.extractfile("samtools-0.1.19/examples/" + m).read())
        toy = open(os.path.join(d, "toy.sam")).read()
        ex1 = gzip.open(os.path.join(d, "ex1.sam.gz"), "rt").read()
        toy_ctg = [(ln.split("\t")[1][3:], int(ln.split("\t")[2][3:])) for ln in toy.splitlines() if ln.startswith("@SQ")]
        # ex1.sam.gz has no header: contig lengths from ex1.fa, as `samtools faidx` + `import -t` would take them
        ex1_ctg, name = [], None
        for ln in open(os.path.join(d, "ex1.fa")):
            if ln[0] == ">":
                name = ln[1:].strip(); ex1_ctg.append([name, 0])
            else:
                ex1_ctg[-1][1] += len(ln.strip())
        ex1_ctg = [tuple(x) for x in ex1_ctg]
        for tag, text, ctg, runs in (("toy", toy, toy_ctg, [("depth", 10), ("wig", 10), ("depth", 7)]),
                                     ("ex1", ex1, ex1_ctg, [("depth", 100), ("wig", 100), ("depth", 20000)])):
            sam = os.path.join(d, tag + ".sam")
            with open(sam, "w") as f:
                for n, L in ctg:
                    f.write(f"@SQ\tSN:{n}\tLN:{L}\n")
                f.write("".join(ln + "\n" for ln in text.splitlines() if ln and ln[0] != "@"))
            bam = os.path.join(d, tag + ".bam")
            subprocess.check_call([os.path.join(REF, "sam2bam"), sam, bam])       # the reference's own libbam: import + sort + index
            exp = {}
            for tool, w in runs:
                res = run_reference(tool, w, bam, d)
                exp[f"{tool}|{w}"] = {k: {"md5": hashlib.md5(v).hexdigest(), "size": len(v)} for k, v in res.items()}
                for k, v in res.items():
                    want = SURVEY_MD5.get((tag, tool, w, k))
                    assert want is None or want == hashlib.md5(v).hexdigest(), (tag, tool, w, k, "differs from the md5 SURVEY.md recorded")
            reads = parse_sam(text, ctg)
            reads.sort(key=lambda r: (r[0] if r[0] >= 0 else 1 << 30, r[1]))      # coordinate order, as the sorted BAM holds them (stable)
            out[tag] = {"contigs": [[n, L] for n, L in ctg], "reads": reads, "expected": exp}
            print(tag, len(reads), "reads", {k: {f: e["size"] for f, e in v.items()} for k, v in exp.items()})
    with open(os.path.join(ROOT, "tests", "golden", "samtools_examples.json"), "w") as f:
        json.dump(out, f, separators=(",", ":"), sort_keys=True)


if __name__ == "__main__":
    main()
