"""Small engine runs for `compute-sanitizer` (tests/test_gpu_sanitizer.py starts this file under memcheck / racecheck):
every kernel of the cover pipeline incl. far deltas, CHAIN entries, region shards with the device-side stitch, streamed
pushes with speculation + repair, the device BGZF inflate, and the classic kernels.  TEST INFRASTRUCTURE."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import bamutil, oracle_py  # noqa: E402


def run(names, lens, recs, mode, window, **kw):
    from highperformancengs_b200 import CoverageEngine
    with CoverageEngine(names, lens, mode=mode, window=window, emit_wig=True, **kw) as eng:
        eng.push_host(recs, last=True)
        eng.finish()
        out = {"wig": eng.output(2)}
        if mode == 0:
            out.update(bedgraph=eng.output(0), depth=eng.output(1))
        return out


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "cover"
    if what == "classic":
        os.environ["NGSCOV_PIPELINE"] = "classic"
    names, lens, reads = bamutil.random_case(31, n_reads=1500, contigs=((b"c1", 30000), (b"c2", 9000), (b"c3", 50000)))
    reads = reads[:-20] + [(2, 100, 0, "700M"), (2, 300, 0, "40M3000N40M"), (2, 900, 16, "5S20M2I20M3D20M900N30M")] + reads[-20:]
    reads.sort(key=lambda r: (r[0] if r[0] >= 0 else 1 << 30, r[1]))
    recs = bamutil.records(reads, seq=True)
    for mode in (0, 1):
        exp = oracle_py.run(recs, names, lens, mode, 500)
        for kw in ({}, {"staging_bytes": 8192}):
            got = run(names, lens, recs, mode, 500, **kw)
            for k in got:
                assert got[k] == exp[k], (what, mode, kw, k)
    if what == "cover":
        # device BGZF inflate in front of the engine
        import tempfile
        from highperformancengs_b200 import CoverageEngine, bamio
        with tempfile.TemporaryDirectory() as d:
            bam = os.path.join(d, "s.bam")
            bamutil.write_bam(bam, names, lens, recs)
            data = open(bam, "rb").read()
        hdr_len = bamio.parse_header(bamio.bgzf_inflate(data))[2]
        exp = oracle_py.run(recs, names, lens, 0, 500)
        for decoders in ("1", "2"):                    # one decoder per warp, and the opt-in two-per-warp state machine
            os.environ["NGSCOV_INF_S"] = decoders
            with CoverageEngine(names, lens, mode=0, window=500, emit_wig=True) as eng:
                eng.push_bgzf(data, skip=hdr_len, last=True)
                eng.finish()
                assert eng.output(0) == exp["bedgraph"] and eng.output(1) == exp["depth"]
        del os.environ["NGSCOV_INF_S"]
    print("sanitize_case ok:", what)


if __name__ == "__main__":
    main()
