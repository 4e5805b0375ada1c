"""Pins the FULL-SIZE config-2 job (30x WGS shape: 618 271 620 reads on 3.09 Gbp, window 20000) against the unmodified
reference: generates the benchmark's BAM, runs oracle/_ref/bam2depth on it (one process, ~25 min, ~10 GB of RAM),
and records size, md5 and the position-weighted 64-bit hash of its bedGraph and .depth files -- the same hash bench.py
computes on the device over the engine's output every run (`parity.hash`), so the benchmark line itself says whether the
measured job reproduced the reference byte for byte.

    python tests/golden/make_config2_full_golden.py [workdir]      (needs oracle/_ref and ~35 GB of disk)
"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from file_hash import file_hash  # noqa: E402


def main():
    from highperformancengs_b200.gen import Generator
    work = sys.argv[1] if len(sys.argv) > 1 else "/tmp/ngscov_full"
    os.makedirs(work, exist_ok=True)
    bam = os.path.join(work, "c2.bam")
    g = Generator(2, scale=1.0)
    if not os.path.exists(bam):
        g.write_bam(bam, level=1)
    bed, dep = os.path.join(work, "c2.bam.1.bedGraph"), os.path.join(work, "ref.1.depth")
    t0 = time.time()
    if not (os.path.exists(bed) and os.path.exists(dep) and os.path.exists(os.path.join(work, "done"))):
        subprocess.check_call([os.path.join(ROOT, "oracle", "_ref", "bam2depth"), "-o", os.path.join(work, "ref"), "-w", "20000", bam],
                              cwd=work, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        open(os.path.join(work, "done"), "w").write("%.1f\n" % (time.time() - t0))
    out = {"config": 2, "scale": 1.0, "reads": int(g.n_reads), "window": 20000,
           "reference_seconds": float(open(os.path.join(work, "done")).read()),
           "made_by": "tests/golden/make_config2_full_golden.py: oracle/_ref/bam2depth -o ref -w 20000 c2.bam (unmodified reference, one process)",
           "bedgraph": {k: v for k, v in file_hash(bed).items() if k != "file"},
           "depth": {k: v for k, v in file_hash(dep).items() if k != "file"}}
    with open(os.path.join(ROOT, "tests", "golden", "config2_full_reference.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
