"""Regenerates tests/golden/reference_outputs.json by running the UNMODIFIED reference binaries
(oracle/_ref/bam2depth, bam2wig, built by oracle/Makefile from /root/reference) on every case of
tests/cases.py.  Only runs where oracle/_ref exists (the build container); the JSON travels.

    python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests import bamutil, cases  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref")


def run_reference(tool, window, bam, workdir):
    """-> dict file -> bytes, as written by the reference tool"""
    pfx = os.path.join(workdir, "o")
    if tool == "depth":
        subprocess.check_call([os.path.join(REF, "bam2depth"), "-o", pfx, "-w", str(window), "-W", bam], cwd=workdir, stderr=subprocess.DEVNULL)
        names = {"bedgraph": os.path.join(workdir, os.path.basename(bam) + ".1.bedGraph"), "depth": pfx + ".1.depth",
                 "wig": pfx + ".1.wig", "chromsize": pfx + ".1.chromSize.txt"}
    else:
        subprocess.check_call([os.path.join(REF, "bam2wig"), "-o", pfx, "-w", str(window), bam], cwd=workdir, stderr=subprocess.DEVNULL)
        names = {"wig": pfx + ".1.wig", "chromsize": pfx + ".1.chromSize.txt"}
    out = {}
    for k, p in names.items():
        with open(p, "rb") as f:
            out[k] = f.read()
        os.remove(p)
    return out


def main():
    only = set(sys.argv[1:])          # optional case names: regenerate just those and merge into the existing file
    path = os.path.join(ROOT, "tests", "golden", "reference_outputs.json")
    gold = json.load(open(path)) if only and os.path.exists(path) else {}
    for name, kind, params, runs in cases.CASES + cases.FULL_CASES:
        if only and name not in only:
            continue
        names, lens, recs, g = cases.case_input(kind, params)
        with tempfile.TemporaryDirectory(prefix="ngscov_gold_") as d:
            bam = os.path.join(d, name + ".bam")
            if g is not None:
                g.write_bam(bam)          # our BGZF + BAI writer: the reference consumes both
            else:
                bamutil.write_bam(bam, names, lens, recs)
            for tool, window in runs:
                out = run_reference(tool, window, bam, d)
                entry = {}
                for k, v in out.items():
                    entry[k] = {"md5": hashlib.md5(v).hexdigest(), "size": len(v)}
                    if len(v) <= 2048:
                        entry[k]["text"] = v.decode()
                gold[f"{name}|{tool}|{window}"] = entry
                print(name, tool, window, {k: e["size"] for k, e in entry.items()})
    with open(path, "w") as f:
        json.dump(gold, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
