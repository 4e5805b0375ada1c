"""Regenerates tests/golden/slide_outputs.json by running the UNMODIFIED reference bam_sliding_count (oracle/_ref, built
by oracle/Makefile from /root/reference with the libgd stubs of oracle/gdstub/) on every case of tests/slide_cases.py.

    python tests/golden/make_slide_golden.py          (needs oracle/_ref: the build container)
"""
import hashlib
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests import bamutil, slide_cases  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "bam_sliding_count")


def run_reference(window, bam, workdir, region=None):
    cmd = [REF, "-o", os.path.join(workdir, "out"), "-w", str(window)] + (["-r", region] if region else []) + [bam]
    r = subprocess.run(cmd, cwd=workdir, capture_output=True, text=True, check=True)
    n_count = int([ln for ln in r.stderr.splitlines() if ln.strip().isdigit()][0])
    with open(os.path.join(workdir, "out.txt"), "rb") as f:
        return (f.read(), n_count, r.stdout) if region else (f.read(), n_count)


def main():
    gold = {}
    for name, kind, params, windows in slide_cases.CASES:
        names, lens, recs, g = slide_cases.case_input(kind, params)
        with tempfile.TemporaryDirectory(prefix="ngscov_slide_") as d:
            bam = os.path.join(d, name + ".bam")
            bamutil.write_bam(bam, names, lens, recs, index=False)        # exactly the records the tests feed, in file order
            for w in windows:
                txt, n_count = run_reference(w, bam, d)
                e = {"md5": hashlib.md5(txt).hexdigest(), "size": len(txt), "n_count": n_count}
                if len(txt) <= 4096:
                    e["text"] = txt.decode()
                gold[f"{name}|{w}"] = e
                print(name, w, e["size"], n_count)
    with open(os.path.join(ROOT, "tests", "golden", "slide_outputs.json"), "w") as f:
        json.dump(gold, f, indent=1, sort_keys=True)
    # the -r mode: the same inputs, indexed (the reference insists on the .bai), one region each
    rgold = {}
    by_name = {c[0]: c for c in slide_cases.CASES}
    for name, w, region, parsed in slide_cases.REGION_CASES:
        _, kind, params, _ = by_name[name]
        names, lens, recs, g = slide_cases.case_input(kind, params)
        with tempfile.TemporaryDirectory(prefix="ngscov_slide_") as d:
            bam = os.path.join(d, name + ".bam")
            bamutil.write_bam(bam, names, lens, recs, index=True)
            txt, n_count, out = run_reference(w, bam, d, region)
            nm, b, e = out.split("\n")[0].split("\t")
            assert (names[parsed[0]].decode(), int(b), int(e)) == (nm, parsed[1], parsed[2]), (out, parsed)
            rgold[f"{name}|{w}|{region}"] = {"md5": hashlib.md5(txt).hexdigest(), "size": len(txt), "n_count": n_count, "stdout": out.split("\n")[0] + "\n"}
            print(name, w, region, len(txt), n_count)
    with open(os.path.join(ROOT, "tests", "golden", "slide_region_outputs.json"), "w") as f:
        json.dump(rgold, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
