"""compute-sanitizer over the kernels: memcheck on both pipelines, racecheck on the cover pipeline
(shared-memory accumulators, the aliased boundary list, word ORs of the line writer, look-back words)."""
import os
import shutil
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CS = shutil.which("compute-sanitizer") or "/usr/local/cuda/bin/compute-sanitizer"


@pytest.mark.parametrize("tool,what", [("memcheck", "cover"), ("memcheck", "classic"), ("racecheck", "cover"), ("synccheck", "cover")])
def test_compute_sanitizer_clean(tool, what, oracle):
    if not os.path.exists(CS):
        pytest.skip("compute-sanitizer is not installed")
    cmd = [CS, "--tool", tool, "--error-exitcode", "86", "--print-limit", "20", sys.executable, os.path.join(ROOT, "tests", "sanitize_case.py"), what]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=1500, cwd=ROOT)
    tail = (r.stdout + r.stderr)[-3000:]
    assert r.returncode == 0, f"compute-sanitizer --tool {tool} ({what}) reported errors / hazards or the run failed:\n{tail}"
    assert "sanitize_case ok" in r.stdout, tail
