"""One resident-input pass of the hot path, for use under ncu (never a benchmark number)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from highperformancengs_b200 import CoverageEngine
from highperformancengs_b200.gen import Generator

ap = argparse.ArgumentParser()
ap.add_argument("--config", type=int, default=2)
ap.add_argument("--scale", type=float, default=0.1)
ap.add_argument("--steps", type=int, default=1)
ap.add_argument("--mode", type=int, default=0)
ap.add_argument("--window", type=int, default=20000)
a = ap.parse_args()
g = Generator(a.config, scale=a.scale, seq=False if a.config != 1 else True)
recs = torch.from_numpy(g.records()).cuda()
eng = CoverageEngine(g.names, g.lens, mode=a.mode, window=a.window)
for _ in range(a.steps):
    eng.reset()
    eng.push_device(recs.data_ptr(), recs.numel(), last=True)
    eng.finish()
st = eng.stats()
print({k: st[k] for k in ("records", "runs", "ms_ingest", "ms_scan", "ms_format", "launches", "spec_repaired")})
eng.close()
