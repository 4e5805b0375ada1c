"""Wall-clock breakdown of the end-to-end path (host buffers in, text out) -- diagnostics, not a benchmark."""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from highperformancengs_b200 import CoverageEngine
from highperformancengs_b200.gen import Generator
ap = argparse.ArgumentParser(); ap.add_argument("--scale", type=float, default=1.0); ap.add_argument("--steps", type=int, default=3)
a = ap.parse_args()
g = Generator(2, scale=a.scale, seq=False)
n = g.records_size()
host = torch.empty(n, dtype=torch.uint8, pin_memory=True)
g.records_into(host.data_ptr(), n)
eng = CoverageEngine(g.names, g.lens, mode=0, window=20000)
for i in range(a.steps):
    torch.cuda.synchronize(); t0 = time.time()
    eng.reset(); t1 = time.time()
    eng.push_host_ptr(host.data_ptr(), n, last=True); t2 = time.time()
    eng.finish(); t3 = time.time()
    b = eng.output_view(0).size; t4 = time.time()
    d = eng.output_view(1).size; t5 = time.time()
    st = eng.stats()
    print(f"step {i}: reset {t1-t0:.3f} push {t2-t1:.3f} finish {t3-t2:.3f} out_bed {t4-t3:.3f} out_depth {t5-t4:.3f} total {t5-t0:.3f} s | "
          f"h2d {st['ms_h2d']:.1f} ms ({n/1e6/max(st['ms_h2d'],1e-9):.1f} GB/s) ingest {st['ms_ingest']:.1f} scan {st['ms_scan']:.1f} format {st['ms_format']:.1f} d2h_wait {st['ms_d2h']:.1f} bed {b/1e9:.2f} GB")
