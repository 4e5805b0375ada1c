"""Device BGZF inflate throughput (never a bench.py number): config-2 BAM at --scale, compressed bytes in pinned host
memory -> ngscov_push_bgzf -> finish.  Prints the inflate kernel time, its GB/s of inflated output and the wall time
next to the same job fed with host-inflated records (ngscov_push_host)."""
import argparse
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from highperformancengs_b200 import CoverageEngine, bamio
from highperformancengs_b200.gen import Generator

ap = argparse.ArgumentParser()
ap.add_argument("--scale", type=float, default=0.1)
ap.add_argument("--level", type=int, default=1)
ap.add_argument("--steps", type=int, default=3)
a = ap.parse_args()
g = Generator(2, scale=a.scale, seq=False)
with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as d:
    bam = os.path.join(d, "x.bam")
    t0 = time.time(); g.write_bam(bam, level=a.level); t_gen = time.time() - t0
    raw = np.fromfile(bam, dtype=np.uint8)
comp = torch.empty(raw.size, dtype=torch.uint8, pin_memory=True)
comp.numpy()[:] = raw
hdr = bamio.bgzf_inflate(bytes(raw[:1 << 20].tobytes()[:0]) or b"") if False else None
# header length: inflate leading blocks until the header parses
pos, inf = 0, b""
for off, bs in bamio.bgzf_blocks(memoryview(raw[:4 << 20])):
    import zlib
    inf += zlib.decompress(bytes(raw[off + 18:off + bs - 8]), -15)
    try:
        _, _, hdr_len = bamio.parse_header(inf)
        break
    except Exception:
        continue
recs = g.records()
host = torch.empty(recs.size, dtype=torch.uint8, pin_memory=True)
host.numpy()[:] = recs
eng = CoverageEngine(g.names, g.lens, mode=0, window=20000)
res = {}
for name in ("bgzf", "host"):
    best = None
    for _ in range(a.steps):
        eng.reset(); torch.cuda.synchronize(); t0 = time.time()
        if name == "bgzf":
            eng.push_bgzf_ptr(comp.data_ptr(), comp.numel(), skip=hdr_len, last=True)
        else:
            eng.push_host_ptr(host.data_ptr(), host.numel(), last=True)
        eng.finish(); n = eng.output_view(0).size; dt = time.time() - t0
        st = eng.stats()
        if best is None or dt < best[0]:
            best = (dt, st, n)
    res[name] = best
st = res["bgzf"][1]
print({"reads": g.n_reads, "compressed_MB": raw.size / 1e6, "inflated_MB": recs.size / 1e6, "ratio": recs.size / raw.size, "gen_s": round(t_gen, 2),
       "inflate_ms": st["ms_inflate"], "inflate_GBps_out": recs.size / st["ms_inflate"] / 1e6, "inflate_GBps_in": raw.size / st["ms_inflate"] / 1e6,
       "wall_bgzf_s": res["bgzf"][0], "wall_host_s": res["host"][0], "same_output": res["bgzf"][2] == res["host"][2],
       "stage_ms_bgzf": {k: st[k] for k in ("ms_h2d", "ms_inflate", "ms_ingest", "ms_scan", "ms_format")}})
eng.close()
