"""Times the device BGZF inflate (k_bgzf_inflate) alone: a generated BAM's blocks in pinned host memory go through
ngscov_push_bgzf; the engine's own CUDA events around every inflate launch give the kernel time.

    python tools/k0_bench.py [--config 2] [--scale 0.05] [--seq none|full] [--level 1] [--reps 3]

Prints one JSON line: inflated GB/s, compressed GB/s, and the md5 of the bedGraph (to compare library builds)."""
import argparse
import hashlib
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=2)
    ap.add_argument("--scale", type=float, default=0.05)
    ap.add_argument("--seq", default=None)
    ap.add_argument("--level", type=int, default=1)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--bam", default=None)
    ap.add_argument("--staging-mb", type=int, default=0, help="chunk size of the engine (0: its default, 256 MiB)")
    a = ap.parse_args()
    import numpy as np
    import torch
    from highperformancengs_b200 import bamio
    from highperformancengs_b200.engine import CoverageEngine
    from highperformancengs_b200.gen import Generator
    kw = {"seq": a.seq == "full"} if a.seq else {}
    g = Generator(a.config, scale=a.scale, **kw)
    path = a.bam or os.path.join("/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir(), "k0_bench_%d_%s_%s.bam" % (a.config, a.scale, a.seq))
    if not os.path.exists(path):
        g.write_bam(path, level=a.level)
    raw = np.fromfile(path, dtype=np.uint8)
    comp = torch.empty(raw.size, dtype=torch.uint8, pin_memory=True)
    comp.numpy()[:] = raw
    import zlib
    inf, hdr_len = b"", None
    for off, bs in bamio.bgzf_blocks(memoryview(raw[:min(raw.size, 8 << 20)])):
        inf += zlib.decompress(bytes(raw[off + 18:off + bs - 8]), -15)
        try:
            hdr_len = bamio.parse_header(inf)[2]
            break
        except Exception:
            continue
    eng = CoverageEngine(g.names, [int(x) for x in g.lens], mode=0, window=20000, emit_wig=False, device=0,
                         **({"staging_bytes": a.staging_mb << 20} if a.staging_mb else {}))
    best, md5 = None, None
    for r in range(a.reps + 1):
        eng.reset()
        t0 = time.time()
        eng.push_bgzf_ptr(comp.data_ptr(), raw.size, skip=hdr_len, last=True)
        eng.finish_runs()
        eng.sync()
        wall = time.time() - t0
        st = eng.stats()
        if r == 0:
            eng.format(eng.default_mask())
            md5 = hashlib.md5(eng.output_view(0)).hexdigest()
            continue
        ms = st["ms_inflate"]
        if best is None or ms < best[0]:
            best = (ms, st["stream_bytes"], wall)
    ms, sb, wall = best
    print(json.dumps({"lib": os.environ.get("NGSCOV_LIB", "default"), "inf_s": os.environ.get("NGSCOV_INF_S"), "staging_mb": a.staging_mb, "config": a.config, "scale": a.scale, "seq": a.seq, "level": a.level,
                      "inflate_ms": ms, "inflated_GBps": sb / ms / 1e6, "compressed_GBps": raw.size / ms / 1e6,
                      "ratio": sb / raw.size, "wall_ms": wall * 1e3, "bedgraph_md5": md5}))


if __name__ == "__main__":
    main()
