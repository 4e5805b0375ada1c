#!/usr/bin/env python
"""bench.py -- throughput of the coverage hot path (BAM record stream -> bedGraph + .depth) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--scale S]

Workload (BASELINE.json configs[1] / configs[2]): synthetic 30x WGS, ~617.65 M x 150 bp reads over an
hg38-shaped 24-contig 3.09 Gbp genome, records without SEQ/QUAL (l_seq = 0, "--seq none": the full-SEQ
stream is 170 GB and cannot be HBM-resident next to its outputs), bam2depth -w 20000.  A step is one
pass of the whole job.  With N GPUs the genome is region-sharded (one process per GPU, halo stitch over
NCCL); total work is fixed.

One JSON line is printed by rank 0:
  value     reads/s with the record stream already resident in HBM (CUDA events on the engine stream)
  e2e       N=1: the same job from the BAM file's compressed BGZF blocks in pinned host memory (ngscov_push_bgzf: H2D of the
            compressed bytes, device inflate, engine, text copied back to the host); e2e.inflated_input = the job fed with
            the already-inflated record stream (ngscov_push_host).  N>1: the push_host form only.
  roofline  the slowest stage's algorithmic bytes / its device time vs the measured HBM copy peak
  cpu_baseline  the unmodified reference bam2depth (oracle/_ref) timed on a bounded sample, 1 core
`--impl reference` times the reference tool itself on all host cores (one process per core, each on
its own sample BAM).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "aligned reads/s BAM->bedGraph depth"
WORKLOAD = ("config2/3: synthetic 30x WGS, 150 bp reads over hg38-shaped 24 contigs (3.09 Gbp), "
            "records without SEQ/QUAL (--seq none), bam2depth -w 20000 -> bedGraph + .depth")


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.stop = index, [], False
        self.t = threading.Thread(target=self.run, daemon=True)

    def run(self):
        while not self.stop:
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.t.join(2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": float(self.rows[0][2]) if self.rows[0][2].replace(".", "").isdigit() else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


# --------------------------------------------------------------------------------------- reference arm
def make_sample_bam(path, reads, seed):
    from highperformancengs_b200 import gen
    gen.build()
    scale = reads / 617654000.0
    subprocess.check_call([os.path.join(ROOT, "tools", "bamgen"), "--config", "2", "--scale", repr(scale), "--seed", str(seed),
                           "--seq", "none", "-o", path], stderr=subprocess.DEVNULL)
    from highperformancengs_b200.gen import Generator
    return Generator(2, scale=scale, seed=seed, seq=False).n_reads


def ref_tool():
    p = os.path.join(ROOT, "oracle", "_ref", "bam2depth")
    if os.path.exists(p):
        return [p], "reference"
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "oracle"])
    return [os.path.join(ROOT, "oracle", "_build", "ngscov_oracle"), "depth"], "port"


def run_reference_once(cmd, bams, workdir):
    """one process per BAM, concurrently; returns wall seconds"""
    t0 = time.time()
    procs = [subprocess.Popen(cmd + ["-o", os.path.join(workdir, f"o{i}"), "-w", "20000", b], cwd=workdir,
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL) for i, b in enumerate(bams)]
    for p in procs:
        if p.wait() != 0:
            raise RuntimeError("reference tool failed")
    return time.time() - t0


def cpu_baseline(sample_reads=2000000):
    cmd, kind = ref_tool()
    with tempfile.TemporaryDirectory(prefix="ngscov_cpu_") as d:
        bam = os.path.join(d, "sample.bam")
        n = make_sample_bam(bam, sample_reads, 11)
        dt = run_reference_once(cmd, [bam], d)
    return {"value": n / dt, "unit": "reads/s", "cores": 1, "kind": kind,
            "sample": f"{n} reads of the same 30x WGS generator (scaled genome), one bam2depth process, {dt:.2f} s wall"}


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cmd, kind = ref_tool()
    cores = os.cpu_count() or 1
    per = args.ref_sample_reads
    with tempfile.TemporaryDirectory(prefix="ngscov_ref_") as d:
        bams, n_tot = [], 0
        for i in range(cores):
            b = os.path.join(d, f"s{i}.bam")
            n_tot += make_sample_bam(b, per, 100 + i)
            bams.append(b)
        for _ in range(args.warmup):
            run_reference_once(cmd, bams, d)
        t = [run_reference_once(cmd, bams, d) for _ in range(args.steps)]
    dt = sum(t) / len(t)
    v = n_tot / dt
    line = {"metric": METRIC, "value": v, "unit": "reads/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int32",
            "data": "synthetic", "impl": "reference",
            "config": {"workload": WORKLOAD, "sample": f"{cores} concurrent single-threaded processes x {per} reads each per step"},
            "cpu_baseline": {"value": v, "unit": "reads/s", "cores": cores, "kind": kind,
                             "sample": f"{n_tot} reads per step: {cores} processes (the tool is single-threaded), each on its own sample BAM"},
            "e2e": {"value": v, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------- our arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--scale", type=float, default=float(os.environ.get("NGSCOV_BENCH_SCALE", "1.0")))
    ap.add_argument("--window", type=int, default=20000)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-bgzf", action="store_true", help="skip the end-to-end leg that starts from the compressed BAM bytes")
    ap.add_argument("--check", action="store_true", help="small scales only: gather every rank's text on rank 0 and compare with the CPU oracle")
    ap.add_argument("--ref-sample-reads", type=int, default=1000000)
    ap.add_argument("--cpu-sample-reads", type=int, default=2000000)
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    from highperformancengs_b200 import CoverageEngine, MODE_DEPTH, OUT_BEDGRAPH, OUT_DEPTH, shard
    from highperformancengs_b200.gen import Generator

    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        # NCCL prints its version banner on stdout when NCCL_DEBUG is set in the environment; stdout is reserved for the
        # one JSON line, so fd 1 points at stderr while the communicator comes up
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            warm = torch.zeros(1, device="cuda")
            dist.all_reduce(warm)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- input: generate this rank's part of the stream straight into pinned host memory
    g = Generator(2, scale=args.scale, seq=False)
    lens, n_ref = g.lens, g.n_ref
    G = sum(lens)
    if world > 1:
        sl_beg, sl_end = shard.plan_slices(lens, world)[rank]
        cb, pb = shard.locate(lens, sl_beg)
        ce, pe = shard.locate(lens, sl_end)
        c0, c1 = cb, (ce if pe else ce - 1)                 # contigs [c0, c1] are touched
        last = rank == world - 1
        c_hi = n_ref + 1 if last else c1 + 1                 # the last rank also walks the unmapped tail
        total = g.records_size(c0, c_hi)
        to = g.tile_offsets(cb, c0)
        off_beg = int(to[max((pb >> 14) - 1, 0)]) if pb else 0
        if last:
            off_end = total
        elif pe == 0:
            off_end = total
        else:
            te = g.tile_offsets(ce, c0)
            off_end = int(te[pe >> 14]) if (pe >> 14) < len(te) else g.records_size(c0, ce + 1)
    else:
        sl_beg = sl_end = 0
        c0, c_hi, total = 0, n_ref + 1, g.records_size(0, n_ref + 1)
        off_beg, off_end = 0, total
    t0 = time.time()
    host = torch.empty(max(total, 1), dtype=torch.uint8, pin_memory=True)
    g.records_into(host.data_ptr(), total, c0, c_hi)
    t_gen = time.time() - t0
    nbytes = off_end - off_beg
    host_ptr = host.data_ptr() + off_beg
    dev = torch.empty(max(nbytes, 1), dtype=torch.uint8, device="cuda")
    dev.copy_(host[off_beg:off_end], non_blocking=False)
    torch.cuda.synchronize()

    eng = CoverageEngine(g.names, lens, mode=MODE_DEPTH, window=args.window, emit_wig=False, device=local,
                         slice_beg=sl_beg, slice_end=sl_end)
    stream = torch.cuda.ExternalStream(eng.stream, device=torch.device("cuda", local))
    mask_all = (1 << OUT_BEDGRAPH) | (1 << OUT_DEPTH)
    launches = [0]

    def stitch_and_format():
        """halo exchange (N>1): merge runs cut by shard boundaries, fix + sum the window bins"""
        if world == 1:
            eng.format(mask_all)
            return
        h = eng.halo()
        t = torch.tensor([h["n_runs"], *h["first"], *h["last"]], dtype=torch.int64, device="cuda")
        allh = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(allh, t)
        halos = []
        for x in allh:
            v = x.tolist()
            halos.append({"n_runs": v[0], "first": tuple(v[1:5]), "last": tuple(v[5:9])})
        decisions, corr = shard.stitch(halos)
        eng.halo_put(*decisions[rank])
        bins = eng.bins()
        if rank == 0:
            for run, sign in corr:
                eng.bins_apply_run(run, sign, bins)
        tb = torch.from_numpy(bins.view(np.int64)).cuda()
        dist.all_reduce(tb)
        if rank == 0:
            eng.bins_put(tb.cpu().numpy().view(np.uint64))
            eng.format(mask_all)
        else:
            eng.format(1 << OUT_BEDGRAPH)

    def step_resident():
        eng.reset()
        eng.push_device(dev.data_ptr() if nbytes else 0, nbytes, last=True)
        eng.finish_runs()
        stitch_and_format()
        launches[0] += eng.stats()["launches"]

    def step_e2e():
        eng.reset()
        eng.push_host_ptr(host_ptr, nbytes, last=True)
        eng.finish_runs()
        stitch_and_format()
        n = eng.output_view(OUT_BEDGRAPH).size
        if rank == 0:
            n += eng.output_view(OUT_DEPTH).size
        return n

    # ---- resident-input timing
    for _ in range(args.warmup):
        step_resident()
    barrier()
    launches[0] = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stage_ms = {"ingest": 0.0, "decode": 0.0, "scan": 0.0, "format": 0.0}
    with ClockSampler(local) as clk:
        e0.record(stream)
        for _ in range(args.steps):
            step_resident()
            st = eng.stats()
            stage_ms["ingest"] += st["ms_ingest"]; stage_ms["decode"] += st["ms_decode"]
            stage_ms["scan"] += st["ms_scan"]; stage_ms["format"] += st["ms_format"]
        e1.record(stream)
        barrier()
    ms_total = e0.elapsed_time(e1)
    st = eng.stats()
    out_bytes = sum(st["out_bytes"])
    tt = torch.tensor([ms_total, stage_ms["ingest"], stage_ms["scan"], stage_ms["format"], stage_ms["decode"]], dtype=torch.float64, device="cuda")
    ts = torch.tensor([st["records"], st["cigar_words"], st["positions"], out_bytes, st["runs"], launches[0], nbytes], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(ts, op=dist.ReduceOp.SUM)
    ms_total, ms_ing, ms_scan, ms_fmt, ms_dec = [x / 1.0 for x in tt.tolist()]
    records, cigar_words, positions, out_b, runs, n_launch, in_bytes = ts.tolist()
    ms_step = ms_total / args.steps
    value = g.n_reads / (ms_step * 1e-3)

    # ---- end-to-end timing (host buffers, H2D + D2H inside)
    def time_e2e(step_fn, h2d_bytes):
        d2h = 0
        for _ in range(min(args.warmup, 2)):
            d2h = step_fn()
        barrier()
        k = max(1, min(args.steps, 3))
        t0 = time.time()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record(stream)
        for _ in range(k):
            d2h = step_fn()
        f1.record(stream)
        barrier()
        ms_e = max(f0.elapsed_time(f1), (time.time() - t0) * 1e3) / k
        te = torch.tensor([ms_e, 0.0], dtype=torch.float64, device="cuda")
        tb2 = torch.tensor([float(h2d_bytes), float(d2h)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
            dist.all_reduce(tb2, op=dist.ReduceOp.SUM)
        return {"value": g.n_reads / (te[0].item() * 1e-3), "unit": "reads/s", "h2d_bytes_per_step": int(tb2[0].item()),
                "d2h_bytes_per_step": int(tb2[1].item()), "ms_per_step": te[0].item(), "steps": k}

    e2e = None
    if not args.no_e2e:
        e2e = time_e2e(step_e2e, nbytes)
        e2e["input"] = "inflated BAM record stream in pinned host memory (ngscov_push_host); host BGZF inflate is I/O, not timed"
        if world == 1 and not args.no_bgzf:
            # the same job from the BAM FILE's bytes: compressed BGZF blocks in pinned host memory, inflated on the device
            # (ngscov_push_bgzf).  This is what the reference arm does from its file, inflate included.
            from highperformancengs_b200 import bamio
            import zlib
            import shutil
            need = int(nbytes * 0.2) + (1 << 30)          # the BAM is ~1/8 of the stream at zlib level 1
            shm_ok = os.path.isdir("/dev/shm") and shutil.disk_usage("/dev/shm").free > need
            tdir = tempfile.mkdtemp(prefix="ngscov_bgzf_", dir="/dev/shm" if shm_ok else None)
            try:
                bam = os.path.join(tdir, "bench.bam")
                t0 = time.time()
                g.write_bam(bam, level=1)
                t_bam = time.time() - t0
                comp = torch.empty(os.path.getsize(bam), dtype=torch.uint8, pin_memory=True)
                with open(bam, "rb") as f:
                    f.readinto(memoryview(comp.numpy()))
            finally:
                shutil.rmtree(tdir, ignore_errors=True)
            inf, hdr_len = b"", None
            for off, bs in bamio.bgzf_blocks(memoryview(comp.numpy()[:8 << 20])):
                inf += zlib.decompress(bytes(comp.numpy()[off + 18:off + bs - 8]), -15)
                try:
                    hdr_len = bamio.parse_header(inf)[2]
                    break
                except Exception:
                    continue

            def step_bgzf():
                eng.reset()
                eng.push_bgzf_ptr(comp.data_ptr(), comp.numel(), skip=hdr_len, last=True)
                eng.finish_runs()
                stitch_and_format()
                return eng.output_view(OUT_BEDGRAPH).size + eng.output_view(OUT_DEPTH).size

            eb = time_e2e(step_bgzf, comp.numel())
            stb = eng.stats()
            eb["input"] = "the BAM file's BGZF blocks (zlib level 1) in pinned host memory, inflated on the device (ngscov_push_bgzf)"
            eb["device_inflate_ms"] = stb["ms_inflate"]
            eb["inflate_GBps_out"] = nbytes / max(stb["ms_inflate"], 1e-9) / 1e6
            eb["bam_write_seconds"] = round(t_bam, 2)
            eb["same_output_bytes"] = eb["d2h_bytes_per_step"] == e2e["d2h_bytes_per_step"]
            e2e = dict(eb, inflated_input=e2e)

    if args.check:
        # parity of the sharded job: concatenate the ranks' bedGraph in rank order, .depth from rank 0
        step_resident()
        bed = eng.output(OUT_BEDGRAPH)
        parts = [None] * world
        if world > 1:
            dist.all_gather_object(parts, bed)
        else:
            parts = [bed]
        if rank == 0:
            from tests import oracle_py
            subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "oracle"])
            exp = oracle_py.run(g.records(), g.names, lens, 0, args.window)
            ok = b"".join(parts) == exp["bedgraph"] and eng.output(OUT_DEPTH) == exp["depth"]
            print(f"check: {world} shard(s), bedGraph {sum(map(len, parts))} bytes, parity {'OK' if ok else 'FAILED'}", file=sys.stderr, flush=True)
            if not ok:
                raise SystemExit("sharded output differs from the oracle")

    if rank == 0:
        peak, peak_src = peaks()
        # algorithmic bytes: records*36 + 4*cigar_words  |  2*4*positions  |  output text
        # per kernel: K1 frames the records (reads 36 + 4*n_cigar bytes of each), K2 decodes them and
        # writes the difference array once, K3+K4 reads it once, K5 writes the text once
        b_frame = 36.0 * records + 4.0 * cigar_words
        b_dec = 4.0 * positions
        b_scan = 4.0 * positions
        b_fmt = float(out_b)
        b_ing = b_frame + b_dec
        stages = {"frame": (b_frame, (ms_ing - ms_dec) / args.steps), "decode": (b_dec, ms_dec / args.steps),
                  "scan": (b_scan, ms_scan / args.steps), "format": (b_fmt, ms_fmt / args.steps)}
        dom = max(stages, key=lambda k: stages[k][1])
        kernels = {"frame": "k_chain_walk", "decode": "k_decode_scatter", "scan": "k_scan_runs", "format": "k_format_bed"}   # frame = K1 + the three verify kernels
        ach = stages[dom][0] / (stages[dom][1] * 1e-3) / 1e9 / world if stages[dom][1] > 0 else 0.0
        b_all = b_ing + b_scan + b_fmt
        # DRAM traffic of the dominant kernel per launch from the committed ncu capture (profiles/r1_traffic.json,
        # full-size single-GPU workload only; null otherwise)
        traffic = None
        try:
            if world == 1 and args.scale == 1.0:
                tj = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))["kernels"]   # ncu launch list of this same command
                key = [k for k in tj if k.startswith(kernels[dom])][0]
                traffic = tj[key]["dram_read_bytes"] + tj[key]["dram_write_bytes"]
        except Exception:
            traffic = None
        line = {"metric": METRIC, "value": value, "unit": "reads/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int32",
                "data": "synthetic",
                "config": {"workload": WORKLOAD, "scale": args.scale, "reads": int(g.n_reads), "genome_bp": int(G), "window": args.window,
                           "parallelism": f"region-shard x{world}", "l2": "inputs (stream + difference array) far larger than the 126 MB L2; no flush needed",
                           "stream_bytes": int(in_bytes), "bases_per_s": value * 150.0, "gen_seconds": round(t_gen, 2)},
                "roofline": {"bound": "hbm", "kernel": kernels[dom], "stage": dom, "achieved": ach, "peak": peak, "unit": "GB/s",
                             "frac": ach / peak, "traffic": traffic, "peak_source": peak_src,
                             "algorithmic_bytes_per_launch": stages[dom][0] / world, "ms_per_launch": stages[dom][1]},
                "engine": {"algorithmic_bytes_per_step": b_all, "bytes_per_read": b_all / max(records, 1),
                           "hbm_frac_whole_step": b_all / (ms_step * 1e-3) / 1e9 / world / peak,
                           "stage_ms": {k: v[1] for k, v in stages.items()},
                           "stage_gbs": {k: (v[0] / (v[1] * 1e-3) / 1e9 / world if v[1] > 0 else 0.0) for k, v in stages.items()},
                           "runs": int(runs), "out_bytes": int(out_b)},
                "clocks": clk.summary(), "gpu_launches": int(n_launch)}
        if e2e:
            line["e2e"] = e2e
        if not args.no_cpu and world == 1:
            try:
                line["cpu_baseline"] = cpu_baseline(args.cpu_sample_reads)
            except Exception as ex:  # keep the GPU line even if the sample run breaks
                line["cpu_baseline"] = {"value": None, "unit": "reads/s", "cores": 1, "kind": "reference", "sample": f"failed: {ex}"}
        print(json.dumps(line), flush=True)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
