#!/usr/bin/env python
"""bench.py -- throughput of the coverage hot path (BAM record stream -> bedGraph + .depth / wig) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 1|2|4|5] [--scale S]

Workloads (BASELINE.json configs; default = config 2, the one the metric is quoted on):
  1  synthetic 1 M x 100 bp single-end reads on one 10 Mbp contig, records with SEQ/QUAL, bam2depth -w 20000
  2  synthetic 30x WGS, ~618 M x 150 bp reads over an hg38-shaped 24-contig 3.09 Gbp genome, records without SEQ/QUAL
     ("--seq none": the full-SEQ stream is 170 GB and cannot be HBM-resident next to its outputs; `--seq full --scale 0.25`
     measures the record shape real BAMs have), bam2depth -w 20000.  With N GPUs (config 3) the genome is region-sharded.
  4  synthetic RNA-seq-like BAM: spliced CIGARs (N/D/I/S), skewed exon coverage, bam2wig -w 20000 (wig + chromSize)
  5  targeted panel: 50 M reads piled into 200 kb of amplicons (depth > 10^4), bam2depth -w 20000
A step is one pass of the whole job.  With N GPUs one process per GPU owns a contiguous genome slice; the halo records
are gathered over NCCL (NVLink) on the engine stream, a one-block kernel applies the stitch rule on every shard, the window
bins are summed in place with one all_reduce.  Total work is fixed ("strong").

One JSON line is printed by rank 0:
  value     reads/s with the record stream already resident in HBM (CUDA events on the engine stream, max over ranks)
  e2e       the same job from the BAM file's compressed BGZF blocks in pinned host memory (ngscov_push_bgzf: H2D of the
            compressed bytes, device inflate, engine, text copied back to pinned host memory); e2e.inflated_input = the job fed
            with the already-inflated record stream (ngscov_push_host)
  roofline  the slowest kernel's algorithmic bytes / its device time vs the measured HBM copy peak
  parity    a check of THIS arm before the timed region: the same sharded job at a small scale compared byte for byte with
            the CPU oracle, plus a 64-bit position-weighted hash of the full-size rank-ordered output (equal for every N);
            for the full-size config-2 job that hash is compared with the one of the files the UNMODIFIED reference wrote
            for the same BAM (parity.reference_full_size, tests/golden/config2_full_reference.json): a mismatch exits 1
  cpu_baseline  the unmodified reference tool (oracle/_ref) timed on one core: a bounded sample for config 2, the full input
            for configs 1/4/5
  cli_wall  (config 1, and config 2 at --cli-scale) wall seconds of the drop-in C driver host/bam2depth and of the reference
            binary on the same BAM file in /dev/shm
`--impl reference` times the reference tool itself on all host cores (one process per core: the tool is single-threaded).
"""
import argparse
import json
import os
import shutil
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "aligned reads/s BAM->bedGraph depth"
CONFIGS = {
    1: dict(mode=0, tool="bam2depth", seq=True, what="config1: synthetic 1 M x 100 bp single-end reads on one 10 Mbp contig, records with SEQ/QUAL"),
    2: dict(mode=0, tool="bam2depth", seq=False, what="config2/3: synthetic 30x WGS, 150 bp reads over hg38-shaped 24 contigs (3.09 Gbp)"),
    4: dict(mode=1, tool="bam2wig", seq=False, what="config4: synthetic RNA-seq-like BAM, spliced CIGARs (N/D/I/S), skewed exon coverage"),
    5: dict(mode=0, tool="bam2depth", seq=False, what="config5: targeted panel, 50 M reads piled into 200 kb of amplicons (depth > 10^4)"),
}
M64 = (1 << 64) - 1


def workload(cfg, seq, window):
    c = CONFIGS[cfg]
    return (f"{c['what']}, records {'with' if seq else 'without'} SEQ/QUAL (--seq {'full' if seq else 'none'}), "
            f"{c['tool']} -w {window} -> {'bedGraph + .depth' if c['mode'] == 0 else 'wig + chromSize'}")


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
FULL_READS = {1: 1000000, 2: 617654000, 4: 20000000, 5: 50000000}


def make_bam(path, cfg, scale, seed, seq):
    """write a BAM + .bai of one of the synthetic configurations with tools/bamgen; returns the number of reads"""
    from highperformancengs_b200 import gen
    gen.build()
    subprocess.check_call([os.path.join(ROOT, "tools", "bamgen"), "--config", str(cfg), "--scale", repr(scale), "--seed", str(seed),
                           "--seq", "full" if seq else "none", "-o", path], stderr=subprocess.DEVNULL)
    from highperformancengs_b200.gen import Generator
    return Generator(cfg, scale=scale, seed=seed, seq=seq).n_reads


def ref_tool(tool):
    p = os.path.join(ROOT, "oracle", "_ref", tool)
    if os.path.exists(p):
        return [p], "reference"
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "oracle"])
    return [os.path.join(ROOT, "oracle", "_build", "ngscov_oracle"), "depth" if tool == "bam2depth" else "wig"], "port"


def run_reference_once(cmd, bams, workdir, window):
    """one process per BAM entry, concurrently; returns wall seconds"""
    t0 = time.time()
    procs = [subprocess.Popen(cmd + ["-o", os.path.join(workdir, f"o{i}"), "-w", str(window), b], cwd=workdir,
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL) for i, b in enumerate(bams)]
    for p in procs:
        if p.wait() != 0:
            raise RuntimeError("reference tool failed")
    return time.time() - t0


def cpu_baseline(cfg, seq, window, sample_reads):
    cmd, kind = ref_tool(CONFIGS[cfg]["tool"])
    with tempfile.TemporaryDirectory(prefix="ngscov_cpu_") as d:
        bam = os.path.join(d, "sample.bam")
        if cfg == 2:
            n = make_bam(bam, 2, sample_reads / FULL_READS[2], 11, seq)
            what = f"{n} reads of the same 30x WGS generator (scaled genome)"
        else:
            n = make_bam(bam, cfg, 1.0, 0, seq)
            what = f"the full config-{cfg} input, {n} reads"
        dt = run_reference_once(cmd, [bam], d, window)
    return {"value": n / dt, "unit": "reads/s", "cores": 1, "kind": kind,
            "sample": f"{what}, one {CONFIGS[cfg]['tool']} process, {dt:.2f} s wall"}


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg, seq = args.config, args.seq_bool
    cmd, kind = ref_tool(CONFIGS[cfg]["tool"])
    cores = os.cpu_count() or 1
    with tempfile.TemporaryDirectory(prefix="ngscov_ref_") as d:
        bams, n_tot = [], 0
        if cfg == 2:
            per = args.ref_sample_reads
            for i in range(cores):
                b = os.path.join(d, f"s{i}.bam")
                n_tot += make_bam(b, 2, per / FULL_READS[2], 100 + i, seq)
                bams.append(b)
            sample = f"{cores} concurrent single-threaded processes x {per} reads each per step"
        else:
            # configs 1/4/5: the reference runs the FULL input; `cores` concurrent copies use all host threads
            b = os.path.join(d, "full.bam")
            n1 = make_bam(b, cfg, args.scale, 0, seq)
            bams, n_tot = [b] * cores, n1 * cores
            sample = f"{cores} concurrent single-threaded processes, each on the full config-{cfg} input ({n1} reads)"
        for _ in range(args.warmup):
            run_reference_once(cmd, bams, d, args.window)
        t = [run_reference_once(cmd, bams, d, args.window) for _ in range(args.steps)]
    dt = sum(t) / len(t)
    v = n_tot / dt
    line = {"metric": METRIC, "value": v, "unit": "reads/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int32",
            "data": "synthetic", "impl": "reference",
            "config": {"workload": workload(cfg, seq, args.window), "sample": sample},
            "cpu_baseline": {"value": v, "unit": "reads/s", "cores": cores, "kind": kind,
                             "sample": f"{n_tot} reads per step: {cores} processes (the tool is single-threaded)"},
            "e2e": {"value": v, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def full_size_gate(ph, n_reads, gold_path=None):
    """The full-size config-2 job is pinned against the UNMODIFIED reference binary run on the same generated BAM
    (tests/golden/config2_full_reference.json, made by tests/golden/make_config2_full_golden.py): the hash bench.py
    computes on the device over the engine's output must be the hash of the files the reference wrote.  ph = parity["hash"]."""
    gold_path = gold_path or os.path.join(ROOT, "tests", "golden", "config2_full_reference.json")
    if not os.path.exists(gold_path):
        return None
    gold = json.load(open(gold_path))
    ok = (int(gold["reads"]) == int(n_reads) and gold["bedgraph"]["hash"] == ph["bedgraph"] and int(gold["bedgraph"]["bytes"]) == int(ph["bedgraph_bytes"])
          and gold["depth"]["hash"] == ph["aux"])
    return {"match": bool(ok), "bedgraph_md5": gold["bedgraph"]["md5"], "reference_seconds": gold.get("reference_seconds"),
            "what": "bedGraph and .depth of oracle/_ref/bam2depth (unmodified reference) on the same 618 M-read BAM: size and hash"}


# --------------------------------------------------------------------------------------- our arm
class Job:
    """one rank's share of a job: generator -> pinned host stream -> device copy, engine, step functions"""

    def __init__(self, cfg, scale, seq, window, rank, world, local, staging_mb=0):
        import torch
        from highperformancengs_b200 import CoverageEngine, OUT_BEDGRAPH, OUT_DEPTH, OUT_WIG, shard
        from highperformancengs_b200.gen import Generator
        self.torch, self.rank, self.world, self.local = torch, rank, world, local
        self.cfg, self.mode, self.window = cfg, CONFIGS[cfg]["mode"], window
        g = self.g = Generator(cfg, scale=scale, seq=seq)
        lens, n_ref = g.lens, g.n_ref
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
            if last or pe == 0:
                off_end = total
            else:
                te = g.tile_offsets(ce, c0)
                off_end = int(te[pe >> 14]) if (pe >> 14) < len(te) else g.records_size(c0, ce + 1)
        else:
            sl_beg = sl_end = 0
            c0, c_hi, total = 0, n_ref + 1, g.records_size(0, n_ref + 1)
            off_beg, off_end = 0, total
        self.c0, self.off_beg, self.off_end = c0, off_beg, off_end
        t0 = time.time()
        self.host = torch.empty(max(total, 1), dtype=torch.uint8, pin_memory=True)
        g.records_into(self.host.data_ptr(), total, c0, c_hi)
        self.t_gen = time.time() - t0
        self.nbytes = off_end - off_beg
        self.host_ptr = self.host.data_ptr() + off_beg
        self.dev = torch.empty(max(self.nbytes, 1), dtype=torch.uint8, device="cuda")
        self.dev.copy_(self.host[off_beg:off_end], non_blocking=False)
        torch.cuda.synchronize()
        self.eng = CoverageEngine(g.names, lens, mode=self.mode, window=window, emit_wig=False, device=local, slice_beg=sl_beg, slice_end=sl_end,
                                  staging_bytes=int(staging_mb) << 20)
        self.stream = torch.cuda.ExternalStream(self.eng.stream, device=torch.device("cuda", local))
        self.BED, self.AUX = OUT_BEDGRAPH, (OUT_DEPTH if self.mode == 0 else OUT_WIG)
        self.mask_all = ((1 << OUT_BEDGRAPH) | (1 << OUT_DEPTH)) if self.mode == 0 else ((1 << OUT_WIG) | (1 << 3))
        self.launches = 0
        self.allh = torch.zeros(9 * world, dtype=torch.int64, device="cuda") if world > 1 else None
        self.comp = None

    # ---- halo exchange (N>1): everything stays on the device and on the engine stream
    def stitch_and_format(self):
        eng, torch = self.eng, self.torch
        if self.world == 1:
            eng.format(self.mask_all)
            return
        import torch.distributed as dist
        from highperformancengs_b200.engine import DeviceArray
        with torch.cuda.stream(self.stream):
            mine = torch.as_tensor(DeviceArray(eng.halo_device(), 9), device="cuda")
            dist.all_gather_into_tensor(self.allh, mine)                       # 72 bytes per rank over NVLink
            eng.stitch_device(self.allh.data_ptr(), self.world, self.rank)        # one kernel: text patch + bin corrections
            bins = torch.as_tensor(DeviceArray(*eng.bins_device()), device="cuda")
            dist.all_reduce(bins)                                              # window bins summed in place
        if self.rank == 0:
            eng.format(self.mask_all)
        elif self.mode == 0:
            eng.format(1 << self.BED)

    def step_resident(self):
        eng = self.eng
        eng.reset()
        eng.push_device(self.dev.data_ptr() if self.nbytes else 0, self.nbytes, last=True)
        eng.finish_runs()
        self.stitch_and_format()
        self.launches += eng.stats()["launches"]

    def fetch(self):
        """device -> pinned host read of this rank's outputs; returns the bytes copied"""
        n = self.eng.output_view(self.BED).size if self.mode == 0 else 0
        if self.rank == 0:
            n += self.eng.output_view(self.AUX).size
        return n

    def step_e2e_host(self):
        eng = self.eng
        eng.reset()
        eng.push_host_ptr(self.host_ptr, self.nbytes, last=True)
        eng.finish_runs()
        self.stitch_and_format()
        return self.fetch()

    # ---- the same job from the BAM file's bytes
    def prepare_bgzf(self, barrier):
        """rank 0 writes the BAM (zlib level 1) once; every rank keeps the BGZF blocks that cover its slice of the record
        stream in pinned memory, with the in-block offset of its first record as `skip`"""
        import numpy as np
        import zlib
        from highperformancengs_b200 import bamio
        torch, g = self.torch, self.g
        path = os.environ.get("NGSCOV_BENCH_BAM") or os.path.join("/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir(),
                                                                   f"ngscov_bench_{os.environ.get('MASTER_PORT', '0')}_{self.cfg}.bam")
        t0 = time.time()
        if self.rank == 0:
            need = int(g.records_size(0, g.n_ref + 1) * 0.2) + (1 << 30)
            if shutil.disk_usage(os.path.dirname(path)).free >= need:
                try:                      # the other ranks wait at the barrier: rank 0 may use every core for the deflate
                    import ctypes
                    gomp = ctypes.CDLL("libgomp.so.1")
                    gomp.omp_set_num_threads(os.cpu_count() or 1)
                    try:
                        os.sched_setaffinity(0, range(os.cpu_count() or 1))
                    except Exception:
                        pass
                except Exception:
                    gomp = None
                g.write_bam(path, level=1)
        barrier()
        if not os.path.exists(path):
            return None
        self.t_bam = time.time() - t0
        try:
            size = os.path.getsize(path)
            mm = np.memmap(path, dtype=np.uint8, mode="r")
            # header length from the first blocks
            inf, hdr_len = b"", None
            for off, bs in bamio.bgzf_blocks(memoryview(mm[:min(size, 8 << 20)])):
                inf += zlib.decompress(bytes(mm[off + 18:off + bs - 8]), -15)
                try:
                    hdr_len = bamio.parse_header(inf)[2]
                    break
                except Exception:
                    continue
            if self.world == 1:
                c_beg, c_end, skip = 0, size, hdr_len
            else:
                # block table: file offset and inflated offset of every block (BSIZE at +16, ISIZE in the last 4 bytes)
                offs, uoffs, off, u = [], [], 0, 0
                while off + 18 <= size:
                    bs = int(mm[off + 16]) + (int(mm[off + 17]) << 8) + 1
                    isz = int.from_bytes(bytes(mm[off + bs - 4:off + bs]), "little")
                    offs.append(off); uoffs.append(u)
                    off += bs; u += isz
                offs.append(off); uoffs.append(u)
                uo = np.asarray(uoffs)
                base = hdr_len + g.records_size(0, self.c0)            # inflated offset of this rank's contig range
                a, b = base + self.off_beg, base + self.off_end
                i0 = int(np.searchsorted(uo, a, side="right") - 1)
                i1 = int(np.searchsorted(uo, b, side="left"))
                c_beg, c_end, skip = offs[i0], offs[min(i1, len(offs) - 1)], a - uoffs[i0]
            self.comp = torch.empty(max(c_end - c_beg, 1), dtype=torch.uint8, pin_memory=True)
            self.comp.numpy()[:c_end - c_beg] = mm[c_beg:c_end]
            self.comp_n, self.skip = c_end - c_beg, int(skip)
            del mm
        finally:
            barrier()
            if self.rank == 0 and not os.environ.get("NGSCOV_BENCH_BAM"):
                os.unlink(path)
        return self.comp_n

    def step_e2e_bgzf(self):
        eng = self.eng
        eng.reset()
        eng.push_bgzf_ptr(self.comp.data_ptr(), self.comp_n, skip=self.skip, last=True, truncated_ok=self.world > 1)
        eng.finish_runs()
        self.stitch_and_format()
        return self.fetch()

    # ---- outputs for the checks
    def texts(self):
        bed = self.eng.output(self.BED) if self.mode == 0 else b""
        aux = self.eng.output(self.AUX) if self.rank == 0 else b""
        return bed, aux

    def device_hash(self, which, base):
        """sum over bytes b_j * ((global offset j + 1) * M mod 2^64) mod 2^64 of one device-resident output"""
        torch = self.torch
        ptr, n = self.eng.output_device(which)
        if not n:
            return 0
        from highperformancengs_b200.engine import DeviceArray
        t = torch.as_tensor(DeviceArray(ptr, n, "|u1"), device="cuda")
        M = -7046029254386353131                                   # 0x9E3779B97F4A7C15 as int64
        h, step = 0, 1 << 26
        for o in range(0, n, step):
            x = t[o:o + step].to(torch.int64)
            w = (torch.arange(base + o + 1, base + o + 1 + x.numel(), dtype=torch.int64, device="cuda")) * M
            h = (h + int((x * w).sum().item())) & M64
        return h

    def close(self):
        self.eng.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--config", type=int, default=2, choices=[1, 2, 4, 5])
    ap.add_argument("--scale", type=float, default=float(os.environ.get("NGSCOV_BENCH_SCALE", "1.0")))
    ap.add_argument("--seq", default=None, choices=["none", "full"], help="records without / with SEQ+QUAL (default: the config's own)")
    ap.add_argument("--window", type=int, default=20000)
    ap.add_argument("--staging-mb", type=int, default=int(os.environ.get("NGSCOV_BENCH_STAGING_MB", "0")),
                    help="cfg.staging_bytes of the engine: the chunk the pushed stream is processed in (0: the engine's default, 256 MiB)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-bgzf", action="store_true", help="skip the end-to-end leg that starts from the compressed BAM bytes")
    ap.add_argument("--no-check", action="store_true", help="skip the small-scale comparison with the CPU oracle before the timed region")
    ap.add_argument("--no-cli", action="store_true", help="skip the wall-clock run of the C driver / reference binary")
    ap.add_argument("--check-scale", type=float, default=None)
    ap.add_argument("--cli-scale", type=float, default=0.02, help="config 2: scale of the BAM the two command-line tools are timed on")
    ap.add_argument("--ref-sample-reads", type=int, default=1000000)
    ap.add_argument("--cpu-sample-reads", type=int, default=2000000)
    args = ap.parse_args()
    args.seq_bool = CONFIGS[args.config]["seq"] if args.seq is None else (args.seq == "full")
    if args.impl == "reference":
        return reference_arm(args)

    import numpy as np
    import torch
    import torch.distributed as dist

    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    numa = bind_to_gpu_numa(local)
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

    cfg, seq = args.config, args.seq_bool

    # ---- parity gate of THIS arm (same N, same step functions) at a scale the CPU oracle can do
    parity = {"oracle_check": None}
    if not args.no_check:
        cs = args.check_scale if args.check_scale is not None else {1: 0.2, 2: 0.004, 4: 0.05, 5: 0.01}[cfg]
        cs = min(cs, args.scale)
        small = Job(cfg, cs, seq, args.window, rank, world, local)
        small.step_resident()
        bed, aux = small.texts()
        parts = [None] * world
        if world > 1:
            dist.all_gather_object(parts, bed)
        else:
            parts = [bed]
        ok = True
        if rank == 0:
            from tests import oracle_py
            subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "oracle"])
            exp = oracle_py.run(small.g.records(), small.g.names, small.g.lens, small.mode, args.window)
            if small.mode == 0:
                ok = b"".join(parts) == exp["bedgraph"] and aux == exp["depth"]
            else:
                ok = aux == exp["wig"]
            parity["oracle_check"] = {"scale": cs, "reads": int(small.g.n_reads), "shards": world, "ok": bool(ok),
                                      "what": "bedGraph (rank-ordered concatenation) and .depth" if small.mode == 0 else "wig"}
            print(f"check: {world} shard(s) at scale {cs}: parity {'OK' if ok else 'FAILED'}", file=sys.stderr, flush=True)
        small.close()
        del small
        okt = torch.tensor([1 if ok else 0], device="cuda")
        if world > 1:
            dist.broadcast(okt, 0)
        if not int(okt.item()):
            raise SystemExit("sharded output differs from the oracle")

    job = Job(cfg, args.scale, seq, args.window, rank, world, local, staging_mb=args.staging_mb)
    g, eng, stream = job.g, job.eng, job.stream
    G = sum(g.lens)

    # ---- resident-input timing
    for _ in range(args.warmup):
        job.step_resident()
    barrier()
    job.launches = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stage_ms = {"ingest": 0.0, "decode": 0.0, "scan": 0.0, "format": 0.0}
    with ClockSampler(local) as clk:
        e0.record(stream)
        for _ in range(args.steps):
            job.step_resident()
            st = eng.stats()
            stage_ms["ingest"] += st["ms_ingest"]; stage_ms["decode"] += st["ms_decode"]
            stage_ms["scan"] += st["ms_scan"]; stage_ms["format"] += st["ms_format"]
        e1.record(stream)
        barrier()
    ms_total = e0.elapsed_time(e1)
    st = eng.stats()
    out_bytes = sum(st["out_bytes"])
    tt = torch.tensor([ms_total, stage_ms["ingest"], stage_ms["scan"], stage_ms["format"], stage_ms["decode"]], dtype=torch.float64, device="cuda")
    ts = torch.tensor([st["records"], st["cigar_words"], st["positions"], out_bytes, st["runs"], job.launches, job.nbytes, st["out_bytes"][0]], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(ts, op=dist.ReduceOp.SUM)
    ms_total, ms_ing, ms_scan, ms_fmt, ms_dec = tt.tolist()
    records, cigar_words, positions, out_b, runs, n_launch, in_bytes, bed_b = ts.tolist()
    ms_step = ms_total / args.steps
    value = g.n_reads / (ms_step * 1e-3)

    # ---- 64-bit hash of the rank-ordered output of the last resident step (equal for every N on the same generator)
    lens_t = torch.tensor([eng.output_device(job.BED)[1] if job.mode == 0 else 0], dtype=torch.int64, device="cuda")
    alll = [torch.zeros_like(lens_t) for _ in range(world)]
    if world > 1:
        dist.all_gather(alll, lens_t)
    else:
        alll = [lens_t]
    base = int(sum(int(x.item()) for x in alll[:rank]))
    hb = job.device_hash(job.BED, base) if job.mode == 0 else 0
    ht = torch.tensor([hb - (1 << 64) if hb >= (1 << 63) else hb], dtype=torch.int64, device="cuda")
    if world > 1:
        dist.all_reduce(ht)                                    # int64 sum wraps modulo 2^64
    parity["hash"] = {"bedgraph": format(int(ht.item()) & M64, "016x") if job.mode == 0 else None,
                      "aux": format(job.device_hash(job.AUX, 0), "016x") if rank == 0 else None,
                      "aux_output": ".depth" if job.mode == 0 else "wig",
                      "bedgraph_bytes": int(sum(int(x.item()) for x in alll)),
                      "rule": "sum_j byte_j * ((j + 1) * 0x9E3779B97F4A7C15 mod 2^64) mod 2^64 over the rank-ordered concatenation"}
    if rank == 0 and cfg == 2 and args.scale == 1.0 and args.window == 20000:
        gate = full_size_gate(parity["hash"], g.n_reads)
        if gate is not None:
            parity["reference_full_size"] = gate
            if not gate["match"]:
                print("bench: PARITY FAILURE: the output of the measured job differs from the reference's (tests/golden/config2_full_reference.json)", file=sys.stderr)

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
    probe = None
    if not args.no_e2e:
        probe = pcie_probe(torch, dist, world)
        probe["numa_binding"] = numa
        e2e = time_e2e(job.step_e2e_host, job.nbytes)
        e2e["input"] = "inflated BAM record stream in pinned host memory (ngscov_push_host); host BGZF inflate is I/O, not timed"
        if not args.no_bgzf:
            # the same job from the BAM FILE's bytes: compressed BGZF blocks in pinned host memory, inflated on the device
            # (ngscov_push_bgzf).  This is what the reference arm does from its file, inflate included.
            if job.prepare_bgzf(barrier) is not None:
                eb = time_e2e(job.step_e2e_bgzf, job.comp_n)
                stb = eng.stats()
                eb["input"] = ("the BAM file's BGZF blocks (zlib level 1) in pinned host memory, inflated on the device (ngscov_push_bgzf)"
                               + ("; every rank takes the blocks that cover its genome slice" if world > 1 else ""))
                eb["device_inflate_ms"] = stb["ms_inflate"]
                eb["chunk_mib"] = args.staging_mb or 256
                # device time of the last step's kernels by stage (CUDA events inside the engine): what the copies hide
                eb["device_ms"] = {"inflate": stb["ms_inflate"], "frame+classify": stb["ms_ingest"], "cover": stb["ms_scan"],
                                   "format_bins": stb["ms_format"], "h2d_copies": stb["ms_h2d"]}
                eb["bam_write_seconds"] = round(job.t_bam, 2)
                eb["same_output_bytes"] = eb["d2h_bytes_per_step"] == e2e["d2h_bytes_per_step"]
                e2e = dict(eb, inflated_input=e2e)

    if rank == 0:
        peak, peak_src = peaks()
        # algorithmic bytes: records*36 + 4*cigar_words  |  2*4*positions  |  output text.
        # per kernel: K1 k_chain_classify frames + decodes the records (reads 36 + 4*n_cigar bytes of
        # each); k_special_reds handles the rare records; k_cover owns what used to be "difference array written once, read
        # once" (2*4 bytes per position, now kept in shared memory) plus the bedGraph text written once; the .depth / wig
        # text of the window bins is the small format kernel
        b_frame = 36.0 * records + 4.0 * cigar_words
        b_cover = 8.0 * positions + float(bed_b)
        b_fmt = float(out_b) - float(bed_b)
        stages = {"frame+classify": (b_frame, (ms_ing - ms_dec) / args.steps), "special": (0.0, ms_dec / args.steps),
                  "cover": (b_cover, ms_scan / args.steps), "format_bins": (b_fmt, ms_fmt / args.steps)}
        dom = max(stages, key=lambda k: stages[k][1])
        kernels = {"frame+classify": "k_chain_classify", "special": "k_special_reds", "cover": "k_cover", "format_bins": "k_format_lines"}
        ach = stages[dom][0] / (stages[dom][1] * 1e-3) / 1e9 / world if stages[dom][1] > 0 else 0.0
        b_all = b_frame + 8.0 * positions + float(out_b)
        traffic = None
        try:
            if world == 1 and args.scale == 1.0 and cfg == 2 and not seq:
                tj = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json")))["kernels"]   # ncu launch list of this same command
                key = [k for k in tj if k.startswith(kernels[dom])][0]
                traffic = tj[key]["dram_read_bytes"] + tj[key]["dram_write_bytes"]
        except Exception:
            traffic = None
        line = {"metric": METRIC, "value": value, "unit": "reads/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int32",
                "data": "synthetic",
                "config": {"workload": workload(cfg, seq, args.window), "baseline_config": cfg if world == 1 or cfg != 2 else 3, "scale": args.scale,
                           "reads": int(g.n_reads), "genome_bp": int(G), "window": args.window,
                           "parallelism": f"region-shard x{world}", "l2": "inputs (record stream, record table) far larger than the 126 MB L2; no flush needed",
                           "stream_bytes": int(in_bytes), "bases_per_s": value * (100.0 if cfg == 1 else 150.0), "gen_seconds": round(job.t_gen, 2)},
                "roofline": {"bound": "hbm", "kernel": kernels[dom], "stage": dom, "achieved": ach, "peak": peak, "unit": "GB/s",
                             "frac": ach / peak, "traffic": traffic, "peak_source": peak_src,
                             "algorithmic_bytes_per_launch": stages[dom][0] / world, "ms_per_launch": stages[dom][1]},
                "engine": {"algorithmic_bytes_per_step": b_all, "bytes_per_read": b_all / max(records, 1),
                           "hbm_frac_whole_step": b_all / (ms_step * 1e-3) / 1e9 / world / peak,
                           "stage_ms": {k: v[1] for k, v in stages.items()},
                           "stage_gbs": {k: (v[0] / (v[1] * 1e-3) / 1e9 / world if v[1] > 0 else 0.0) for k, v in stages.items()},
                           "runs": int(runs), "out_bytes": int(out_b)},
                "parity": parity,
                "clocks": clk.summary(), "gpu_launches": int(n_launch)}
        if e2e:
            e2e["pcie_probe"] = probe
            line["e2e"] = e2e
        if not args.no_cpu and world == 1:
            try:
                line["cpu_baseline"] = cpu_baseline(cfg, seq, args.window, args.cpu_sample_reads)
            except Exception as ex:  # keep the GPU line even if the sample run breaks
                line["cpu_baseline"] = {"value": None, "unit": "reads/s", "cores": 1, "kind": "reference", "sample": f"failed: {ex}"}
        if not args.no_cli and world == 1 and cfg in (1, 2):
            try:
                line["cli_wall"] = cli_wall(cfg, seq, args.window, 1.0 if cfg == 1 else min(args.cli_scale, args.scale))
            except Exception as ex:
                line["cli_wall"] = {"failed": str(ex)}
        print(json.dumps(line), flush=True)
    job.close()
    if world > 1:
        dist.destroy_process_group()
    if rank == 0 and parity.get("reference_full_size", {}).get("match") is False:
        sys.exit(1)                                   # a fast wrong answer is not a benchmark result


def bind_to_gpu_numa(local):
    """Pin this process (and with it the pinned host buffers it allocates from now on) to the CPUs of the NUMA node the
    GPU hangs off: with one process per GPU and unbound processes, half of the device<->host copies cross the socket
    interconnect.  Also undoes torchrun's OMP_NUM_THREADS=1 for the input generator."""
    info = {"node": None}
    try:
        bus = subprocess.run(["nvidia-smi", f"--id={local}", "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                             capture_output=True, text=True, timeout=10).stdout.strip().lower()
        if bus.startswith("00000000:"):
            bus = bus[4:]
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        info = {"node": None, "sysfs_numa_node": node}
        if node >= 0:
            cpus = set()
            for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
            cpus &= os.sched_getaffinity(0)
            if cpus:
                os.sched_setaffinity(0, cpus)
                info = {"node": node, "cpus": len(cpus)}
    except Exception as ex:
        info = {"node": None, "error": str(ex)[:80]}
    world = int(os.environ.get("WORLD_SIZE", "1"))
    nthr = max(1, len(os.sched_getaffinity(0)) // (world if info.get("node") is None else 1))
    os.environ["OMP_NUM_THREADS"] = str(nthr)
    try:                                  # libgomp is already initialised (torch loaded it) with torchrun's OMP_NUM_THREADS=1
        import ctypes
        ctypes.CDLL("libgomp.so.1").omp_set_num_threads(nthr)
    except Exception:
        pass
    info["omp_threads"] = nthr
    return info


def pcie_probe(torch, dist, world, nbytes=1 << 30):
    """plain pinned H2D / D2H copy bandwidth of every rank, all ranks at once: the ceiling of the end-to-end legs"""
    h = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    d = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    res = []
    for src, dst in ((h, d), (d, h)):
        dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            dst.copy_(src, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        res.append(3 * nbytes / (e0.elapsed_time(e1) * 1e-3) / 1e9)
    t = torch.tensor(res, dtype=torch.float64, device="cuda")
    lo, hi = t.clone(), t.clone()
    if world > 1:
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    return {"h2d_GBps_per_rank_min": round(lo[0].item(), 1), "h2d_GBps_per_rank_max": round(hi[0].item(), 1),
            "d2h_GBps_per_rank_min": round(lo[1].item(), 1), "d2h_GBps_per_rank_max": round(hi[1].item(), 1),
            "what": f"{nbytes >> 20} MiB pinned copies, {world} rank(s) concurrently"}


def cli_wall(cfg, seq, window, scale):
    """what the user of the drop-in sees: wall seconds of host/bam2depth (this repo's C driver, device inflate) and of the
    unmodified reference binary on the same BAM file in /dev/shm, outputs compared byte for byte"""
    tool = CONFIGS[cfg]["tool"]
    ours = os.path.join(ROOT, "highperformancengs_b200", "host", tool)
    ref, kind = ref_tool(tool)
    base = "/dev/shm" if os.path.isdir("/dev/shm") else None
    with tempfile.TemporaryDirectory(prefix="ngscov_cli_", dir=base) as d:
        bam = os.path.join(d, "in.bam")
        n = make_bam(bam, cfg, scale, 0, seq)
        res = {"reads": int(n), "bam_bytes": os.path.getsize(bam), "scale": scale}
        outs = {}
        for name, cmd, env in (("ours", [ours], dict(os.environ, NGSCOV_GPU_INFLATE="1")), ("reference", ref, dict(os.environ))):
            od = os.path.join(d, name)
            os.makedirs(od)
            best = None
            for _ in range(2 if name == "ours" else 1):              # the first run of ours pays CUDA context creation; keep the second
                t0 = time.time()
                subprocess.check_call(cmd + ["-o", "o", "-w", str(window), bam], cwd=od, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, env=env)
                best = time.time() - t0
            res[f"{name}_wall_s"] = round(best, 3)
            outs[name] = {f: open(os.path.join(od, f), "rb").read() for f in sorted(os.listdir(od))}
        res["same_files"] = outs["ours"] == outs["reference"]
        res["output_bytes"] = sum(len(v) for v in outs["ours"].values())
        res["reference_kind"] = kind
        res["speedup"] = round(res["reference_wall_s"] / res["ours_wall_s"], 2)
    return res


if __name__ == "__main__":
    main()
