"""Python entry points with the reference tools' command lines (reference bam2depth.c:257-355,
bam2wig.c:254-336): same getopt strings, defaults, ignored -r/-s, output file names, truncation and
"-"-prefixed-name-means-stdout rule (IO_stream.h:69-97), .bai requirement and stderr progress lines.
The C drivers in host/ are the production drop-ins; these call the same C ABI from Python.
"""
import getopt
import os
import sys
import time

from . import bamio
from .engine import (CoverageEngine, MODE_DEPTH, MODE_WIG, OUT_BEDGRAPH, OUT_DEPTH, OUT_WIG, OUT_CHROMSIZE)

USAGE = """
Usage: %s [-o OUTFILE] [-w WINDOW_SIZE] [-r chr1:1-2000000]%s [-s 0] [-h] bamFile1 bamFile2 ..
  B200-native drop-in for the HighPerformanceNGS coverage tools.
   [-o OUTPUT_FILE]  = output prefix.                                  [required]
   [-w WINDOW_SIZE]  = window size, default 20000.                     [option]
%s   [-r]              = accepted and ignored, like the reference.       [option]
   [-s]              = accepted and ignored, like the reference.       [option]
   [-h]              = this help.                                      [option]
   Infiles           = coordinate-sorted, indexed BAM file(s).         [required]
"""


def _atoi(s):
    """C atoi: leading whitespace, optional sign, digits; anything else ends the number."""
    s = s.lstrip()
    i, sign = 0, 1
    if s[:1] in "+-":
        sign = -1 if s[0] == "-" else 1
        i = 1
    v = 0
    while i < len(s) and s[i].isdigit():
        v = v * 10 + int(s[i])
        i += 1
    return sign * v


def _usage(prog, depth_mode):
    sys.stderr.write(USAGE % (prog, " [-W is wig]" if depth_mode else "",
                              "   [-W is wig]       = also write wig + chromSize, default not.        [option]\n" if depth_mode else ""))
    raise SystemExit(1)


def _open_out(name):
    """fopen_output_stream: names that start with '-' (or are empty) mean stdout; else O_TRUNC create."""
    if name.startswith("-") or name == "":
        return os.fdopen(os.dup(1), "wb")
    return open(name, "wb")


def _run(argv, depth_mode):
    prog = argv[0]
    if len(argv) < 2:
        _usage(prog, depth_mode)
    try:
        opts, files = getopt.getopt(argv[1:], "o:w:r:s:Wh?" if depth_mode else "o:w:r:s:h?")
    except getopt.GetoptError:
        _usage(prog, depth_mode)
    outfile, window, wig = "-", 20000, 0
    for k, v in opts:
        if k == "-o":
            outfile = v
        elif k == "-w":
            window = _atoi(v) & 0xFFFFFFFF
        elif k == "-W":
            wig += 1
        elif k in ("-h", "-?"):
            _usage(prog, depth_mode)
        # -r and -s are parsed and never used (reference bam2depth.c:281-285)
    begin = time.time()
    for i, fn in enumerate(files):
        if not os.path.exists(fn):
            sys.stderr.write(f"{os.path.basename(prog)}: bam2bed: Fail to open BAM file {fn}\n")
            raise SystemExit(1)
        n = i + 1
        streams = {}
        if depth_mode:
            streams[OUT_BEDGRAPH] = _open_out(f"{os.path.basename(fn)}.{n}.bedGraph")
            streams[OUT_DEPTH] = _open_out(f"{outfile}.{n}.depth")
        if wig or not depth_mode:
            streams[OUT_WIG] = _open_out(f"{outfile}.{n}.wig")
            streams[OUT_CHROMSIZE] = _open_out(f"{outfile}.{n}.chromSize.txt")
        if bamio.find_bai(fn) is None:
            sys.stderr.write("bam2bed: BAM indexing file is not available.\n")
            raise SystemExit(1)
        gpu_inflate = os.environ.get("NGSCOV_GPU_INFLATE", "0") not in ("", "0")
        if gpu_inflate:
            # like the C drivers with NGSCOV_GPU_INFLATE=1: only the header blocks are inflated on the host
            import zlib
            with open(fn, "rb") as f:
                raw = f.read()
            inf, hdr = b"", None
            for off, bs in bamio.bgzf_blocks(raw):
                inf += zlib.decompress(raw[off + 18:off + bs - 8], -15)
                try:
                    hdr = bamio.parse_header(inf)
                    break
                except Exception:
                    continue
            if hdr is None:
                sys.stderr.write(f"bam2bed: {fn} has no BAM header\n")
                raise SystemExit(1)
            names, lens, hdr_len = hdr
        else:
            names, lens, records = bamio.load_bam(fn)
        eng = CoverageEngine(names, lens, mode=MODE_DEPTH if depth_mode else MODE_WIG, window=window,
                             emit_wig=bool(wig), device=int(os.environ.get("NGSCOV_DEVICE", "0")))
        try:
            if gpu_inflate:
                eng.push_bgzf(raw, skip=hdr_len, last=True)
            else:
                eng.push_host(records, last=True)
            eng.finish()
            for which, f in streams.items():
                f.write(eng.output(which))
        finally:
            eng.close()
        for nm in names:
            sys.stderr.write("%s at %.3f s\n" % (nm.decode(), time.time() - begin))
        for f in streams.values():
            f.close()
        sys.stderr.write("Converted %s to wig format at %.3f s\n" % (fn, time.time() - begin))
    return 0


def bam2depth(argv):
    """bam2depth [-o PFX] [-w WIN] [-W] [-r REG] [-s N] [-h] a.bam [b.bam ..]"""
    return _run(list(argv), True)


def bam2wig(argv):
    """bam2wig [-o PFX] [-w WIN] [-r REG] [-s N] [-h] a.bam [b.bam ..]"""
    return _run(list(argv), False)


if __name__ == "__main__":
    tool = os.path.basename(sys.argv[1]) if len(sys.argv) > 1 else ""
    if tool not in ("bam2depth", "bam2wig"):
        sys.stderr.write("usage: python -m highperformancengs_b200.cli bam2depth|bam2wig [args]\n")
        raise SystemExit(1)
    raise SystemExit((bam2depth if tool == "bam2depth" else bam2wig)(sys.argv[1:]))
