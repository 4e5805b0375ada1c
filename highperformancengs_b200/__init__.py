"""highperformancengs_b200 -- B200-native coverage engine behind the bam2depth / bam2wig tools.

Only the one hot path of xiongxu/HighPerformanceNGS is here: BAM record stream -> per-base depth ->
bedGraph / .depth / wig.  The compute lives in ``libngscov_cuda.so`` (hand-written sm_100a kernels
behind the C ABI of ``include/ngscov.h``); this package is the thin host-side mirror used by the
tests, the benchmark and the Python entry points.  There is no CPU fallback.
"""
from ._lib import load_library, NgscovError  # noqa: F401
from .engine import CoverageEngine, MODE_DEPTH, MODE_WIG  # noqa: F401
from .engine import OUT_BEDGRAPH, OUT_DEPTH, OUT_WIG, OUT_CHROMSIZE  # noqa: F401
from .cli import bam2depth, bam2wig  # noqa: F401
