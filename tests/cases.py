"""The parity cases shared by the golden-fixture generator and the tests (inputs are regenerated
deterministically: hand-written/random read lists from tests/bamutil.py, synthetic configs from
tools/bamgen.c).  TEST INFRASTRUCTURE."""
import numpy as np

from tests import bamutil

# (name, kind, params, [(tool, window), ..])
CASES = [
    ("toy", "reads", "toy", [("depth", 10), ("depth", 1), ("wig", 10), ("wig", 7)]),
    ("quirk", "reads", "quirk", [("depth", 25), ("wig", 25), ("wig", 1), ("depth", 20000)]),
    ("random1", "reads", 1, [("depth", 150), ("wig", 150), ("depth", 999), ("wig", 1000)]),
    ("random2", "reads", 2, [("depth", 1000), ("wig", 999), ("wig", 20000)]),
    ("random3", "reads", 3, [("depth", 7), ("wig", 50000)]),
    ("overhang", "reads", "overhang", [("depth", 300), ("wig", 300)]),
    ("c1_s0.05", "gen", (1, 0.05), [("depth", 20000), ("wig", 20000)]),
    ("c2_s0.0005", "gen", (2, 0.0005), [("depth", 20000), ("depth", 500), ("wig", 20000)]),
    ("c4_s0.01", "gen", (4, 0.01), [("wig", 20000), ("wig", 100), ("depth", 100)]),
    ("c5_s0.002", "gen", (5, 0.002), [("depth", 20000), ("wig", 1000)]),
    ("c1_full", "gen", (1, 1.0), [("depth", 20000)]),
]

# full-size configurations 4 and 5 (BASELINE.json): 20 M spliced reads / 50 M reads piled on 200 kb.  The reference binary
# ran them here once (tests/golden/make_golden.py); the CPU oracle needs minutes on them, so only the GPU tests (md5 of
# the engine's output against the reference's) and the generator consume these.
FULL_CASES = [
    ("c4_full", "gen", (4, 1.0), [("wig", 20000), ("wig", 100)]),
    ("c5_full", "gen", (5, 1.0), [("depth", 20000)]),
]


def case_input(kind, params):
    """-> (names, lens, records uint8 array, generator or None)"""
    if kind == "reads":
        if params == "toy":
            n, l, r = bamutil.toy_case()
        elif params == "quirk":
            n, l, r = bamutil.quirk_case()
        elif params == "overhang":
            n, l, r = [b"c"], [1000], [(0, 100, 0, "50M"), (0, 950, 0, "80M"), (0, 960, 16, "60M")]
        else:
            n, l, r = bamutil.random_case(params)
        return n, l, bamutil.records(r, seq=(params != 2)), None
    from highperformancengs_b200.gen import Generator
    g = Generator(params[0], scale=params[1])
    return g.names, g.lens, g.records(), g


def files_of(tool, emit_wig=True):
    return ("bedgraph", "depth", "wig", "chromsize") if tool == "depth" else ("wig", "chromsize")
