"""Parity cases of the bam_sliding_count path (NGSCOV_MODE_SLIDE), shared by the golden generator and the tests.
TEST INFRASTRUCTURE."""
from tests import bamutil

# (name, kind, params, [window, ..])
CASES = [
    ("random1", "reads", 1, [150, 20000, 7]),
    ("random3", "reads", 3, [1000, 33]),
    ("c1_s0.05", "gen", (1, 0.05), [20000, 777, 200000]),
    ("c2_s0.0005", "gen", (2, 0.0005), [20000, 1000000]),
    ("c4_s0.01", "gen", (4, 0.01), [20000]),
    # every read of a contig in ONE window: the per-window GC sum passes 2^24, where the reference's float accumulation
    # depends on the order of the reads
    ("c5_s0.02_onewindow", "gen", (5, 0.02), [200000000, 20000]),
]

# the `-r` mode: (case name, window, region string, (refID, beg, end) as bam_parse_region yields them)
REGION_CASES = [
    ("random1", 150, "c1:1000-5000", (0, 999, 5000)),
    ("random1", 20000, "c3", (2, 0, 1 << 29)),
    ("random3", 33, "c1:2,001-2,500", (0, 2000, 2500)),
    ("random3", 1000, "c3:69000", (2, 68999, 1 << 29)),
    ("c1_s0.05", 777, "chr1:1-20000", (0, 0, 20000)),
    ("c2_s0.0005", 20000, "chr2:30000-90000", (1, 29999, 90000)),
]


def case_input(kind, params):
    """-> (names, lens, records uint8 array, generator or None), records WITH SEQ"""
    if kind == "reads":
        n, l, r = bamutil.random_case(params)
        return n, l, bamutil.records(r, seq=True, seq_seed=params), None
    from highperformancengs_b200.gen import Generator
    g = Generator(params[0], scale=params[1], seq=True)
    return g.names, g.lens, g.records(), g
