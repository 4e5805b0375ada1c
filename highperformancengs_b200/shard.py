"""Region sharding across GPUs (no counterpart in the single-process reference).

The genome is cut into contiguous slices of concatenated-genome coordinates at multiples of 16 384 bp
inside contigs (the BAI linear-index granularity, samtools bam_index.c:43).  Each GPU accumulates every
read overlapping its slice with deltas clipped to the slice, so depth inside a slice is exact; the only
exchange is the halo: the last run of slice g and the first run of slice g+1 merge into one bedGraph
line when they abut at the cut with equal depth.  Window bins are corrected for merged runs (the
bam2wig window arithmetic is not additive over pieces of a run) and summed over ranks.
"""
import numpy as np

TILE_BP = 16384


def plan_slices(ref_lens, n):
    """-> list of n (beg, end) in concatenated-genome coordinates covering [0, G)."""
    lens = [int(x) for x in ref_lens]
    G = sum(lens)
    starts = np.concatenate([[0], np.cumsum(lens)])
    cuts = [0]
    for r in range(1, n):
        g = G * r // n
        c = int(np.searchsorted(starts, g, side="right") - 1)
        c = min(c, len(lens) - 1)
        pos = (g - int(starts[c])) // TILE_BP * TILE_BP
        cut = int(starts[c]) + pos
        cuts.append(max(cut, cuts[-1]))
    cuts.append(G)
    return [(cuts[i], cuts[i + 1]) for i in range(n)]


def locate(ref_lens, g):
    """concatenated coordinate -> (contig, position); g == G maps to (n_ref, 0)."""
    acc = 0
    for c, L in enumerate(ref_lens):
        if g < acc + L:
            return c, g - acc
        acc += L
    return len(ref_lens), 0


def stitch(halos):
    """halos: per rank dict(n_runs, first=(contig,start,end,depth), last=...) in rank order.

    Returns (decisions, corrections): decisions[r] = (drop_first, new_last_end) for ngscov_halo_put;
    corrections = list of (run, sign) to apply to the summed window bins with ngscov_bins_apply_run.
    """
    n = len(halos)
    decisions = [[False, 0] for _ in range(n)]
    corrections = []
    carry = None          # [contig, start, end, depth, owner_rank]: the run still open at the current cut
    for r in range(n):
        h = halos[r]
        if h["n_runs"] == 0:
            carry = None
            continue
        f = h["first"]
        merged = False
        if carry is not None and carry[0] == f[0] and carry[2] == f[1] and carry[3] == f[3]:
            merged = True
            corrections.append(((carry[0], carry[1], carry[2], carry[3]), -1))
            corrections.append((tuple(f), -1))
            carry[2] = f[2]
            corrections.append(((carry[0], carry[1], carry[2], carry[3]), +1))
            decisions[r][0] = True
            decisions[carry[4]][1] = carry[2]
        if merged and h["n_runs"] == 1:
            continue          # the whole shard was one run: it stays open into the next shard
        l = h["last"]
        carry = [l[0], l[1], l[2], l[3], r]
    return [tuple(d) for d in decisions], corrections
