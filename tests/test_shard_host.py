"""Host-side shard logic (no GPU): slice planning and halo stitching, checked by cutting the oracle's
whole-genome runs at the slice boundaries and stitching them back; plus a 2-rank gloo exchange."""
import os
import sys

import numpy as np
import pytest

from tests import bamutil
from highperformancengs_b200 import shard


def parse_bedgraph(text, names):
    idx = {n.decode(): i for i, n in enumerate(names)}
    runs = []
    for line in text.decode().splitlines():
        c, s, e, d = line.split("\t")
        runs.append((idx[c], int(s), int(e), int(d)))
    return runs


def cut_runs(runs, lens, slices):
    """what each shard would hold: runs clipped to its slice, in contig coordinates"""
    starts = np.concatenate([[0], np.cumsum(lens)])
    out = []
    for beg, end in slices:
        mine = []
        for c, s, e, d in runs:
            gs, ge = int(starts[c]) + s, int(starts[c]) + e
            lo, hi = max(gs, beg), min(ge, end)
            # runs are not clipped at the contig end, only at slice cuts inside the contig
            if ge > int(starts[c + 1]) and end >= int(starts[c + 1]) and beg < int(starts[c + 1]):
                hi = ge if lo < int(starts[c + 1]) else hi
            if lo < hi:
                mine.append((c, lo - int(starts[c]), hi - int(starts[c]), d))
        out.append(mine)
    return out


def halo_of(runs):
    if not runs:
        return {"n_runs": 0, "first": (-1, 0, 0, 0), "last": (-1, 0, 0, 0)}
    return {"n_runs": len(runs), "first": runs[0], "last": runs[-1]}


def apply(decisions, per_shard):
    out = []
    for (drop, new_end), runs in zip(decisions, per_shard):
        rs = list(runs)
        if new_end:
            c, s, e, d = rs[-1]
            rs[-1] = (c, s, new_end, d)
        if drop:
            rs = rs[1:]
        out += rs
    return out


@pytest.mark.parametrize("n", [2, 3, 4, 8])
@pytest.mark.parametrize("seed", [1, 5])
def test_stitch_reassembles_whole_genome(oracle, n, seed):
    names, lens, reads = bamutil.random_case(seed, n_reads=4000, contigs=((b"c1", 100000), (b"c2", 40000), (b"c3", 200000)))
    recs = bamutil.records(reads, seq=False)
    whole = parse_bedgraph(oracle.run(recs, names, lens, 0, 1000)["bedgraph"], names)
    slices = shard.plan_slices(lens, n)
    assert slices[0][0] == 0 and slices[-1][1] == sum(lens)
    assert all(slices[i][1] == slices[i + 1][0] for i in range(n - 1))
    per = cut_runs(whole, lens, slices)
    decisions, corr = shard.stitch([halo_of(r) for r in per])
    assert apply(decisions, per) == whole


def test_stitch_chain_through_single_run_shards():
    # one long run cut into four shards; middle shards hold exactly one run each
    per = [[(0, 5, 100, 3)], [(0, 100, 200, 3)], [(0, 200, 300, 3)], [(0, 300, 350, 3), (0, 350, 360, 1)]]
    decisions, corr = shard.stitch([halo_of(r) for r in per])
    assert apply(decisions, per) == [(0, 5, 350, 3), (0, 350, 360, 1)]
    assert decisions[0] == (False, 350) and decisions[1][0] and decisions[2][0] and decisions[3][0]


def test_no_merge_when_depth_differs_or_gap():
    per = [[(0, 5, 100, 3)], [(0, 100, 200, 4)], [], [(0, 300, 350, 4)]]
    decisions, corr = shard.stitch([halo_of(r) for r in per])
    assert all(d == (False, 0) for d in decisions) and corr == []


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    per = [[(0, 5, 100, 3), (0, 100, 16384, 2)], [(0, 16384, 20000, 2), (0, 20000, 20010, 7)]]
    mine = halo_of(per[rank])
    halos = [None] * world
    dist.all_gather_object(halos, mine)
    decisions, corr = shard.stitch(halos)
    q.put((rank, decisions[rank], len(corr)))
    dist.destroy_process_group()


def test_gloo_two_rank_halo_exchange():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    ps = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(60)
    assert res[0] == (0, (False, 20000), 3)
    assert res[1] == (1, (True, 0), 3)
