"""Host-side shard logic (no GPU): slice planning and halo stitching, checked by cutting the oracle's
whole-genome runs at the slice boundaries and stitching them back; plus a 2-rank gloo exchange."""
import os
import subprocess
import sys

import numpy as np
import pytest

from tests import bamutil
from highperformancengs_b200 import shard

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


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


# ---- the C drivers' copy of this logic (host/ngscov_shard.h), compiled without a GPU -------------------------------
@pytest.fixture(scope="module")
def shard_check():
    out = os.path.join(ROOT, "tests", "_build")
    os.makedirs(out, exist_ok=True)
    exe = os.path.join(out, "shard_check")
    subprocess.check_call(["gcc", "-O2", "-Wall", os.path.join(ROOT, "tests", "shard_check.c"), "-o", exe])
    return exe


def test_c_plan_slices_equals_python(shard_check):
    rng = np.random.RandomState(5)
    cases = [([248956422, 242193529, 198295559, 190214555, 181538259, 170805979, 159345973, 145138636, 138394717, 133797422,
               135086622, 133275309, 114364328, 107043718, 101991189, 90338345, 83257441, 80373285, 58617616, 64444167,
               46709983, 50818468, 156040895, 57227415], n) for n in (1, 2, 4, 8, 64)]
    for _ in range(40):
        lens = [int(x) for x in rng.randint(1, 400000, size=rng.randint(1, 12))]
        cases.append((lens, int(rng.randint(1, 20))))
    cases.append(([10, 20, 30], 7))                      # more shards than 16 kb tiles: empty slices
    for lens, n in cases:
        out = subprocess.run([shard_check, "plan", str(n)] + [str(x) for x in lens], capture_output=True, text=True, check=True).stdout.split()
        got = [(int(out[4 * i]), int(out[4 * i + 1])) for i in range(n)]
        assert got == shard.plan_slices(lens, n), (lens, n)
        for i, (b, _) in enumerate(got):
            assert (int(out[4 * i + 2]), int(out[4 * i + 3])) == tuple(shard.locate(lens, b))


def test_c_stitch_equals_python(shard_check):
    rng = np.random.RandomState(9)
    for _ in range(300):
        n = int(rng.randint(1, 9))
        halos, args = [], []
        pos = 0
        for r in range(n):
            k = int(rng.choice([0, 1, 1, 2, 5]))
            c = int(rng.randint(0, 2))
            if k == 0:
                h = {"n_runs": 0, "first": (-1, 0, 0, 0), "last": (-1, 0, 0, 0)}
            else:
                # runs that often abut the previous shard's last run with the same depth (the merge case)
                s0 = pos if rng.rand() < 0.7 else pos + int(rng.randint(1, 50))
                d0 = int(rng.randint(1, 3))
                f = (c, s0, s0 + int(rng.randint(1, 100)), d0)
                l = f if k == 1 else (c, f[2] + int(rng.randint(0, 50)), f[2] + 60 + int(rng.randint(1, 100)), int(rng.randint(1, 3)))
                pos = l[2]
                h = {"n_runs": k, "first": f, "last": l}
            halos.append(h)
            args += [str(h["n_runs"])] + [str(x) for x in h["first"]] + [str(x) for x in h["last"]]
        out = subprocess.run([shard_check, "stitch", str(n)] + args, capture_output=True, text=True, check=True).stdout.strip().split("\n")
        dec, corr = shard.stitch(halos)
        got_dec = [(bool(int(l.split()[0])), int(l.split()[1])) for l in out[:n]]
        assert got_dec == [(bool(a), int(b)) for a, b in dec]
        got_corr = [tuple(int(x) for x in l.split()[1:]) for l in out[n:]]
        assert got_corr == [tuple(int(x) for x in run) + (sign,) for run, sign in corr]


def test_c_bai_seek_matches_python_reader(shard_check, tmp_path):
    """virtual offset chosen for a slice start = first non-zero linear-index entry at or after its 16 kb tile, else the
    first chunk of the next contig that has reads (samtools bam_index.c:92-111,605-615 semantics)"""
    from highperformancengs_b200 import bamio
    from highperformancengs_b200.gen import Generator
    bam = str(tmp_path / "x.bam")
    g = Generator(4, scale=0.002)
    g.write_bam(bam)
    refs, _ = bamio.read_bai(bam + ".bai")
    for c in range(len(refs)):
        for p in (0, 16384, 5 * 16384, max(0, g.lens[c] - 1)):
            exp = 0
            lin = refs[c]["linear"]
            for w in range(p >> 14, len(lin)):
                if lin[w]:
                    exp = lin[w]
                    break
            if not exp:
                for c2 in range(c + 1, len(refs)):
                    beg = [ch[0] for b, chs in refs[c2]["bins"].items() if b != 37450 for ch in chs]
                    if beg:
                        exp = min(beg)
                        break
            out = subprocess.run([shard_check, "seek", bam, str(c), str(p)], capture_output=True, text=True, check=True).stdout.split()
            assert int(out[0]) == len(refs) and int(out[1]) == exp, (c, p)
