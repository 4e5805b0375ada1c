"""CPU: the oracle (oracle/ngscov_oracle.c) against the committed outputs of the unmodified reference
binaries (tests/golden/reference_outputs.json, made by tests/golden/make_golden.py).  This is what pins
the oracle where /root/reference is absent (the GPU box)."""
import hashlib
import json
import os

import pytest

from tests import cases

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_outputs.json")))
_cache = {}


def _input(name, kind, params):
    if name not in _cache:
        _cache.clear()
        _cache[name] = cases.case_input(kind, params)
    return _cache[name]


def check_against_gold(got, key):
    gold = GOLD[key]
    for f, meta in gold.items():
        data = got[f]
        assert len(data) == meta["size"], f"{key}: {f} size {len(data)} != reference {meta['size']}"
        assert hashlib.md5(data).hexdigest() == meta["md5"], f"{key}: {f} differs from the reference output"
        if "text" in meta:
            assert data.decode() == meta["text"]


PARAMS = [(n, k, p, tool, w) for n, k, p, runs in cases.CASES for tool, w in runs]


@pytest.mark.parametrize("name,kind,params,tool,window", PARAMS, ids=[f"{p[0]}-{p[3]}-{p[4]}" for p in PARAMS])
def test_oracle_matches_reference_outputs(oracle, name, kind, params, tool, window):
    names, lens, recs, g = _input(name, kind, params)
    got = oracle.run(recs, names, lens, 0 if tool == "depth" else 1, window, emit_wig=True)
    check_against_gold(got, f"{name}|{tool}|{window}")


def test_gold_covers_every_case():
    assert set(GOLD) == {f"{n}|{t}|{w}" for n, k, p, runs in cases.CASES + cases.FULL_CASES for t, w in runs}


# ---- the inputs the reference itself holds: samtools-0.1.19 examples/toy.sam and examples/ex1.sam.gz
ST = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "samtools_examples.json")))
ST_PARAMS = [(tag, key) for tag in sorted(ST) for key in sorted(ST[tag]["expected"])]
SURVEY_MD5 = {("toy", "depth|10", "bedgraph"): "4f4f60ece97bd03e9c84e2960c26def2", ("ex1", "depth|100", "bedgraph"): "8a06af30086d31a290d43f2ac3ad2343",
              ("ex1", "depth|100", "depth"): "0ae8d46b075ea0bac0fd01e1a61b3226", ("ex1", "depth|100", "wig"): "84fce416017db6a96779e0a8f30b2007",
              ("ex1", "wig|100", "wig"): "1385dda8cc49d86f0e1f104f509ee33d"}


def samtools_example(tag):
    from tests import bamutil
    c = ST[tag]
    names = [n.encode() for n, _ in c["contigs"]]
    lens = [L for _, L in c["contigs"]]
    return names, lens, bamutil.records([tuple(r) for r in c["reads"]], seq=False)


def check_samtools_example(got, tag, key):
    for f, meta in ST[tag]["expected"][key].items():
        assert len(got[f]) == meta["size"], f"{tag} {key}: {f} size"
        assert hashlib.md5(got[f]).hexdigest() == meta["md5"], f"{tag} {key}: {f} differs from the reference output"
        if (tag, key, f) in SURVEY_MD5:
            assert meta["md5"] == SURVEY_MD5[(tag, key, f)]            # the value SURVEY.md section 8(c) recorded


@pytest.mark.parametrize("tag,key", ST_PARAMS)
def test_oracle_on_samtools_examples(oracle, tag, key):
    """toy.sam (12 reads, M/I/D/N/S/H/P) and ex1.sam.gz (3 307 reads): the oracle reproduces the outputs of the reference
    binaries on the files the reference ships, with the md5s SURVEY.md recorded"""
    names, lens, recs = samtools_example(tag)
    tool, w = key.split("|")
    got = oracle.run(recs, names, lens, 0 if tool == "depth" else 1, int(w), emit_wig=True)
    check_samtools_example(got, tag, key)
