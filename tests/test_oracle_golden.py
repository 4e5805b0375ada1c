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
    assert set(GOLD) == {f"{n}|{t}|{w}" for n, k, p, runs in cases.CASES for t, w in runs}
