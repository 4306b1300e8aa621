"""Committed golden fixtures (tests/golden/make_golden.py): the oracle and the generator must
reproduce them on any host; the GPU path must reproduce them on the GPU box."""
import os
import sys
import zlib

import numpy as np
import pytest

import helpers
import oracle_binding as ob
from rucene_b200 import codec

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
sys.path.insert(0, GOLD)
import make_golden  # noqa: E402


def _load():
    g = np.load(os.path.join(GOLD, "topdocs_v1.npz"))
    seed, max_doc, n_terms, k = (int(x) for x in g["params"])
    seg = codec.synth_segment(seed, max_doc, n_terms, doc_version=1)
    return g, seg, k


def _as_triple(g):
    hits = np.zeros(g["docs"].shape, ob.HIT_DTYPE)
    hits["doc"] = g["docs"]
    hits["score"] = g["scores"].view(np.float32)
    return hits, g["counts"], g["total"]


def test_generator_and_oracle_reproduce_golden():
    g, seg, k = _load()
    assert zlib.crc32(seg.doc_file.tobytes()) == int(g["doc_file_crc"])
    assert zlib.crc32(seg.norms.tobytes()) == int(g["norms_crc"])
    assert np.array_equal(make_golden.flat_specs(), g["specs"])
    q, c = ob.make_queries(make_golden.specs())
    got = helpers.oracle_index([seg]).search_batch(q, c, k)
    helpers.assert_same_topdocs(got, _as_triple(g), "oracle vs golden")


@pytest.mark.gpu
def test_gpu_reproduces_golden():
    from rucene_b200 import search
    g, seg, k = _load()
    s = search.GpuIndexSearcher(search.IndexReader([seg]))
    try:
        got = s.search_batch(helpers.to_queries(make_golden.specs()), k)
    finally:
        s.engine.close()
    helpers.assert_same_topdocs(got, _as_triple(g), "gpu vs golden")
