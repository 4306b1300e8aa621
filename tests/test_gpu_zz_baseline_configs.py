"""BASELINE.json's other configurations as parity cases: C1 on the CPU (plumbing of the
oracle through the reference-style API) and a scaled-down C5 on the GPU — 8 leaves generated like the
bench index (`seed + leaf`), a batch that alternates 2–3-term MUST and 3–5-term SHOULD queries with
log-uniform term ranks, k = 100, `search` and `search_parallel` collection.  (This file sorts after the
other GPU tests on purpose: it only recombines features they already cover.)"""
import numpy as np
import pytest

import helpers
import oracle_binding as ob
from rucene_b200 import codec, search


def test_c1_term_query_on_10k_docs_cpu():
    """C1: N = 10 000, one TermQuery on rank 5, k = 10 — the oracle alone."""
    seg = codec.synth_segment(0x5EED0001, 10000, 2000, doc_version=1)
    ix = helpers.oracle_index([seg])
    q, c = ob.make_queries([("term", 4)])           # rank 5 -> term id 4, df ~ N/6
    hits, counts, total = ix.search_batch(q, c, 10)
    df = int(seg.terms["doc_freq"][4])
    assert total[0] == df and 1200 < df < 2200 and counts[0] == 10
    scores = hits[0]["score"][:10]
    assert np.all(np.diff(scores) <= 0) and scores[0] > 0
    docs, freqs = ix.postings(0, 4, df + 1)
    assert set(hits[0]["doc"][:10].tolist()) <= set(docs.tolist())


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [0, 1])
def test_c5_mixed_and_or_over_eight_leaves(mode):
    n_terms = 3000
    segs = [codec.synth_segment(0x5EED0005 + s, 40000, n_terms, doc_version=1) for s in range(8)]
    rng = np.random.default_rng(0x5EED0005)
    specs = []
    for i in range(96):
        if i % 2 == 0:
            ts = helpers.distinct_query_terms(rng, n_terms, 1, 2, 3)[0]
            specs.append(("bool", [(ob.MUST, t) for t in ts], 0))
        else:
            ts = helpers.distinct_query_terms(rng, n_terms, 1, 3, 5)[0]
            specs.append(("bool", [(ob.SHOULD, t) for t in ts], 0))
    ix = helpers.oracle_index(segs)
    q, c = ob.make_queries(specs)
    want = ix.search_batch(q, c, 100, parallel_mode=mode, n_threads=4)
    s = search.GpuIndexSearcher(search.IndexReader(segs))
    try:
        got = s.search_batch(helpers.to_queries(specs), 100, mode=mode)
    finally:
        s.engine.close()
    helpers.assert_same_topdocs(got, want, "C5 mode %d" % mode)
