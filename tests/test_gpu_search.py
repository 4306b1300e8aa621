"""GPU parity: IndexSearcher::search through the C ABI vs the oracle — TopDocs identical
(docids bit-exact, BM25 scores bit-exact f32 — tighter than the 1e-5 relative the north star
allows —, same order under ties, same total_hits)."""
import os

import numpy as np
import pytest

import helpers
import oracle_binding as ob
from rucene_b200 import codec, engine, search

pytestmark = pytest.mark.gpu

DFS = [0, 1, 2, 100, 127, 128, 129, 255, 256, 257, 1000, 5000, 20000, 40000, 59000]


def _mixed_specs(rng, n_terms, n, kinds=("term", "and", "or")):
    specs = []
    for i in range(n):
        kind = kinds[i % len(kinds)]
        if kind == "term":
            specs.append(("term", int(rng.integers(0, n_terms))))
        else:
            t = int(rng.integers(2, 6))
            terms = rng.choice(n_terms, size=t, replace=False)
            occ = ob.MUST if kind == "and" else ob.SHOULD
            specs.append(("bool", [(occ, int(x)) for x in terms], 0))
    return specs


def _run_both(segs, specs, k, mode=0, range_postings=0, threads=4, extra_flags=0):
    ix = helpers.oracle_index(segs)
    q, c = ob.make_queries(specs)
    want = ix.search_batch(q, c, k, parallel_mode=mode, n_threads=threads)
    # every evaluation route must equal the oracle: a score column for every clause of df >= max_doc/64
    # (k_eval_or_ms: presence bitmaps + bit-sliced per-document bound), the same with tf-norm planes, the same
    # columns read by the exhaustive kernel (every other disjunction clause of df >= 256 streamed from its scored
    # list), the same without lists, no bitmaps / columns / lists at all (block streams only), the planner's choice
    got = None
    for flags in (engine.CFG_EAGER_COLUMNS | engine.CFG_MAXSCORE, engine.CFG_EAGER_COLUMNS | engine.CFG_MAXSCORE | engine.CFG_TFPLANES,
                  engine.CFG_EAGER_COLUMNS, engine.CFG_EAGER_COLUMNS | engine.CFG_NO_LISTS,
                  engine.CFG_NO_BITMAPS | engine.CFG_NO_LISTS, engine.CFG_MAXSCORE, 0):
        s = search.GpuIndexSearcher(search.IndexReader(segs), range_postings=range_postings,
                                    flags=flags | extra_flags)
        try:
            got = s.search_batch(helpers.to_queries(specs), k, mode=mode)
        finally:
            s.engine.close()
        if flags:
            helpers.assert_same_topdocs(got, want, "engine flags %d" % flags)
    return got, want


@pytest.mark.parametrize("version", [1, 0])
@pytest.mark.parametrize("k", [1, 10, 100])
def test_small_index_every_shape(version, k):
    rng = np.random.default_rng(1000 + version)
    seg, _ = helpers.build_segment(rng, 60000, DFS, doc_version=version, dense_terms=(11,))
    specs = [("term", t) for t in range(len(DFS))]
    specs += _mixed_specs(rng, len(DFS), 90, kinds=("and", "or"))
    specs += [("bool", [(ob.SHOULD, 12, 2.0), (ob.SHOULD, 3, 0.5)], 0),
              ("bool", [(ob.MUST, 14), (ob.MUST, 13), (ob.MUST, 12), (ob.MUST, 11)], 0),
              ("bool", [(ob.MUST, 14), (ob.MUST, 0)], 0),
              ("bool", [(ob.SHOULD, 0), (ob.SHOULD, 1)], 0),
              ("bool", [(ob.SHOULD, 7)], 0), ("bool", [(ob.MUST, 9)], 0)]
    got, want = _run_both([seg], specs, k)
    helpers.assert_same_topdocs(got, want, "v%d k%d" % (version, k))


def test_ranges_split_queries_across_ctas():
    """Tiny range_postings forces many (query, docid-range) work items per query: candidate
    lists, theta chaining and the replay must still reproduce the sequential collector."""
    rng = np.random.default_rng(7)
    seg, _ = helpers.build_segment(rng, 200000, [150000, 90000, 30000, 5000, 300, 1])
    specs = _mixed_specs(rng, 6, 60)
    for rp in (1 << 18, 4096, 700):
        got, want = _run_both([seg], specs, 10, range_postings=rp)
        helpers.assert_same_topdocs(got, want, "range_postings=%d" % rp)


def test_score_ties_follow_heap_layout():
    """All postings share freq and norm => massive score ties; which docs survive and in what
    order is decided by the BinaryHeap layout, not by docid."""
    max_doc = 50000
    w = codec.PostingsWriter(doc_version=1, max_doc=max_doc)
    rng = np.random.default_rng(3)
    for df in (30000, 12000, 700):
        docs = np.sort(rng.choice(max_doc, df, replace=False)).astype(np.int32)
        w.add_term(docs, np.ones(df, np.int32))
    seg = w.finish(norms=np.full(max_doc, 120, np.uint8))
    specs = [("term", 0), ("term", 1), ("term", 2),
             ("bool", [(ob.SHOULD, 0), (ob.SHOULD, 1)], 0),
             ("bool", [(ob.MUST, 0), (ob.MUST, 1)], 0),
             ("bool", [(ob.SHOULD, 0), (ob.SHOULD, 1), (ob.SHOULD, 2)], 0)]
    for k in (3, 10, 100):
        for rp in (0, 2000):
            got, want = _run_both([seg], specs, k, range_postings=rp)
            helpers.assert_same_topdocs(got, want, "ties k=%d rp=%d" % (k, rp))


@pytest.mark.parametrize("mode", [0, 1])
def test_multi_segment_and_live_docs(mode):
    """Several leaves on one GPU: sequential-leaf collector (mode 0) and search_parallel's
    per-leaf heaps merged in leaf order (mode 1); statistics from the largest leaf only."""
    rng = np.random.default_rng(11 + mode)
    dfs = [0, 1, 90, 128, 400, 3000, 9000, 20000]
    segs = []
    for s, md in enumerate((30000, 52000, 41000)):
        d = list(dfs)
        if s == 1:
            d[4] = 0      # term absent from the largest leaf (the one that supplies doc_freq)
        if s == 2:
            d[6] = 0
        seg, _ = helpers.build_segment(rng, md, d, live_fraction=0.8 if s != 0 else None)
        segs.append(seg)
    specs = [("term", t) for t in range(len(dfs))] + _mixed_specs(rng, len(dfs), 60, kinds=("and", "or"))
    got, want = _run_both(segs, specs, 10, mode=mode, range_postings=3000)
    helpers.assert_same_topdocs(got, want, "mode %d" % mode)


def test_synthetic_zipf_index_c3_c4_shapes():
    """Scaled-down BASELINE configs 3 and 4: Zipfian synthetic segment, log-uniform query terms."""
    seg = codec.synth_segment(0x5EED0001, 300000, 5000, doc_version=1)
    rng = np.random.default_rng(0x5EED0003)
    and_q = helpers.distinct_query_terms(rng, 5000, 48, 2, 2)
    or_q = helpers.distinct_query_terms(rng, 5000, 48, 5, 5)
    specs = [("bool", [(ob.MUST, t) for t in ts], 0) for ts in and_q]
    specs += [("bool", [(ob.SHOULD, t) for t in ts], 0) for ts in or_q]
    got, want = _run_both([seg], specs, 10)
    helpers.assert_same_topdocs(got, want, "k=10")
    got, want = _run_both([seg], specs, 100, range_postings=20000)
    helpers.assert_same_topdocs(got, want, "k=100")


def test_large_k_and_many_leaves():
    """k = 1000 (heap/theta capacity 1024) over five leaves, both collector modes."""
    rng = np.random.default_rng(31)
    dfs = [3, 200, 1500, 6000, 15000, 30000]
    segs = [helpers.build_segment(rng, 40000 + 1000 * i, dfs)[0] for i in range(5)]
    specs = [("term", 5), ("term", 2)] + _mixed_specs(rng, len(dfs), 22, kinds=("and", "or"))
    for mode in (0, 1):
        got, want = _run_both(segs, specs, 1000, mode=mode, range_postings=5000)
        helpers.assert_same_topdocs(got, want, "k=1000 mode %d" % mode)


def test_must_not_clauses_req_not_scorer():
    """ReqNotScorer (search/scorer/req_not_scorer.rs; wiring boolean_query.rs:253-278): MUST_NOT with
    pure-MUST and pure-SHOULD queries, several MUST_NOT clauses, absent clauses, tails, ranges."""
    rng = np.random.default_rng(21)
    dfs = [0, 1, 60, 128, 300, 2500, 9000, 26000, 52000]
    segs = [helpers.build_segment(rng, 70000, dfs, live_fraction=lf)[0] for lf in (None, 0.85)]
    specs = []
    for i in range(80):
        n_pos = int(rng.integers(1, 4))
        n_neg = int(rng.integers(1, 3))
        terms = [int(x) for x in rng.choice(len(dfs), size=n_pos + n_neg, replace=False)]
        occ = ob.MUST if i % 2 else ob.SHOULD
        specs.append(("bool", [(occ, t) for t in terms[:n_pos]] + [(ob.MUST_NOT, t) for t in terms[n_pos:]], 0))
    specs += [("bool", [(ob.SHOULD, 8), (ob.MUST_NOT, 7), (ob.MUST_NOT, 6), (ob.MUST_NOT, 0)], 0),
              ("bool", [(ob.MUST, 7), (ob.MUST, 8), (ob.MUST_NOT, 6), (ob.MUST_NOT, 1)], 0),
              ("bool", [(ob.MUST, 0), (ob.MUST_NOT, 5)], 0),
              ("bool", [(ob.SHOULD, 0), (ob.MUST_NOT, 5)], 0)]
    for k, rp in ((10, 0), (100, 1500)):
        for mode in (0, 1):
            got, want = _run_both(segs, specs, k, mode=mode, range_postings=rp)
            helpers.assert_same_topdocs(got, want, "must_not k=%d rp=%d mode=%d" % (k, rp, mode))


def test_req_opt_scorer_must_plus_should():
    """MUST + SHOULD in one query -> ReqOptScorer (search/scorer/req_opt_scorer.rs:19-65; wiring
    boolean_query.rs:253-262), optionally under a ReqNotScorer: the running-mean skip after 100
    scored docs is sequential state per (query, leaf).  Leaves where no SHOULD term exists fall back
    to the plain conjunction; leaves where a MUST term is missing produce nothing."""
    rng = np.random.default_rng(77)
    dfs = [0, 1, 2, 90, 129, 700, 5000, 14000, 26000, 33000]
    segs = []
    for s, lf in enumerate((None, 0.8, None)):
        d = list(dfs)
        if s == 2:
            d[5] = 0       # a SHOULD/MUST term missing from the last leaf
            d[8] = 0
        segs.append(helpers.build_segment(rng, 36000 + 700 * s, d, live_fraction=lf)[0])
    specs = [("bool", [(ob.MUST, 9), (ob.SHOULD, 8)], 0),
             ("bool", [(ob.MUST, 8), (ob.MUST, 9), (ob.SHOULD, 7), (ob.SHOULD, 5)], 0),
             ("bool", [(ob.SHOULD, 6), (ob.MUST, 7), (ob.SHOULD, 9), (ob.SHOULD, 2)], 0),
             ("bool", [(ob.MUST, 7), (ob.SHOULD, 0)], 0),
             ("bool", [(ob.MUST, 0), (ob.SHOULD, 7)], 0),
             ("bool", [(ob.MUST, 5), (ob.SHOULD, 9, 3.0), (ob.SHOULD, 1)], 1),
             ("bool", [(ob.MUST, 9), (ob.SHOULD, 8), (ob.MUST_NOT, 7)], 0),
             ("bool", [(ob.MUST, 9), (ob.MUST, 7), (ob.SHOULD, 6), (ob.SHOULD, 4), (ob.MUST_NOT, 5), (ob.MUST_NOT, 3)], 0),
             ("bool", [(ob.MUST, 4), (ob.SHOULD, 9)], 0),       # lead list with a vint tail only + 1 block
             ("bool", [(ob.MUST, 2), (ob.SHOULD, 3)], 0)]
    for i in range(40):
        n_must, n_should = int(rng.integers(1, 3)), int(rng.integers(1, 4))
        n_not = int(rng.integers(0, 2))
        terms = [int(x) for x in rng.choice(len(dfs), size=n_must + n_should + n_not, replace=False)]
        cl = [(ob.MUST, t) for t in terms[:n_must]] + [(ob.SHOULD, t) for t in terms[n_must:n_must + n_should]]
        cl += [(ob.MUST_NOT, t) for t in terms[n_must + n_should:]]
        specs.append(("bool", [cl[j] for j in rng.permutation(len(cl))], 0))
    specs += _mixed_specs(rng, len(dfs), 10, kinds=("and", "or"))   # same batch as the other kernels
    for k, rp in ((10, 0), (100, 2000)):
        for mode in (0, 1):
            got, want = _run_both(segs, specs, k, mode=mode, range_postings=rp)
            helpers.assert_same_topdocs(got, want, "req_opt k=%d rp=%d mode=%d" % (k, rp, mode))


def test_min_should_match_greater_than_one():
    """min_should_match > 1 (disjunction_scorer.rs:317-329): a doc needs that many SHOULD clauses of
    the leaf.  Only the top-level SHOULD side is filtered; beside a MUST the SHOULD disjunction sits
    behind ReqOptScorer::score -> advance(), which ignores it (:350-363) — the oracle restates both."""
    rng = np.random.default_rng(91)
    dfs = [0, 1, 130, 900, 5000, 14000, 26000, 33000, 36000]
    segs = []
    for s, lf in enumerate((None, 0.85)):
        d = list(dfs)
        if s == 1:
            d[5] = 0
        segs.append(helpers.build_segment(rng, 38000 + 900 * s, d, live_fraction=lf)[0])
    specs = [("bool", [(ob.SHOULD, 8), (ob.SHOULD, 7)], 2),
             ("bool", [(ob.SHOULD, 8), (ob.SHOULD, 7), (ob.SHOULD, 6), (ob.SHOULD, 5), (ob.SHOULD, 4)], 3),
             ("bool", [(ob.SHOULD, 8), (ob.SHOULD, 7), (ob.SHOULD, 6), (ob.SHOULD, 5), (ob.SHOULD, 4)], 5),
             ("bool", [(ob.SHOULD, 4), (ob.SHOULD, 5), (ob.SHOULD, 0)], 2),
             ("bool", [(ob.SHOULD, 3), (ob.SHOULD, 2)], 3),                       # msm > clauses: nothing
             ("bool", [(ob.SHOULD, 8), (ob.MUST_NOT, 7)], 2),                     # one SHOULD, msm 2: nothing
             ("bool", [(ob.SHOULD, 8), (ob.SHOULD, 6), (ob.SHOULD, 4), (ob.MUST_NOT, 7), (ob.MUST_NOT, 5)], 2),
             ("bool", [(ob.MUST, 7), (ob.SHOULD, 8), (ob.SHOULD, 6)], 2),         # beside a MUST: ignored
             ("bool", [(ob.MUST, 7), (ob.MUST, 8)], 2),
             ("bool", [(ob.SHOULD, 8)], 4)]                                        # collapses to the clause
    for i in range(30):
        t = int(rng.integers(2, 6))
        terms = [int(x) for x in rng.choice(len(dfs), size=t, replace=False)]
        specs.append(("bool", [(ob.SHOULD, x) for x in terms], int(rng.integers(2, t + 1))))
    specs += _mixed_specs(rng, len(dfs), 8, kinds=("and", "or"))
    for k, rp in ((10, 0), (100, 2500)):
        for mode in (0, 1):
            got, want = _run_both(segs, specs, k, mode=mode, range_postings=rp)
            helpers.assert_same_topdocs(got, want, "msm k=%d rp=%d mode=%d" % (k, rp, mode))


@pytest.mark.parametrize("version,with_pf", [(1, True), (1, False), (0, False)])
def test_ef_and_bitset_doc_blocks(version, with_pf):
    """EncodeType::EF / EncodeType::BITSET doc blocks (codec/postings/for_util.rs:337-372,417-468; read
    side posting_reader.rs:501-561,612-647,733-779): written where the reference's (dormant) writer rule
    picks them, decoded on the device by rank/select over the block's bitmaps.  Every query shape, both
    collector modes, docid ranges, live docs, two leaves."""
    rng = np.random.default_rng(123 + version)
    dfs = [0, 1, 127, 128, 129, 300, 1000, 5000, 20000, 33000, 47000]
    segs, n_other = [], 0
    for s, lf in enumerate((None, 0.8)):
        cnt = []
        seg, _ = helpers.build_segment(rng, 52000 + 300 * s, dfs, doc_version=version, live_fraction=lf,
                                       dense_terms=(7,), use_ef=True, with_pf=with_pf, counts=cnt)
        segs.append(seg)
        n_other += cnt[0][1] + cnt[0][2]
        assert cnt[0][1] > 0 and cnt[0][2] > 0      # both encodings occur
    specs = [("term", t) for t in range(len(dfs))]
    specs += _mixed_specs(rng, len(dfs), 60, kinds=("and", "or"))
    specs += [("bool", [(ob.MUST, 10), (ob.SHOULD, 9), (ob.SHOULD, 6)], 0),
              ("bool", [(ob.MUST, 9), (ob.MUST, 10), (ob.MUST_NOT, 8)], 0),
              ("bool", [(ob.SHOULD, 10), (ob.SHOULD, 8), (ob.SHOULD, 5), (ob.MUST_NOT, 9)], 0),
              ("bool", [(ob.SHOULD, 10), (ob.SHOULD, 9), (ob.SHOULD, 8)], 2),
              ("bool", [(ob.MUST, 4), (ob.MUST, 10)], 0), ("bool", [(ob.MUST, 5), (ob.MUST, 6), (ob.MUST, 9)], 0)]
    for k, rp in ((10, 0), (100, 3000)):
        for mode in (0, 1):
            got, want = _run_both(segs, specs, k, mode=mode, range_postings=rp)
            helpers.assert_same_topdocs(got, want, "ef v%d pf%d k=%d rp=%d mode=%d" % (version, with_pf, k, rp, mode))


def test_disjunction_max_query():
    """DisjunctionMaxQuery over TermQuerys (search/query/disjunction_max_query.rs:51-155;
    DisjunctionMaxScorer, search/scorer/disjunction_scorer.rs:106-186,241-263): the union of the disjuncts,
    score = max + (sum - max) * tie_breaker_multiplier.  One disjunct (in the query, or present in a leaf)
    is that TermScorer.  Mixed into a batch with the other shapes; both collector modes; ranges; live docs."""
    rng = np.random.default_rng(97)
    dfs = [0, 1, 130, 900, 5000, 14000, 26000, 33000, 36000]
    segs = []
    for s, lf in enumerate((None, 0.85)):
        d = list(dfs)
        if s == 1:
            d[5] = 0
        segs.append(helpers.build_segment(rng, 38000 + 900 * s, d, live_fraction=lf)[0])
    specs = [("dismax", [(8,), (7,)], 0.0),
             ("dismax", [(8,), (7,), (6, 2.0), (5,), (4,)], 0.3),
             ("dismax", [(4,), (5,), (0,)], 1.0),
             ("dismax", [(3,)], 0.5),                                 # one disjunct: the TermQuery itself
             ("dismax", [(5,), (2,)], 0.1),                           # one of them missing from the second leaf
             ("dismax", [(8,), (7,), (6,), (5,), (4,), (3,), (2,), (1,), (0,)], 0.25),
             ("dismax", [(6,), (6,)], 0.5)]
    for i in range(24):
        t = int(rng.integers(2, 6))
        terms = [int(x) for x in rng.choice(len(dfs), size=t, replace=False)]
        specs.append(("dismax", [(x, float(rng.choice([1.0, 0.5, 3.0]))) for x in terms],
                      float(rng.choice([0.0, 0.1, 0.5, 1.0]))))
    specs += _mixed_specs(rng, len(dfs), 12, kinds=("term", "and", "or"))
    specs += [("bool", [(ob.SHOULD, 8), (ob.SHOULD, 7), (ob.SHOULD, 6)], 2),
              ("bool", [(ob.SHOULD, 8), (ob.SHOULD, 6), (ob.MUST_NOT, 7)], 0)]
    for k, rp in ((10, 0), (100, 2500)):
        for mode in (0, 1):
            got, want = _run_both(segs, specs, k, mode=mode, range_postings=rp)
            helpers.assert_same_topdocs(got, want, "dismax k=%d rp=%d mode=%d" % (k, rp, mode))


def test_filter_clauses_and_pure_must_not():
    """FILTER clauses (required, score 0f32: boolean_query.rs:108-110, searcher.rs:158-197), a lone FILTER
    (ConstantScoreQuery boost 0) and pure MUST_NOT queries (MatchAllDocsQuery, boolean_query.rs:76-79), mixed into a
    batch with the other shapes; leaves with live docs; both collector modes."""
    rng = np.random.default_rng(131)
    dfs = [0, 1, 60, 129, 900, 4000, 9000, 21000, 30000]
    segs = [helpers.build_segment(rng, 34000 + 500 * s, dfs, live_fraction=lf)[0] for s, lf in enumerate((None, 0.85))]
    specs = [("bool", [(ob.FILTER, 7)], 0),
             ("bool", [(ob.MUST, 8), (ob.FILTER, 7)], 0),
             ("bool", [(ob.FILTER, 8), (ob.MUST, 5), (ob.MUST, 7)], 0),
             ("bool", [(ob.FILTER, 8), (ob.FILTER, 7)], 0),
             ("bool", [(ob.FILTER, 8), (ob.SHOULD, 7), (ob.SHOULD, 4)], 0),
             ("bool", [(ob.MUST, 6), (ob.FILTER, 8), (ob.SHOULD, 7), (ob.MUST_NOT, 5)], 0),
             ("bool", [(ob.FILTER, 8), (ob.MUST_NOT, 7)], 0),
             ("bool", [(ob.MUST, 8), (ob.FILTER, 0)], 0),
             ("bool", [(ob.MUST_NOT, 8)], 0),
             ("bool", [(ob.MUST_NOT, 7), (ob.MUST_NOT, 4), (ob.MUST_NOT, 0)], 0),
             ("bool", [(ob.MUST_NOT, 0)], 0),                   # nothing excluded: every live doc, score 0
             ("bool", [(ob.SHOULD, 5), (ob.MUST_NOT, 8)], 0)]
    for i in range(30):
        n_f, n_m, n_n = int(rng.integers(1, 3)), int(rng.integers(0, 3)), int(rng.integers(0, 2))
        terms = [int(x) for x in rng.choice(len(dfs), size=n_f + n_m + n_n, replace=False)]
        cl = [(ob.FILTER, t) for t in terms[:n_f]] + [(ob.MUST, t) for t in terms[n_f:n_f + n_m]]
        cl += [(ob.MUST_NOT, t) for t in terms[n_f + n_m:]]
        specs.append(("bool", [cl[j] for j in rng.permutation(len(cl))], 0))
    specs += _mixed_specs(rng, len(dfs), 12, kinds=("term", "and", "or"))
    for k, rp in ((10, 0), (100, 2500)):
        for mode in (0, 1):
            got, want = _run_both(segs, specs, k, mode=mode, range_postings=rp)
            helpers.assert_same_topdocs(got, want, "filter k=%d rp=%d mode=%d" % (k, rp, mode))


def test_wide_disjunctions_follow_the_disi_priority_queue():
    """>= 10 sub-scorers in a leaf: DisjunctionSumScorer / DisjunctionMaxScorer keep them in a DisiPriorityQueue
    (disjunction_scorer.rs:41-45,118-139) and add the scores in the order of its top_list() walk (util/disi.rs:190-231),
    which k_eval_dpq replays literally.  Leaves where fewer than ten of the terms exist fall back to the SimpleQueue
    kernels; live docs; both collector modes; mixed into a batch with the other shapes."""
    rng = np.random.default_rng(171)
    dfs = [9000, 8000, 7000, 6500, 6000, 5000, 4500, 4000, 3000, 2500, 2000, 1200, 600, 129, 40, 1, 0]
    segs = []
    for s, lf in enumerate((None, 0.85, None)):
        d = list(dfs)
        if s == 2:
            for t in (1, 3, 5, 7, 9, 11, 12, 13):   # only 8 of the terms exist in the last leaf: SimpleQueue there
                d[t] = 0
        segs.append(helpers.build_segment(rng, 15000 + 400 * s, d, live_fraction=lf)[0])
    specs = [("bool", [(ob.SHOULD, t) for t in range(10)], 0),
             ("bool", [(ob.SHOULD, t) for t in range(17)], 0),
             ("bool", [(ob.SHOULD, t, float(rng.choice([0.5, 1.0, 2.0]))) for t in rng.permutation(16)], 0),
             ("bool", [(ob.SHOULD, t) for t in (15, 14, 13, 12, 0, 1, 2, 3, 4, 5, 6)], 1),
             ("dismax", [(t,) for t in range(12)], 0.3),
             ("dismax", [(int(t), float(rng.choice([1.0, 3.0]))) for t in rng.permutation(14)], 0.0),
             ("dismax", [(t,) for t in range(10)], 1.0)]
    for i in range(8):
        n = int(rng.integers(10, 17))
        specs.append(("bool", [(ob.SHOULD, int(t)) for t in rng.choice(17, size=n, replace=False)], 0))
    specs += _mixed_specs(rng, len(dfs), 12, kinds=("term", "and", "or"))
    for k in (10, 100):
        for mode in (0, 1):
            got, want = _run_both(segs, specs, k, mode=mode)
            helpers.assert_same_topdocs(got, want, "dpq k=%d mode=%d" % (k, mode))
    s = search.GpuIndexSearcher(search.IndexReader(segs))
    try:   # wider than the kernel takes, or wide and not a plain disjunction: the caller falls back
        for bad in (("bool", [(ob.SHOULD, t % 17) for t in range(33)], 0),
                    ("bool", [(ob.SHOULD, t) for t in range(11)] + [(ob.MUST_NOT, 12)], 0),
                    ("bool", [(ob.SHOULD, t) for t in range(11)], 2)):
            with pytest.raises(engine.Unsupported):
                s.search_batch(helpers.to_queries([bad]), 10)
    finally:
        s.engine.close()


def test_reference_style_api():
    """Reads like examples/example.rs:111-117."""
    rng = np.random.default_rng(5)
    seg, posts = helpers.build_segment(rng, 10000, [1666, 40, 7])
    reader = search.IndexReader([seg], term_ids={("body", b"hello"): 0, ("body", b"world"): 1})
    searcher = search.GpuIndexSearcher(reader)
    try:
        query = search.TermQuery.new(search.Term.new("body", b"hello"), 1.0, None)
        collector = search.TopDocsCollector.new(10)
        searcher.search(query, collector)
        top = collector.top_docs()
        assert top.total_hits() == 1666 and len(top.score_docs()) == 10
        scores = [d.score for d in top.score_docs()]
        assert scores == sorted(scores, reverse=True)
        collector = search.TopDocsCollector.new(10)
        searcher.search(search.TermQuery.new(search.Term.new("body", b"absent"), 1.0, None), collector)
        assert collector.top_docs().total_hits() == 0 and collector.top_docs().score_docs() == []
        q = search.BooleanQuery.build([query], [search.TermQuery.new(search.Term.new("body", b"world"))], [], [], 0)
        collector = search.TopDocsCollector.new(10)   # MUST + SHOULD: ReqOptScorer
        searcher.search(q, collector)
        assert collector.top_docs().total_hits() == 1666
        collector = search.TopDocsCollector.new(10)   # MUST + FILTER: the filter restricts, only the MUST scores
        searcher.search(search.BooleanQuery.build([query], [], [search.TermQuery.new(search.Term.new("body", b"world"))], [], 0),
                        collector)
        both = np.intersect1d(posts[0][0], posts[1][0])
        assert collector.top_docs().total_hits() == len(both)
        lone = search.BooleanQuery.build([], [], [query], [], 0)   # a lone FILTER: ConstantScoreQuery(boost 0)
        assert isinstance(lone, search.ConstantScoreQuery)
        collector = search.TopDocsCollector.new(10)
        searcher.search(lone, collector)
        assert collector.top_docs().total_hits() == 1666 and all(d.score == 0.0 for d in collector.top_docs().score_docs())
        only_not = search.BooleanQuery.build([], [], [], [query], 0)   # pure MUST_NOT: MatchAllDocsQuery minus the term
        assert isinstance(only_not.must_queries[0], search.MatchAllDocsQuery)
        collector = search.TopDocsCollector.new(10)
        searcher.search(only_not, collector)
        assert collector.top_docs().total_hits() == 10000 - 1666
        with pytest.raises(search.IllegalArgument):
            search.BooleanQuery.build([], [], [], [], 0)
    finally:
        searcher.engine.close()


def test_candidate_arena_overflow_splits_the_batch():
    """A 1 MB candidate arena cannot hold a 300-query, k=1000 batch: rg_search_batch halves the batch (recursively) on
    its own and still equals the oracle; the split calls report RG_ENOMEM to the caller."""
    rng = np.random.default_rng(4242)
    seg, _ = helpers.build_segment(rng, 150000, [60000, 40000, 25000, 12000, 8000, 5000, 3000, 2000, 900, 400])
    ix = helpers.oracle_index([seg])
    specs = _mixed_specs(rng, 10, 300, kinds=("or",))
    q, c = ob.make_queries(specs)
    want = ix.search_batch(q, c, 1000, parallel_mode=0, n_threads=4)
    s = search.GpuIndexSearcher(search.IndexReader([seg]), cand_arena_bytes=1 << 20)
    try:
        got = s.search_batch(helpers.to_queries(specs), 1000)
        helpers.assert_same_topdocs(got, want, "auto split")
        qa, ca = s.compile_batch(helpers.to_queries(specs))
        b = s.engine.prepare(qa, ca, 1000, k1=s.similarity.k1)
        b.run()
        with pytest.raises(engine.EngineError):
            b.fetch()
        b.close()
    finally:
        s.engine.close()
