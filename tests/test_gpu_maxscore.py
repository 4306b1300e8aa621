"""GPU parity of the presence-bitmap / score-column / non-essential-clause route (k_eval_or_ms) and of
what surrounds it: persistent columns across batches, invalidation, stale batches, the segment-wide
block decode (BASELINE config 2, realistic blocks)."""
import numpy as np
import pytest

import helpers
import oracle_binding as ob
from rucene_b200 import codec, engine, search

pytestmark = pytest.mark.gpu


def _or_specs(rng, n_terms, n, t_min=1, t_max=7, boosts=False):
    specs = []
    for ts in helpers.distinct_query_terms(rng, n_terms, n, t_min, t_max):
        if boosts:
            specs.append(("bool", [(ob.SHOULD, t, float(rng.choice([0.25, 1.0, 1.0, 3.0, 10.0]))) for t in ts], 0))
        else:
            specs.append(("bool", [(ob.SHOULD, t) for t in ts], 0))
    return specs


def _check(searcher, ix, specs, k, label, mode=0):
    q, c = ob.make_queries(specs)
    want = ix.search_batch(q, c, k, parallel_mode=mode, n_threads=4)
    got = searcher.search_batch(helpers.to_queries(specs), k, mode=mode)
    helpers.assert_same_topdocs(got, want, label)


@pytest.mark.parametrize("k", [1, 10, 100])
def test_zipf_disjunctions_all_routes(k):
    """Log-uniform term ranks over a Zipfian segment: dense clauses become score columns (non-essential once
    theta has risen), sparse ones stay block streams; many docid ranges per query, theta chaining."""
    seg = codec.synth_segment(0x5EED0001, 400000, 4000, doc_version=1)
    ix = helpers.oracle_index([seg])
    rng = np.random.default_rng(20 + k)
    specs = _or_specs(rng, 4000, 70) + [("term", 0), ("term", 7), ("term", 300)]
    for flags in (engine.CFG_EAGER_COLUMNS | engine.CFG_MAXSCORE, engine.CFG_MAXSCORE | engine.CFG_TFPLANES, engine.CFG_MAXSCORE, 0):
        for rp in (0, 6000):
            s = search.GpuIndexSearcher(search.IndexReader([seg]), range_postings=rp, flags=flags)
            try:
                _check(s, ix, specs, k, "zipf k=%d flags=%d rp=%d" % (k, flags, rp))
                if flags & engine.CFG_EAGER_COLUMNS:
                    assert s.engine.column_stats()["cached"] > 0
            finally:
                s.engine.close()


def test_boosts_change_the_essential_order_and_negative_boost_is_never_pruned():
    seg = codec.synth_segment(0x5EED0011, 250000, 2500, doc_version=1)
    ix = helpers.oracle_index([seg])
    rng = np.random.default_rng(5)
    specs = _or_specs(rng, 2500, 60, 2, 6, boosts=True)
    specs += [("bool", [(ob.SHOULD, 0, -1.0), (ob.SHOULD, 3), (ob.SHOULD, 40)], 0),
              ("bool", [(ob.SHOULD, 1, 0.0), (ob.SHOULD, 2), (ob.SHOULD, 900)], 0),
              ("bool", [(ob.SHOULD, 0), (ob.SHOULD, 1), (ob.SHOULD, 2), (ob.SHOULD, 3), (ob.SHOULD, 4)], 0)]
    s = search.GpuIndexSearcher(search.IndexReader([seg]), range_postings=8000, flags=engine.CFG_EAGER_COLUMNS | engine.CFG_MAXSCORE | engine.CFG_TFPLANES)
    try:
        for k in (5, 50):
            _check(s, ix, specs, k, "boosts k=%d" % k)
    finally:
        s.engine.close()


@pytest.mark.parametrize("mode", [0, 1])
def test_leaves_live_docs_and_clustered_streams(mode):
    """Several leaves with deleted docs; sparse terms whose postings are clustered (consecutive docids) so a
    stream's cached block ends inside a window and cuts it; massive score ties (constant freq / norm)."""
    rng = np.random.default_rng(77 + mode)
    segs = []
    for s_i, md in enumerate((90000, 70000, 81000)):
        w = codec.PostingsWriter(doc_version=1, max_doc=md)
        dfs = [60000, 30000, 9000, 3000, 700, 300, 129, 5, 1]
        for t, df in enumerate(dfs):
            if t in (3, 4):  # clustered: runs of consecutive docids
                starts = np.sort(rng.choice(md - 600, size=df // 100, replace=False))
                docs = np.unique(np.concatenate([np.arange(a, a + 100) for a in starts]))[:df].astype(np.int32)
            else:
                docs = np.sort(rng.choice(md, size=df, replace=False)).astype(np.int32)
            freqs = np.ones(len(docs), np.int32) if t in (0, 4) else (1 + rng.geometric(0.5, size=len(docs)) - 1).clip(1, 255).astype(np.int32)
            w.add_term(docs, freqs)
        norms = np.full(md, 120, np.uint8) if s_i == 1 else rng.integers(100, 140, md).astype(np.uint8)
        live = None
        if s_i != 0:
            bits = rng.random(md) < 0.8
            words = np.zeros((md + 63) // 64, np.uint64)
            idx = np.nonzero(bits)[0]
            np.bitwise_or.at(words, idx >> 6, np.uint64(1) << (idx & 63).astype(np.uint64))
            live = words
        segs.append(w.finish(norms=norms, live_docs=live))
    ix = helpers.oracle_index(segs)
    specs = [("term", t) for t in range(9)]
    for i in range(60):
        t = int(rng.integers(2, 7))
        specs.append(("bool", [(ob.SHOULD, int(x)) for x in rng.choice(9, size=t, replace=False)], 0))
    for flags in (engine.CFG_EAGER_COLUMNS | engine.CFG_MAXSCORE | engine.CFG_TFPLANES, engine.CFG_MAXSCORE, 0):
        s = search.GpuIndexSearcher(search.IndexReader(segs), range_postings=4000, flags=flags)
        try:
            for k in (3, 100):
                _check(s, ix, specs, k, "leaves mode=%d flags=%d k=%d" % (mode, flags, k), mode=mode)
        finally:
            s.engine.close()


def test_columns_persist_across_batches_and_follow_the_norm_cache():
    seg = codec.synth_segment(0x5EED0021, 300000, 3000, doc_version=1)
    rng = np.random.default_rng(9)
    s = search.GpuIndexSearcher(search.IndexReader([seg]))
    try:
        ix = helpers.oracle_index([seg])
        a = _or_specs(rng, 3000, 40, 2, 5)
        _check(s, ix, a, 10, "first batch")
        st1 = s.engine.column_stats()
        assert st1["built"] > 0 and st1["cached"] == st1["built"]
        b = a[:20] + _or_specs(rng, 3000, 30, 2, 5)
        _check(s, ix, b, 10, "second batch")
        st2 = s.engine.column_stats()
        assert st2["hits"] > st1["hits"]          # columns of the first batch were reused
        l1 = s.engine.list_stats()
        assert l1["built"] > 0 and l1["hits"] > 0  # and so were the scored posting lists of its mid-frequency clauses
        # a batch prepared before the cache changes must not run afterwards
        q, c = s.compile_batch(helpers.to_queries(a))
        stale = s.engine.prepare(q, c, 10)
        # new similarity parameters: same cache id, new contents -> cached columns are dropped
        cache2 = codec.bm25_norm_cache(1.2, 0.3, s._avgdl)
        s.engine.set_norm_cache(0, cache2)
        with pytest.raises(engine.EngineError):
            stale.run()
        stale.close()
        assert s.engine.column_stats()["cached"] == 0
        assert s.engine.list_stats()["cached"] == 0
        ix2 = helpers.oracle_index([seg], b=0.3)
        _check(s, ix2, b, 10, "after the norm cache changed")
        assert s.engine.list_stats()["cached"] > 0
    finally:
        s.engine.close()


def test_segment_decode_matches_the_oracle_reader():
    """rg_segment_decode: every block pair of the uploaded segment -> 128 doc deltas + 128 freqs."""
    seg = codec.synth_segment(0x5EED0031, 120000, 600, doc_version=1)
    ix = helpers.oracle_index([seg])
    eng = engine.Engine()
    try:
        eng.upload_segment(seg)
        stats, out = eng.segment_decode(0, fetch=True)
        assert stats["blocks"] == stats["segment_blocks"] == out.shape[0] > 100
        blk = 0
        for t in range(600):
            df = int(seg.terms["doc_freq"][t])
            nb = df // 128
            if nb == 0:
                continue
            docs, freqs = ix.postings(0, t, df + 1)
            d = docs[:nb * 128].astype(np.int64)
            deltas = np.diff(np.concatenate([[0], d])).reshape(nb, 128)
            assert np.array_equal(out[blk:blk + nb, 0, :], deltas), t
            assert np.array_equal(out[blk:blk + nb, 1, :], freqs[:nb * 128].reshape(nb, 128)), t
            blk += nb
        assert blk == out.shape[0]
    finally:
        eng.close()


def test_device_terms_dictionary_resolves_a_batch():
    """rg_terms_upload / rg_terms_lookup: exact lookups of the batch's Term bytes on the device (the reference's
    per-query SegmentTermIterator::seek_exact, blocktree_reader.rs:1364) — absent terms, prefixes of present terms,
    the empty term, long terms, bytes 0x00 / 0xff, terms present in one leaf only — then the same searches as with
    the host-side dictionary."""
    rng = np.random.default_rng(99)
    names = [b"", b"\x00", b"\x00\x00", b"a", b"ab", b"abc", b"abd", b"b" * 300, b"b" * 301, b"\xff", b"\xff\xff",
             b"zebra", b"zebr\xc3\xa4", b"hello", b"world"]
    dfs = [int(x) for x in rng.integers(50, 9000, len(names))]
    segs = []
    for s in range(2):
        d = list(dfs)
        if s == 1:
            d[3] = 0      # b"a" lives in the first leaf only
            d[13] = 0
        segs.append(helpers.build_segment(rng, 20000 + 500 * s, d)[0])
    term_ids = {("body", n): i for i, n in enumerate(names)}
    reader = search.IndexReader(segs, term_ids=term_ids)
    dev = search.GpuIndexSearcher(reader, device_terms=True)
    host = search.GpuIndexSearcher(reader)
    try:
        probes = names + [b"abcd", b"aa", b"\x00\x00\x00", b"b" * 299, b"b" * 302, b"zebr", b"zebrb", b"\xfe", b"nope"]
        ids, df = dev.engine.lookup_terms(probes)
        for i, pb in enumerate(probes):
            want = term_ids.get(("body", pb))
            assert (None if ids[i] == 0xFFFFFFFF else int(ids[i])) == want, pb
            for s in range(2):
                assert df[s][i] == (0 if want is None else int(segs[s].terms["doc_freq"][want])), (pb, s)
        T = lambda b, boost=1.0: search.TermQuery.new(search.Term.new("body", b), boost, None)   # noqa: E731
        queries = [T(b"hello"), T(b"nope"), T(b"a"),
                   search.BooleanQuery.build([T(b"abc"), T(b"abd")], [], [], [], 0),
                   search.BooleanQuery.build([], [T(b""), T(b"\xff"), T(b"b" * 300), T(b"absent"), T(b"zebra", 2.0)], [], [], 0),
                   search.BooleanQuery.build([T(b"world")], [T(b"a")], [T(b"ab")], [T(b"\x00")], 0),
                   search.DisjunctionMaxQuery.build([T(b"hello"), T(b"world"), T(b"zebr\xc3\xa4")], 0.2)]
        helpers.assert_same_topdocs(dev.search_batch(queries, 10), host.search_batch(queries, 10), "device dictionary")
        with pytest.raises(engine.EngineError):   # not in dictionary order
            dev.engine.upload_terms(0, [b"b", b"a"], [0, 1])
    finally:
        dev.engine.close()
        host.engine.close()


def test_abi_rejects_bad_clause_and_corrupt_index_fields():
    """The C ABI validates what the kernels later use as indices: a MUST_NOT clause with an unset norm cache id
    (kernels form cache pointers from it even though the clause never scores) and index fields that point outside
    the segment are RG_EINVAL, not out-of-bounds device accesses."""
    rng = np.random.default_rng(3)
    seg, _ = helpers.build_segment(rng, 20000, [5000, 900, 1])
    eng = engine.Engine()
    try:
        eng.upload_segment(seg)
        eng.set_norm_cache(0, codec.bm25_norm_cache(1.2, 0.75, 100.0))
        q = np.zeros(1, engine.QUERY_DTYPE)
        q["n_clauses"], q["flags"] = 2, engine.Q_BOOLEAN
        c = np.zeros(2, engine.CLAUSE_DTYPE)
        c["occur"] = [engine.SHOULD, engine.MUST_NOT]
        c["term_id"] = [0, 1]
        c["weight"] = 1.0
        c["cache_id"] = [0, 0x7FFFFFFF]      # garbage on the clause that "does not score"
        with pytest.raises(engine.EngineError):
            eng.search_batch(q, c, 10)
        c["cache_id"] = [0, 0]
        eng.search_batch(q, c, 10)
        # corrupt BlockTermState rows
        bad = np.array(seg.terms, copy=True)
        bad["singleton_doc_id"][2] = 20000          # singleton docid == max_doc
        e2 = engine.Engine()
        try:
            with pytest.raises(engine.EngineError):
                e2.upload_segment(type("S", (), dict(terms=bad, doc_file=seg.doc_file, norms=seg.norms, live_docs=None, max_doc=20000))())
            with pytest.raises(engine.EngineError):    # max_doc smaller than the docids in the file
                e2.upload_segment(type("S", (), dict(terms=seg.terms, doc_file=seg.doc_file, norms=seg.norms[:6000], live_docs=None, max_doc=6000))())
        finally:
            e2.close()
    finally:
        eng.close()


def test_zero_score_matches_stay_on_the_general_disjunction_kernel():
    """Lucene's norm table maps norm byte 0 to an infinite field length, so such a doc's BM25 contribution is exactly
    0f32: it still matches (total_hits, and it is collected while the heap is not full).  The plain-sum kernel
    variant tells matches by a non-zero sum and score columns use +0.0f for "no posting" — both must step aside for a
    leaf whose norms select such a cache entry, and for zero / negative boosts."""
    seg = codec.synth_segment(0x5EED0077, 120000, 1500, doc_version=1)
    seg.norms[::5] = 0
    ix = helpers.oracle_index([seg])
    rng = np.random.default_rng(77)
    specs = _or_specs(rng, 1500, 40) + [("term", 0), ("term", 900),
                                         ("bool", [(ob.SHOULD, 0, 0.0), (ob.SHOULD, 5, 0.0)], 0),
                                         ("bool", [(ob.SHOULD, 2, 1e-30), (ob.SHOULD, 700, 1e-30)], 0)]
    for flags in (0, engine.CFG_EAGER_COLUMNS, engine.CFG_EAGER_COLUMNS | engine.CFG_MAXSCORE):
        s = search.GpuIndexSearcher(search.IndexReader([seg]), range_postings=5000, flags=flags)
        try:
            for k in (3, 1000):
                _check(s, ix, specs, k, "zero scores flags=%d k=%d" % (flags, k))
            assert s.engine.column_stats()["cached"] == 0  # no clause of this leaf is provably positive
        finally:
            s.engine.close()
    # the same queries over a leaf with ordinary norms: columns are built and the plain-sum variant runs
    seg2 = codec.synth_segment(0x5EED0077, 120000, 1500, doc_version=1)
    ix2 = helpers.oracle_index([seg2])
    s = search.GpuIndexSearcher(search.IndexReader([seg2]), range_postings=5000, flags=engine.CFG_EAGER_COLUMNS)
    try:
        _check(s, ix2, specs, 10, "ordinary norms")
        assert s.engine.column_stats()["cached"] > 0
    finally:
        s.engine.close()


def test_scored_lists_every_block_layout_and_tails():
    """Scored posting lists (k_build_columns<4>) must reproduce stream_refill for PF / EF / BITSET doc blocks, both
    .doc versions, vint tails, docid ranges that cut blocks, live docs and several leaves; with and without them the
    TopDocs are the oracle's."""
    rng = np.random.default_rng(404)
    dfs = [40000, 25000, 12000, 6000, 3000, 1500, 700, 385, 300, 129, 128, 127, 5]
    for version, use_ef in ((1, False), (0, False), (1, True)):
        segs = []
        for i in range(2):
            seg, _ = helpers.build_segment(rng, 50000, dfs, doc_version=version, use_ef=use_ef, live_fraction=0.9 if i else None)
            segs.append(seg)
        ix = helpers.oracle_index(segs)
        specs = [("term", t) for t in range(len(dfs))] + _or_specs(rng, len(dfs), 60, 2, 6)
        for flags in (engine.CFG_EAGER_COLUMNS, engine.CFG_EAGER_COLUMNS | engine.CFG_NO_COLUMNS, engine.CFG_NO_LISTS):
            for mode in (0, 1):
                s = search.GpuIndexSearcher(search.IndexReader(segs), range_postings=3000, flags=flags)
                try:
                    _check(s, ix, specs, 20, "lists v%d ef%d flags=%d mode=%d" % (version, use_ef, flags, mode), mode=mode)
                    assert (s.engine.list_stats()["cached"] > 0) == (not flags & engine.CFG_NO_LISTS)
                finally:
                    s.engine.close()


def test_two_batches_in_flight():
    """rg_batch_prepare of the next batch while one runs (plan uploads and result fetches ride the copy stream): both
    orders of fetching, a batch closed without being fetched, and scored lists built by the second prepare while
    the first batch is still queued."""
    seg = codec.synth_segment(0x5EED0099, 200000, 2000, doc_version=1)
    ix = helpers.oracle_index([seg])
    rng = np.random.default_rng(99)
    sets = [_or_specs(rng, 2000, 50, 2, 6) for _ in range(4)]
    want = []
    for sp in sets:
        q, c = ob.make_queries(sp)
        want.append(ix.search_batch(q, c, 10, parallel_mode=0, n_threads=4))
    s = search.GpuIndexSearcher(search.IndexReader([seg]), range_postings=4000, flags=engine.CFG_EAGER_COLUMNS)
    try:
        arrays = [s.compile_batch(helpers.to_queries(sp)) for sp in sets]
        k1 = s.similarity.k1
        a = s.engine.prepare(arrays[0][0], arrays[0][1], 10, k1=k1)
        a.run()
        b = s.engine.prepare(arrays[1][0], arrays[1][1], 10, k1=k1)   # planned while a is (possibly) still running
        b.run()
        c = s.engine.prepare(arrays[2][0], arrays[2][1], 10, k1=k1)
        c.run()
        d = s.engine.prepare(arrays[3][0], arrays[3][1], 10, k1=k1)
        c.close()                                                     # never fetched
        got_b = b.fetch()                                             # out of order
        got_a = a.fetch()
        d.run()
        got_d = d.fetch()
        got_a2 = a.fetch()                                            # a second fetch of a finished batch
        for x in (a, b, d):
            x.close()
        helpers.assert_same_topdocs(got_a, want[0], "in flight: a")
        helpers.assert_same_topdocs(got_a2, want[0], "in flight: a again")
        helpers.assert_same_topdocs(got_b, want[1], "in flight: b")
        helpers.assert_same_topdocs(got_d, want[3], "in flight: d")
    finally:
        s.engine.close()


def test_scored_list_arena_wraps_and_reclaims(monkeypatch):
    """A list arena of 768 KB: successive batches over different terms fill it, the ring wraps, the oldest slabs are
    reclaimed (their lists leave the cache), a slab a prepared batch still references is never reclaimed — and every
    batch equals the oracle throughout."""
    monkeypatch.setenv("RG_LIST_ARENA_KB", "768")
    seg = codec.synth_segment(0x5EED00AA, 150000, 3000, doc_version=1)
    ix = helpers.oracle_index([seg])
    rng = np.random.default_rng(170)
    s = search.GpuIndexSearcher(search.IndexReader([seg]), range_postings=6000, flags=engine.CFG_EAGER_COLUMNS | engine.CFG_NO_COLUMNS)
    try:
        k1 = s.similarity.k1
        first = _or_specs(rng, 3000, 12, 2, 4)
        qa, ca = s.compile_batch(helpers.to_queries(first))
        pinned = s.engine.prepare(qa, ca, 10, k1=k1)           # holds its lists (the first slab) until closed
        seen_max, reclaimed = 0, False
        for i in range(40):
            specs = _or_specs(rng, 3000, 12, 2, 4)
            _check(s, ix, specs, 10, "arena round %d" % i)
            st = s.engine.list_stats()
            assert st["bytes"] <= 768 * 1024
            reclaimed = reclaimed or st["cached"] < seen_max
            seen_max = max(seen_max, st["cached"])
            if i == 20:                                          # the ring may now pass the first slab
                pinned.run()
                helpers.assert_same_topdocs(pinned.fetch(), ix.search_batch(*ob.make_queries(first), 10, parallel_mode=0, n_threads=2), "pinned batch")
                pinned.close()
        assert s.engine.list_stats()["built"] > seen_max        # more lists were built than ever fit: space was reused
        assert reclaimed
    finally:
        s.engine.close()
