"""Oracle vs an independent numpy model for the non-scoring pieces of BooleanQuery: FILTER clauses
(boolean_query.rs:108-110 -> NonScoringSimilarity, score 0f32, searcher.rs:158-197), a lone FILTER
(ConstantScoreQuery::with_boost(filter, 0), :66-75) and a pure MUST_NOT query (MatchAllDocsQuery, :76-79)."""
import numpy as np

import helpers
import oracle_binding as ob


def _live(seg, docs):
    if seg.live_docs is None:
        return np.ones(len(docs), bool)
    return ((seg.live_docs[docs >> 6] >> (docs & 63).astype(np.uint64)) & np.uint64(1)) == 1


def _model(segs, posts, ix, clauses):
    """-> (docs, scores) in collection order for musts/filters/shoulds/must_nots of TermQuerys (no MUST+SHOULD
    with scoring MUSTs here: the ReqOpt chain has its own test)."""
    out_d, out_s = [], []
    base = 0
    for seg, postings in zip(segs, posts):
        def lst(t, boost=1.0):
            w, _idf, _avgdl, cache = ix.term_weight(t, boost)
            d, f = postings[t]
            return d, helpers.bm25_scores_numpy(w, 1.2, f, seg.norms[d], cache)
        musts = [lst(c[1]) for c in clauses if c[0] == ob.MUST]
        filters = [postings[c[1]][0] for c in clauses if c[0] == ob.FILTER]
        shoulds = [lst(c[1]) for c in clauses if c[0] == ob.SHOULD]
        nots = [postings[c[1]][0] for c in clauses if c[0] == ob.MUST_NOT]
        shoulds = [s for s in shoulds if len(s[0])]
        nots = [n for n in nots if len(n)]
        if musts or filters:
            req_lists = [m[0] for m in musts] + filters
            if any(len(x) == 0 for x in req_lists):
                base += seg.max_doc
                continue
            docs = req_lists[0]
            for x in req_lists[1:]:
                docs = np.intersect1d(docs, x)
            score = np.zeros(len(docs), np.float32)
            order = sorted(range(len(musts)), key=lambda i: len(musts[i][0]))  # cost order among the scoring ones
            for j, i in enumerate(order):
                part = musts[i][1][np.searchsorted(musts[i][0], docs)]
                score = part.copy() if j == 0 else (score + part).astype(np.float32)
            if shoulds:   # ReqOptScorer with a required side that scores 0 (filters only) or musts (not used here)
                assert not musts
                od = np.unique(np.concatenate([s[0] for s in shoulds]))
                ov = np.zeros(len(od), np.float32)
                for d, s in shoulds:
                    p = np.searchsorted(od, d)
                    ov[p] = (ov[p] + s).astype(np.float32)
                p = np.minimum(np.searchsorted(od, docs), len(od) - 1)
                has = od[p] == docs
                score = np.where(has, (score + ov[p]).astype(np.float32), score)
        elif shoulds:
            docs = np.unique(np.concatenate([s[0] for s in shoulds]))
            score = np.zeros(len(docs), np.float32)
            for d, s in shoulds:
                p = np.searchsorted(docs, d)
                score[p] = (score[p] + s).astype(np.float32)
        else:  # MatchAllDocsQuery
            docs = np.arange(seg.max_doc, dtype=np.int32)
            score = np.zeros(len(docs), np.float32)
        if nots:
            excl = np.unique(np.concatenate(nots))
            keep = ~np.isin(docs, excl)
            docs, score = docs[keep], score[keep]
        keep = _live(seg, docs)
        out_d.append(docs[keep] + base)
        out_s.append(score[keep])
        base += seg.max_doc
    if not out_d:
        return np.zeros(0, np.int32), np.zeros(0, np.float32)
    return np.concatenate(out_d).astype(np.int32), np.concatenate(out_s).astype(np.float32)


def test_filter_clauses_and_match_all_follow_the_reference():
    rng = np.random.default_rng(12)
    dfs = [0, 1, 40, 129, 900, 4000, 9000, 15000]
    segs, posts = [], []
    for s in range(2):
        seg, p = helpers.build_segment(rng, 20000 + 300 * s, dfs, live_fraction=0.85 if s else None)
        segs.append(seg)
        posts.append(p)
    ix = helpers.oracle_index(segs)
    specs = [[(ob.FILTER, 6)],                                   # lone filter: its docs, score 0
             [(ob.MUST, 7), (ob.FILTER, 6)],
             [(ob.FILTER, 7), (ob.MUST, 5), (ob.MUST, 6)],
             [(ob.FILTER, 7), (ob.FILTER, 6)],
             [(ob.FILTER, 7), (ob.SHOULD, 6), (ob.SHOULD, 4)],   # required side scores 0; optional side adds
             [(ob.FILTER, 7), (ob.MUST_NOT, 6)],
             [(ob.MUST, 7), (ob.FILTER, 0)],                     # filter term absent: nothing
             [(ob.MUST_NOT, 7)],                                 # MatchAllDocsQuery minus the term
             [(ob.MUST_NOT, 6), (ob.MUST_NOT, 4), (ob.MUST_NOT, 0)],
             [(ob.SHOULD, 5), (ob.MUST_NOT, 7)]]
    q, c = ob.make_queries([("bool", cl, 0) for cl in specs])
    for k in (1, 10, 100):
        hits, counts, total = ix.search_batch(q, c, k)
        for i, cl in enumerate(specs):
            d, s = _model(segs, posts, ix, cl)
            want, _ = ob.topk_stream(d, s, k)
            assert total[i] == len(d), (cl, total[i], len(d))
            assert counts[i] == len(want)
            got = hits[i][:counts[i]]
            assert np.array_equal(got["doc"], want["doc"]), cl
            assert np.array_equal(got["score"].view(np.uint32), want["score"].view(np.uint32)), cl
