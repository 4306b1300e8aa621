"""Shared test helpers: random posting lists, a brute-force numpy model of the scoring path
(an independent second opinion on the oracle for small cases), query generators."""
import numpy as np

import oracle_binding as ob
from rucene_b200 import codec


def random_postings(rng, max_doc, df):
    docs = np.sort(rng.choice(max_doc, size=df, replace=False)).astype(np.int32)
    freqs = (1 + rng.geometric(0.5, size=df) - 1).astype(np.int32)
    freqs = np.minimum(freqs, 255)
    return docs, freqs


def build_segment(rng, max_doc, dfs, doc_version=1, live_fraction=None, dense_terms=(), use_ef=False,
                  with_pf=True, counts=None):
    """dfs: list of document frequencies (0 allowed = absent term).  use_ef/with_pf: write EF / BITSET
    doc blocks where the reference's (dormant) writer rule picks them; counts: list that receives
    (blocks, ef_blocks, bitset_blocks)."""
    w = codec.PostingsWriter(doc_version=doc_version, max_doc=max_doc, use_ef=use_ef, with_pf=with_pf)
    postings = []
    for t, df in enumerate(dfs):
        if df == 0:
            w.add_term([], [])
            postings.append((np.zeros(0, np.int32), np.zeros(0, np.int32)))
            continue
        if t in dense_terms:  # consecutive docids + constant freq => all-equal blocks
            start = int(rng.integers(0, max_doc - df + 1))
            docs = np.arange(start, start + df, dtype=np.int32)
            freqs = np.full(df, 3, np.int32)
        else:
            docs, freqs = random_postings(rng, max_doc, df)
        w.add_term(docs, freqs)
        postings.append((docs, freqs))
    lens = np.clip(np.round(np.exp(rng.normal(np.log(200), 0.5, max_doc))), 1, 10000).astype(np.int32)
    norms = np.array([codec.encode_norm_value(1.0, int(x)) for x in lens], dtype=np.uint8)
    live = None
    if live_fraction is not None:
        bits = rng.random(max_doc) < live_fraction
        words = np.zeros((max_doc + 63) // 64, np.uint64)
        idx = np.nonzero(bits)[0]
        np.bitwise_or.at(words, idx >> 6, np.uint64(1) << (idx & 63).astype(np.uint64))
        live = words
    if counts is not None:
        counts.append(w.block_counts())
    seg = w.finish(norms=norms, live_docs=live)
    return seg, postings


def oracle_index(segs, k1=1.2, b=0.75):
    ix = ob.Index(k1, b)
    for s in segs:
        ix.add_segment(s)
    return ix


def bm25_scores_numpy(weight, k1, freqs, norms_bytes, cache):
    """weight*(k1+1)*freq/(freq+cache[norm]) in f32, left to right (bm25_similarity.rs:203-212)."""
    f = freqs.astype(np.float32)
    t1 = np.float32(weight) * (np.float32(k1) + np.float32(1.0))
    t2 = (t1 * f).astype(np.float32)
    t3 = (f + cache[norms_bytes]).astype(np.float32)
    return (t2 / t3).astype(np.float32)


def log_uniform_ranks(rng, n_terms, size):
    """r = floor(V^U) in [1,V]; returned as 0-based term ids."""
    r = np.floor(np.power(float(n_terms), rng.random(size))).astype(np.int64)
    return np.clip(r, 1, n_terms) - 1


def distinct_query_terms(rng, n_terms, n_queries, t_min, t_max):
    out = []
    for _ in range(n_queries):
        t = int(rng.integers(t_min, t_max + 1))
        s = []
        while len(s) < t:
            c = int(log_uniform_ranks(rng, n_terms, 1)[0])
            if c not in s:
                s.append(c)
        out.append(s)
    return out


def to_queries(specs):
    """specs (see oracle_binding.make_queries) -> rucene_b200.search Query objects."""
    from rucene_b200 import search as S
    out = []
    for s in specs:
        if s[0] == "term":
            out.append(S.TermQuery.new(S.Term.new("body", s[1]), s[2] if len(s) > 2 else 1.0, None))
        elif s[0] == "dismax":
            out.append(S.DisjunctionMaxQuery.build(
                [S.TermQuery.new(S.Term.new("body", cl[0]), cl[1] if len(cl) > 1 else 1.0, None) for cl in s[1]], s[2]))
        else:
            musts, shoulds, filters, nots = [], [], [], []
            for cl in s[1]:
                q = S.TermQuery.new(S.Term.new("body", cl[1]), cl[2] if len(cl) > 2 else 1.0, None)
                (musts if cl[0] == ob.MUST else shoulds if cl[0] == ob.SHOULD else filters if cl[0] == ob.FILTER
                 else nots).append(q)
            out.append(S.BooleanQuery.build(musts, shoulds, filters, nots, s[2] if len(s) > 2 else 0))
    return out


def assert_same_topdocs(got, want, label=""):
    """got/want: (hits, counts, total) triples; bit-exact docids, scores, order and total_hits."""
    gh, gc, gt = got
    wh, wc, wt = want
    assert np.array_equal(gt, wt), (label, "total_hits", np.nonzero(gt != wt)[0][:5], gt[:5], wt[:5])
    assert np.array_equal(gc, wc), (label, "counts")
    for i in range(len(gc)):
        n = int(wc[i])
        if not np.array_equal(gh[i][:n]["doc"], wh[i][:n]["doc"]):
            j = int(np.nonzero(gh[i][:n]["doc"] != wh[i][:n]["doc"])[0][0])
            raise AssertionError((label, "query", i, "rank", j, gh[i][j], wh[i][j]))
        assert np.array_equal(gh[i][:n]["score"].view(np.uint32), wh[i][:n]["score"].view(np.uint32)), \
            (label, "scores of query", i)
