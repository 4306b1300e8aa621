"""Property tests (hypothesis, CPU): random posting lists through the product's writer and the
oracle's BlockDocIterator — next() stream, advance() landing spots, skip data across block and
level boundaries — for both `.doc` layouts; TopDocsCollector vs a brute-force model of the heap."""
import numpy as np
from hypothesis import given, settings, strategies as st

import helpers
import oracle_binding as ob
from rucene_b200 import codec


@settings(max_examples=25, deadline=None)
@given(seed=st.integers(0, 2**31 - 1), version=st.sampled_from([0, 1]),
       df=st.one_of(st.integers(1, 300), st.sampled_from([127, 128, 129, 1023, 1024, 1025, 8192, 8193, 9000])),
       density=st.sampled_from([0.9, 0.3, 0.01]), ef=st.sampled_from([0, 0, 1, 2]))
def test_random_list_round_trip_and_advance(seed, version, df, density, ef):
    rng = np.random.default_rng(seed)
    max_doc = max(int(df / density) + 10, df + 10)
    docs = np.sort(rng.choice(max_doc, size=df, replace=False)).astype(np.int32)
    freqs = np.minimum(rng.geometric(0.4, size=df), 10**6).astype(np.int32)
    # ef: the reference's dormant EF / BITSET doc-block encodings (1: only where no larger than PF, 2: EF
    # wherever it fits MAX_ENCODED_SIZE) — next() and advance() must see the same posting list
    w = codec.PostingsWriter(doc_version=version, max_doc=max_doc, use_ef=ef > 0, with_pf=ef != 2)
    w.add_term(docs, freqs)
    seg = w.finish(norms=np.full(max_doc, 100, np.uint8))
    ix = helpers.oracle_index([seg])
    d, f = ix.postings(0, 0, df + 3)
    assert np.array_equal(d, docs) and np.array_equal(f, freqs)
    n = int(rng.integers(1, 40))
    targets = np.unique(rng.integers(0, max_doc + 50, size=n)).astype(np.int32)
    out_d, out_f = ix.advance_seq(0, 0, targets)
    cur = -1
    for tg, dd, ff in zip(targets, out_d, out_f):
        if tg <= cur:
            break
        i = int(np.searchsorted(docs, tg))
        if i >= df:
            assert dd == ob.NO_MORE_DOCS
            break
        assert dd == docs[i] and ff == freqs[i]
        cur = int(dd)


@settings(max_examples=30, deadline=None)
@given(seed=st.integers(0, 2**31 - 1), version=st.sampled_from([0, 1]), ef=st.sampled_from([0, 1, 2]),
       df=st.sampled_from([130, 257, 1000, 1024, 4097, 9000]), density=st.sampled_from([0.95, 0.4, 0.02]))
def test_interleaved_next_and_advance(seed, version, ef, df, density):
    """What ConjunctionScorer does to a list: next() and advance(target > doc) in any interleaving, over PF,
    EF and BITSET blocks (EliasFanoDecoder::advance_to_value then next_value, FixedBitSet advance then
    next_set_bit, skip data in between) — against a plain sorted-array model."""
    rng = np.random.default_rng(seed)
    max_doc = max(int(df / density) + 10, df + 10)
    docs = np.sort(rng.choice(max_doc, size=df, replace=False)).astype(np.int32)
    freqs = np.minimum(rng.geometric(0.4, size=df), 10**6).astype(np.int32)
    w = codec.PostingsWriter(doc_version=version, max_doc=max_doc, use_ef=ef > 0, with_pf=ef != 2)
    w.add_term(docs, freqs)
    seg = w.finish(norms=np.full(max_doc, 100, np.uint8))
    ix = helpers.oracle_index([seg])
    ops, want_d, want_f = [], [], []
    pos, cur = -1, -1            # index of the current posting, current docid
    for _ in range(int(rng.integers(5, 120))):
        if pos >= df:
            break
        if rng.random() < 0.5:
            ops.append(-1)
            pos += 1
        else:
            jump = int(rng.choice([1, 2, 5, 40, 200, 3000]))
            tgt = cur + 1 + int(rng.integers(0, jump))
            ops.append(tgt)
            pos = max(pos + 1, int(np.searchsorted(docs, tgt)))
        if pos >= df:
            want_d.append(ob.NO_MORE_DOCS)
            want_f.append(0)
        else:
            cur = int(docs[pos])
            want_d.append(cur)
            want_f.append(int(freqs[pos]))
    got_d, got_f = ix.advance_seq(0, 0, np.array(ops, np.int32))
    assert list(got_d) == want_d
    assert [int(f) for d, f in zip(got_d, got_f) if d != ob.NO_MORE_DOCS] == [f for d, f in zip(want_d, want_f) if d != ob.NO_MORE_DOCS]


def _heap_model(stream, k):
    """Direct transcription of SURVEY Appendix B (independent of the oracle's C++)."""
    data = []

    def le(a, b):
        return a[1] >= b[1]

    def ge(a, b):
        return a[1] <= b[1]

    def sift_up(start, pos):
        e = data[pos]
        while pos > start:
            p = (pos - 1) // 2
            if le(e, data[p]):
                break
            data[pos] = data[p]
            pos = p
        data[pos] = e

    def sift_down_range(pos, end):
        e = data[pos]
        c = 2 * pos + 1
        while c < end:
            r = c + 1
            if r < end and le(data[c], data[r]):
                c = r
            if ge(e, data[c]):
                break
            data[pos] = data[c]
            pos = c
            c = 2 * pos + 1
        data[pos] = e

    def sift_down_to_bottom(pos):
        end, start = len(data), pos
        e = data[pos]
        c = 2 * pos + 1
        while c < end:
            r = c + 1
            if r < end and le(data[c], data[r]):
                c = r
            data[pos] = data[c]
            pos = c
            c = 2 * pos + 1
        data[pos] = e
        sift_up(start, pos)

    total = 0
    for d, s in stream:
        total += 1
        if len(data) < k:
            data.append((d, s))
            sift_up(0, len(data) - 1)
        elif data[0][1] < s:
            data[0] = (d, s)
            sift_down_range(0, len(data))
    out = []
    for _ in range(min(total, len(data))):
        item = data.pop()
        if data:
            item, data[0] = data[0], item
            sift_down_to_bottom(0)
        out.append(item)
    return out[::-1]


@settings(max_examples=60, deadline=None)
@given(scores=st.lists(st.integers(0, 6), min_size=0, max_size=200), k=st.integers(1, 12))
def test_top_docs_collector_matches_heap_model(scores, k):
    """Heavy ties (scores from a 7-value alphabet): survivors and order follow the heap layout."""
    docs = list(range(len(scores)))
    sc = [float(s) for s in scores]
    got, _heap = ob.topk_stream(docs, sc, k)
    want = _heap_model(list(zip(docs, sc)), k)
    assert [(int(h["doc"]), float(h["score"])) for h in got] == want
