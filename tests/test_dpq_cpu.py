""">= 10 SHOULD clauses: DisjunctionSumScorer switches from SimpleQueue to DisiPriorityQueue
(search/scorer/disjunction_scorer.rs:41-45); the f32 summation order then follows the heap's top_list()
walk (util/disi.rs:190-231), i.e. the whole history of next() calls.  The oracle's C++ restatement is
checked against a second, independent Python model of the same heap."""
import numpy as np

import helpers
import oracle_binding as ob


class _Sub:
    def __init__(self, docs, scores):
        self.docs, self.scores, self.i = docs, scores, -1

    def doc(self):
        return -1 if self.i < 0 else (0x7FFFFFFF if self.i >= len(self.docs) else int(self.docs[self.i]))

    def next(self):
        self.i += 1

    def score(self):
        return self.scores[self.i]


class _Dpq:
    def __init__(self, subs):
        self.heap, self.size = [None] * len(subs), 0
        for s in subs:                       # do_push + up_heap
            i = self.size
            self.heap[i] = s
            while i > 0:
                j = ((i + 1) >> 1) - 1
                if s.doc() >= self.heap[j].doc():
                    break
                self.heap[i] = self.heap[j]
                i = j
            self.heap[i] = s
            self.size += 1

    def update_top(self):                    # down_heap(size)
        size, h = self.size, self.heap
        i, node, j = 0, h[0], 1
        if j < size:
            k = j + 1
            if k < size and h[k].doc() < h[j].doc():
                j = k
            if h[j].doc() < node.doc():
                while True:
                    h[i] = h[j]
                    i = j
                    j = ((i + 1) << 1) - 1
                    k = j + 1
                    if k < size and h[k].doc() < h[j].doc():
                        j = k
                    if j >= size or h[j].doc() >= node.doc():
                        break
                h[i] = node

    def top_list(self):                      # returns the scorers in summation order (head first)
        h, size = self.heap, self.size
        lst = [h[0]]

        def to(i):
            w = h[i]
            if w.doc() == lst[0].doc():
                lst.insert(0, w)
                left, right = ((i + 1) << 1) - 1, ((i + 1) << 1)
                if right < size:
                    to(left)
                    to(right)
                elif left < size and h[left].doc() == lst[0].doc():
                    lst.insert(0, h[left])
        if size >= 3:
            to(1)
            to(2)
        elif size == 2 and h[1].doc() == lst[0].doc():
            lst.insert(0, h[1])
        return lst


def _model(seg, postings, ix, terms):
    subs = []
    for t in terms:
        w, _idf, _avgdl, cache = ix.term_weight(t, 1.0)
        d, f = postings[t]
        if len(d):
            subs.append(_Sub(d, helpers.bm25_scores_numpy(w, 1.2, f, seg.norms[d], cache)))
    q = _Dpq(subs)
    docs, scores = [], []
    while True:
        doc = q.heap[0].doc()
        while True:
            q.heap[0].next()
            q.update_top()
            if q.heap[0].doc() != doc:
                break
        d = q.heap[0].doc()
        if d == 0x7FFFFFFF:
            break
        s = np.float32(0.0)
        for sub in q.top_list():
            s = np.float32(s + sub.score())
        docs.append(d)
        scores.append(s)
    return np.array(docs, np.int32), np.array(scores, np.float32)


def test_disi_priority_queue_sum_order_matches_python_model():
    rng = np.random.default_rng(41)
    dfs = [900, 800, 700, 650, 600, 500, 450, 400, 300, 250, 200, 120, 60, 0]
    seg, posts = helpers.build_segment(rng, 1500, dfs)     # dense lists: many docs carry 3+ clauses
    ix = helpers.oracle_index([seg])
    for terms in (list(range(10)), list(range(13)), [12, 3, 7, 0, 5, 9, 1, 11, 13, 2, 8, 4]):
        q, c = ob.make_queries([("bool", [(ob.SHOULD, t) for t in terms], 0)])
        d, s = _model(seg, posts[:], ix, terms)
        for k in (5, 200):
            hits, counts, total = ix.search_batch(q, c, k)
            want, _ = ob.topk_stream(d, s, k)
            assert total[0] == len(d)
            got = hits[0][:counts[0]]
            assert np.array_equal(got["doc"], want["doc"])
            assert np.array_equal(got["score"].view(np.uint32), want["score"].view(np.uint32))
    # the order does matter on this data: a plain clause-order sum differs in the last bits for some docs
    q, c = ob.make_queries([("bool", [(ob.SHOULD, t) for t in range(13)], 0)])
    d, s = _model(seg, posts, ix, list(range(13)))
    plain = np.zeros(1500, np.float32)
    for t in range(13):
        w, _i, _a, cache = ix.term_weight(t, 1.0)
        dd, ff = posts[t]
        plain[dd] = (plain[dd] + helpers.bm25_scores_numpy(w, 1.2, ff, seg.norms[dd], cache)).astype(np.float32)
    assert np.any(plain[d].view(np.uint32) != s.view(np.uint32))
