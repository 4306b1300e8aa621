"""ctypes binding of oracle/liboracle.so — TEST INFRASTRUCTURE (see oracle/oracle.h).

Lives under tests/ on purpose: the product package (rucene_b200/) must never import the oracle.
bench.py's cpu_baseline / --impl reference legs and __graft_entry__.smoke() import this module
as the checker / the timed CPU reference only.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ODIR = os.path.join(ROOT, "oracle")

TERM_STATE_DTYPE = np.dtype([("doc_freq", "<i4"), ("singleton_doc_id", "<i4"),
                             ("total_term_freq", "<i8"), ("doc_start_fp", "<i8"),
                             ("skip_offset", "<i8")])
CLAUSE_DTYPE = np.dtype([("occur", "<i4"), ("term_id", "<u4"), ("boost", "<f4")])
QUERY_DTYPE = np.dtype([("clause_begin", "<u4"), ("n_clauses", "<u4"),
                        ("min_should_match", "<i4"), ("is_boolean", "<i4")])
HIT_DTYPE = np.dtype([("doc", "<i4"), ("score", "<f4")])

MUST, SHOULD, MUST_NOT, FILTER = 0, 1, 2, 3
NO_MORE_DOCS = 0x7FFFFFFF

_lib = None


def build():
    so = os.path.join(ODIR, "liboracle.so")
    srcs = [os.path.join(ODIR, "oracle.cpp"), os.path.join(ODIR, "oracle.h")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        p = subprocess.run(["make", "-C", ODIR], stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                           text=True)
        if p.returncode != 0:
            raise RuntimeError("oracle build failed:\n" + p.stdout)
    return so


def lib():
    global _lib
    if _lib is not None:
        return _lib
    L = C.CDLL(build())
    vp = C.c_void_p
    L.orc_last_error.restype = C.c_char_p
    L.orc_index_create.restype = vp
    L.orc_index_create.argtypes = [C.c_float, C.c_float]
    L.orc_index_destroy.argtypes = [vp]
    L.orc_index_add_segment.argtypes = [vp, vp, C.c_size_t, C.c_int32, vp, vp, vp, C.c_uint32,
                                        C.c_int64, C.c_int64, C.c_int64]
    L.orc_search_batch.argtypes = [vp, vp, C.c_uint32, vp, C.c_uint32, C.c_int, C.c_int, vp, vp, vp]
    L.orc_term_weight.argtypes = [vp, C.c_uint32, C.c_float, vp, vp, vp, vp]
    L.orc_postings.restype = C.c_int64
    L.orc_postings.argtypes = [vp, C.c_uint32, C.c_uint32, vp, vp, C.c_int64]
    L.orc_advance_seq.argtypes = [vp, C.c_uint32, C.c_uint32, vp, C.c_uint32, vp, vp]
    L.orc_forutil_decode.argtypes = [vp, C.c_size_t, vp, C.c_uint32, C.c_int, vp, vp, C.c_int]
    for f in ("orc_simd_pack", "orc_simd_unpack"):
        getattr(L, f).argtypes = [vp, vp, C.c_int]
        getattr(L, f).restype = None
    for f in ("orc_simd_delta_pack", "orc_simd_delta_unpack"):
        getattr(L, f).argtypes = [vp, vp, C.c_uint32, C.c_int]
        getattr(L, f).restype = None
    L.orc_simd_max_bits.argtypes = [vp]
    L.orc_packed_decode.argtypes = [C.c_int, C.c_int, vp, C.c_size_t, vp, C.c_int]
    L.orc_packed_encode.argtypes = [C.c_int, C.c_int, vp, vp, C.c_int]
    L.orc_packed_iterations.argtypes = [C.c_int, C.c_int]
    L.orc_packed_encoded_size.argtypes = [C.c_int, C.c_int]
    L.orc_fastest_format.argtypes = [C.c_int, C.c_float, C.POINTER(C.c_int)]
    L.orc_block_advance.argtypes = [vp, C.c_int32]
    L.orc_mock_conjunction.argtypes = [vp, vp, C.c_uint32, vp, vp, C.c_uint32]
    L.orc_mock_disjunction.argtypes = [vp, vp, C.c_uint32, C.c_int32, vp, vp, C.c_uint32]
    L.orc_mock_req_opt.argtypes = [vp, C.c_uint32, vp, C.c_uint32, vp, vp, C.c_uint32]
    L.orc_mock_req_not.argtypes = [vp, C.c_uint32, vp, C.c_uint32, vp, vp, C.c_uint32]
    L.orc_topk_stream.argtypes = [vp, vp, C.c_uint64, C.c_uint32, vp, vp, vp]
    L.orc_topk_merge.argtypes = [vp, vp, C.c_uint32, C.c_uint32, vp, vp]
    L.orc_float_to_byte315.restype = C.c_uint8
    L.orc_float_to_byte315.argtypes = [C.c_float]
    L.orc_byte315_to_float.restype = C.c_float
    L.orc_byte315_to_float.argtypes = [C.c_uint8]
    L.orc_norm_table.restype = C.c_float
    L.orc_norm_table.argtypes = [C.c_int]
    L.orc_bm25_idf.restype = C.c_float
    L.orc_bm25_idf.argtypes = [C.c_int64, C.c_int64]
    L.orc_bm25_avgdl.restype = C.c_float
    L.orc_bm25_avgdl.argtypes = [C.c_int64, C.c_int64, C.c_int64]
    L.orc_bm25_score.restype = C.c_float
    L.orc_bm25_score.argtypes = [C.c_float] * 4
    L.orc_bm25_cache.argtypes = [C.c_float, C.c_float, C.c_float, vp]
    L.orc_encode_norm.restype = C.c_uint8
    L.orc_encode_norm.argtypes = [C.c_float, C.c_int32]
    _lib = L
    return L


def err():
    return lib().orc_last_error().decode()


class OracleError(RuntimeError):
    pass


def _p(a):
    return None if a is None else a.ctypes.data


class Index:
    """The reference's IndexReader + DefaultIndexSearcher as the oracle models them."""

    def __init__(self, k1=1.2, b=0.75):
        self.h = lib().orc_index_create(k1, b)
        self._keep = []

    def add_segment(self, seg):
        """seg: rucene_b200.codec.Segment-like (doc_file, norms, terms, stats, max_doc, live_docs)."""
        terms = np.ascontiguousarray(seg.terms).astype(TERM_STATE_DTYPE, copy=False)
        doc_file = np.ascontiguousarray(seg.doc_file)
        norms = None if seg.norms is None else np.ascontiguousarray(seg.norms)
        live = None if seg.live_docs is None else np.ascontiguousarray(seg.live_docs, dtype=np.uint64)
        self._keep += [terms, doc_file, norms, live, seg]
        rc = lib().orc_index_add_segment(self.h, _p(doc_file), doc_file.size, seg.max_doc, _p(norms),
                                         _p(live), _p(terms), len(terms), seg.doc_count,
                                         seg.sum_total_term_freq, seg.sum_doc_freq)
        if rc != 0:
            raise OracleError(err())

    def search_batch(self, queries, clauses, k, parallel_mode=0, n_threads=1):
        q = np.ascontiguousarray(queries, dtype=QUERY_DTYPE)
        c = np.ascontiguousarray(clauses, dtype=CLAUSE_DTYPE)
        hits = np.zeros((len(q), k), dtype=HIT_DTYPE)
        counts = np.zeros(len(q), dtype=np.uint32)
        total = np.zeros(len(q), dtype=np.uint64)
        rc = lib().orc_search_batch(self.h, _p(q), len(q), _p(c), k, parallel_mode, n_threads,
                                    _p(hits), _p(counts), _p(total))
        if rc != 0:
            raise OracleError(err())
        return hits, counts, total

    def leaf_records(self, seg, queries, clauses, k, n_threads=1):
        q = np.ascontiguousarray(queries, dtype=QUERY_DTYPE)
        c = np.ascontiguousarray(clauses, dtype=CLAUSE_DTYPE)
        out = np.zeros((len(q), 16 + 8 * k), dtype=np.uint8)
        lib().orc_search_leaf_records.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32,
                                                  C.c_void_p, C.c_uint32, C.c_int, C.c_void_p]
        rc = lib().orc_search_leaf_records(self.h, seg, _p(q), len(q), _p(c), k, n_threads, _p(out))
        if rc != 0:
            raise OracleError(err())
        return out

    def term_weight(self, term_id, boost=1.0):
        w, idf, avgdl = C.c_float(), C.c_float(), C.c_float()
        cache = np.zeros(256, dtype=np.float32)
        rc = lib().orc_term_weight(self.h, term_id, boost, C.byref(w), C.byref(idf), C.byref(avgdl),
                                   _p(cache))
        if rc != 0:
            raise OracleError(err())
        return np.float32(w.value), np.float32(idf.value), np.float32(avgdl.value), cache

    def postings(self, seg, term_id, cap):
        docs = np.zeros(cap, dtype=np.int32)
        freqs = np.zeros(cap, dtype=np.int32)
        n = lib().orc_postings(self.h, seg, term_id, _p(docs), _p(freqs), cap)
        if n < 0:
            raise OracleError(err())
        return docs[:n], freqs[:n]

    def advance_seq(self, seg, term_id, targets):
        t = np.ascontiguousarray(targets, dtype=np.int32)
        docs = np.zeros(len(t), dtype=np.int32)
        freqs = np.zeros(len(t), dtype=np.int32)
        rc = lib().orc_advance_seq(self.h, seg, term_id, _p(t), len(t), _p(docs), _p(freqs))
        if rc != 0:
            raise OracleError(err())
        return docs, freqs

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_index_destroy(self.h)
            self.h = None


def make_queries(specs):
    """specs: list of ("term", term_id[, boost]), ("bool", [(occur, term_id[, boost])...], msm) or
    ("dismax", [(term_id[, boost])...], tie_breaker_multiplier).
    Returns (queries, clauses) structured arrays for Index.search_batch."""
    qs, cs = [], []
    for s in specs:
        if s[0] == "term":
            boost = s[2] if len(s) > 2 else 1.0
            qs.append((len(cs), 1, 0, 0))
            cs.append((SHOULD, s[1], boost))
        elif s[0] == "dismax":  # ("dismax", [(term_id[, boost])...], tie_breaker_multiplier)
            begin = len(cs)
            for cl in s[1]:
                cs.append((SHOULD, cl[0], cl[1] if len(cl) > 1 else 1.0))
            tie_bits = int(np.array([s[2]], np.float32).view(np.int32)[0])
            qs.append((begin, len(s[1]), tie_bits, 2))
        else:
            begin = len(cs)
            for cl in s[1]:
                cs.append((cl[0], cl[1], cl[2] if len(cl) > 2 else 1.0))
            qs.append((begin, len(s[1]), s[2] if len(s) > 2 else 0, 1))
    return np.array(qs, dtype=QUERY_DTYPE), np.array(cs, dtype=CLAUSE_DTYPE)


def forutil_decode(stream, offsets, doc_version, table, n_threads=1):
    stream = np.ascontiguousarray(stream, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    table = np.ascontiguousarray(table, dtype=np.int32)
    out = np.zeros((len(offsets), 128), dtype=np.int32)
    rc = lib().orc_forutil_decode(_p(stream), stream.size, _p(offsets), len(offsets), doc_version,
                                  _p(table), _p(out), n_threads)
    if rc != 0:
        raise OracleError(err())
    return out


def simd_pack(data, bits):
    d = np.ascontiguousarray(data, dtype=np.uint32)
    enc = np.zeros(512, dtype=np.uint8)
    lib().orc_simd_pack(_p(d), _p(enc), bits)
    return enc


def simd_unpack(enc, bits):
    e = np.ascontiguousarray(enc, dtype=np.uint8)
    assert e.size >= 16 * bits
    out = np.zeros(128, dtype=np.uint32)
    lib().orc_simd_unpack(_p(e), _p(out), bits)
    return out


def simd_delta_pack(data, base, bits):
    d = np.ascontiguousarray(data, dtype=np.uint32)
    enc = np.zeros(512, dtype=np.uint8)
    lib().orc_simd_delta_pack(_p(d), _p(enc), base, bits)
    return enc


def simd_delta_unpack(enc, base, bits):
    e = np.ascontiguousarray(enc, dtype=np.uint8)
    out = np.zeros(128, dtype=np.uint32)
    lib().orc_simd_delta_unpack(_p(e), _p(out), base, bits)
    return out


def packed_decode(format_id, bpv, blocks, iterations, cap=256):
    b = np.ascontiguousarray(blocks, dtype=np.uint8)
    out = np.zeros(max(cap, 256), dtype=np.int32)
    n = lib().orc_packed_decode(format_id, bpv, _p(b), b.size, _p(out), iterations)
    if n < 0:
        raise OracleError(err())
    return out[:n]


def packed_encode(format_id, bpv, values, iterations):
    v = np.ascontiguousarray(values, dtype=np.int32)
    out = np.zeros(1024, dtype=np.uint8)
    n = lib().orc_packed_encode(format_id, bpv, _p(v), _p(out), iterations)
    if n < 0:
        raise OracleError(err())
    return out[:n]


def _mock_lists(lists):
    flat = np.concatenate([np.asarray(x, dtype=np.int32) for x in lists]) if lists else np.zeros(0, np.int32)
    lens = np.array([len(x) for x in lists], dtype=np.uint32)
    return np.ascontiguousarray(flat), lens


def mock_conjunction(lists, cap=1 << 16):
    flat, lens = _mock_lists(lists)
    docs = np.zeros(cap, np.int32)
    scores = np.zeros(cap, np.float32)
    n = lib().orc_mock_conjunction(_p(flat), _p(lens), len(lens), _p(docs), _p(scores), cap)
    if n < 0:
        raise OracleError(err())
    return docs[:n], scores[:n]


def mock_disjunction(lists, min_should_match=1, cap=1 << 16):
    flat, lens = _mock_lists(lists)
    docs = np.zeros(cap, np.int32)
    scores = np.zeros(cap, np.float32)
    n = lib().orc_mock_disjunction(_p(flat), _p(lens), len(lens), min_should_match, _p(docs),
                                   _p(scores), cap)
    if n < 0:
        raise OracleError(err())
    return docs[:n], scores[:n]


def mock_req_opt(req, opt, cap=1 << 16):
    r = np.ascontiguousarray(req, dtype=np.int32)
    o = np.ascontiguousarray(opt, dtype=np.int32)
    docs = np.zeros(cap, np.int32)
    scores = np.zeros(cap, np.float32)
    n = lib().orc_mock_req_opt(_p(r), len(r), _p(o), len(o), _p(docs), _p(scores), cap)
    if n < 0:
        raise OracleError(err())
    return docs[:n], scores[:n]


def mock_req_not(req, nots, cap=1 << 16):
    r = np.ascontiguousarray(req, dtype=np.int32)
    o = np.ascontiguousarray(nots, dtype=np.int32)
    docs = np.zeros(cap, np.int32)
    scores = np.zeros(cap, np.float32)
    n = lib().orc_mock_req_not(_p(r), len(r), _p(o), len(o), _p(docs), _p(scores), cap)
    if n < 0:
        raise OracleError(err())
    return docs[:n], scores[:n]


def topk_stream(docs, scores, k):
    d = np.ascontiguousarray(docs, dtype=np.int32)
    s = np.ascontiguousarray(scores, dtype=np.float32)
    out = np.zeros(k, HIT_DTYPE)
    heap = np.zeros(k, HIT_DTYPE)
    cnt = C.c_uint32()
    rc = lib().orc_topk_stream(_p(d), _p(s), len(d), k, _p(out), _p(heap), C.byref(cnt))
    if rc != 0:
        raise OracleError(err())
    return out[:cnt.value], heap[:min(k, len(d))]


def topk_merge(leaf_hits, k):
    flat = np.concatenate([np.asarray(h, dtype=HIT_DTYPE) for h in leaf_hits])
    counts = np.array([len(h) for h in leaf_hits], dtype=np.uint32)
    out = np.zeros(k, HIT_DTYPE)
    cnt = C.c_uint32()
    rc = lib().orc_topk_merge(_p(np.ascontiguousarray(flat)), _p(counts), len(counts), k, _p(out),
                              C.byref(cnt))
    if rc != 0:
        raise OracleError(err())
    return out[:cnt.value]


def mock_run(spec, ops):
    """spec: prefix-coded scorer tree (see oracle.cpp parse_mock); ops: [(op, target)...] with
    op 0=next, 1=advance(target), 2=score only.  Returns (doc_ids, scores)."""
    s = np.ascontiguousarray(spec, dtype=np.int32)
    o = np.ascontiguousarray(ops, dtype=np.int32).reshape(-1, 2)
    lib().orc_mock_run.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p,
                                   C.c_void_p]
    docs = np.zeros(len(o), np.int32)
    scores = np.zeros(len(o), np.float32)
    rc = lib().orc_mock_run(_p(s), len(s), _p(o), len(o), _p(docs), _p(scores))
    if rc != 0:
        raise OracleError(err())
    return docs, scores


def leaf(docs):
    return [0, len(docs)] + list(docs)


def conj(*children):
    out = [1, len(children)]
    for c in children:
        out += c
    return out


def disj(msm, *children):
    out = [2, msm, len(children)]
    for c in children:
        out += c
    return out


def req_opt(a, b):
    return [3] + a + b


def req_not(a, b):
    return [4] + a + b
