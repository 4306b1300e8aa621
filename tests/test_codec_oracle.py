"""Host logic on CPU: the product's write side (codec.cpp) against the oracle's read side,
BM25 host math against the oracle's restatement, synthetic segments, and the oracle's search
against a brute-force numpy model."""
import numpy as np
import pytest

import helpers
import oracle_binding as ob
from rucene_b200 import codec

DFS = [0, 1, 2, 5, 127, 128, 129, 255, 256, 257, 300, 1023, 1024, 1025, 4000, 8193, 20000]


@pytest.mark.parametrize("version", [0, 1])
def test_writer_reader_round_trip(version):
    rng = np.random.default_rng(100 + version)
    seg, postings = helpers.build_segment(rng, 60000, DFS, doc_version=version, dense_terms=(13, 14))
    ix = helpers.oracle_index([seg])
    for t, (docs, freqs) in enumerate(postings):
        st = seg.terms[t]
        assert st["doc_freq"] == len(docs)
        if len(docs) == 0:
            continue
        assert st["total_term_freq"] == int(freqs.sum())
        assert (st["skip_offset"] != -1) == (len(docs) > 128)
        assert (st["singleton_doc_id"] != -1) == (len(docs) == 1)
        got_docs, got_freqs = ix.postings(0, t, len(docs) + 5)
        assert np.array_equal(got_docs, docs)
        assert np.array_equal(got_freqs, freqs)


@pytest.mark.parametrize("version", [0, 1])
def test_advance_uses_skip_data_correctly(version):
    """BlockDocIterator::advance + Lucene50SkipReader land on the first doc >= target."""
    rng = np.random.default_rng(200 + version)
    max_doc = 3_000_000
    dfs = [129, 1000, 1024, 1025, 9000, 70000, 131072 + 77]  # up to 3 skip levels (8^2*128=8192)
    seg, postings = helpers.build_segment(rng, max_doc, dfs, doc_version=version)
    ix = helpers.oracle_index([seg])
    for t, (docs, freqs) in enumerate(postings):
        for trial in range(3):
            n = [5, 60, 700][trial]
            targets = np.sort(rng.choice(max_doc + 1000, size=n, replace=False)).astype(np.int32)
            # advance() is only defined for strictly increasing targets beyond the current doc
            out_docs, out_freqs = ix.advance_seq(0, t, targets)
            cur = -1
            for tg, d, f in zip(targets, out_docs, out_freqs):
                if tg <= cur:
                    break  # undefined territory in the reference's contract
                i = np.searchsorted(docs, tg, side="left")
                if i >= len(docs):
                    assert d == ob.NO_MORE_DOCS
                    break
                assert d == docs[i], (t, tg)
                assert f == freqs[i]
                cur = d


def test_doc_file_header_and_footer():
    rng = np.random.default_rng(5)
    seg, _ = helpers.build_segment(rng, 5000, [10, 500])
    raw = seg.doc_file.tobytes()
    assert raw[:4] == bytes([0x3F, 0xD7, 0x6C, 0x17])           # codec_util.rs:30 CODEC_MAGIC
    assert raw[4] == 25 and raw[5:30] == b"Lucene50PostingsWriterDoc"
    assert raw[30:34] == bytes([0, 0, 0, 1])                      # VERSION_CURRENT = 1
    hdr = 9 + 25 + 16 + 1                                         # index_header_length
    assert raw[hdr] == 2                                          # PackedInts VERSION_CURRENT
    codes = list(raw[hdr + 1:hdr + 33])
    assert codes == [(1 << 5) | (b - 1) if b in (1, 2, 4) else (b - 1) for b in range(1, 33)]
    assert seg.terms[0]["doc_start_fp"] == hdr + 33
    import zlib
    assert raw[-16:-12] == bytes([0xC0, 0x28, 0x93, 0xE8]) and raw[-12:-8] == bytes(4)
    assert int.from_bytes(raw[-8:], "big") == zlib.crc32(raw[:-8])


def test_block_byte_sizes():
    """1 + 16*b bytes for every packed block in both versions; all-equal = code 0 + vint."""
    rng = np.random.default_rng(6)
    for version in (0, 1):
        for b in range(1, 33):
            vals = rng.integers(0, 1 << b, 128, dtype=np.uint64).astype(np.uint32)
            vals[3] |= 1 << (b - 1)
            vals[4] = vals[3] ^ 1
            enc = codec.write_block(vals.astype(np.int32), version)
            assert enc[0] == b and len(enc) == 1 + 16 * b
            dec = ob.forutil_decode(np.concatenate([enc, np.zeros(64, np.uint8)]), [0], version,
                                    codec.forutil_table())
            assert np.array_equal(dec[0].astype(np.uint32), vals)
        enc = codec.write_block(np.full(128, 300, np.int32), version)
        assert enc.tolist() == [0, 0xAC, 0x02]
    # SIMD layout written by the product == the oracle's restatement of pack_bits!
    for b in range(1, 33):
        vals = rng.integers(0, 1 << b, 128, dtype=np.uint64).astype(np.uint32)
        vals[0] |= 1 << (b - 1)
        vals[1] = vals[0] ^ 1
        assert np.array_equal(codec.write_block(vals.astype(np.int32), 1)[1:], ob.simd_pack(vals, b)[:16 * b])


def test_bm25_host_math_matches_oracle():
    L = ob.lib()
    for df, dc in [(1, 11), (1, 32), (17, 100000), (0, 5), (49_999_999, 100_000_000)]:
        assert np.float32(codec.bm25_idf(df, dc)) == np.float32(L.orc_bm25_idf(df, dc))
    for args in [(0, 5, 11), (8, 2, 3), (9, -1, 3), (2_700_000_123, 100_000_000, 100_000_000)]:
        assert codec.bm25_avg_field_length(*args) == L.orc_bm25_avgdl(*args)
    for k1, b, avgdl in [(1.2, 0.75, 27.3), (0.9, 0.4, 200.0), (2.0, 1.0, 1.0)]:
        want = np.zeros(256, np.float32)
        L.orc_bm25_cache(k1, b, avgdl, want.ctypes.data)
        assert np.array_equal(codec.bm25_norm_cache(k1, b, avgdl).view(np.uint32), want.view(np.uint32))
    for ln in (1, 2, 120, 1000, 10000, 2**31 - 1):
        assert codec.encode_norm_value(1.0, ln) == L.orc_encode_norm(1.0, ln)


def test_synth_segment_is_deterministic_and_zipfian():
    a = codec.synth_segment(0x5EED0001, 20000, 3000, doc_version=1, n_threads=1)
    b = codec.synth_segment(0x5EED0001, 20000, 3000, doc_version=1, n_threads=4)
    assert a.doc_file.tobytes() == b.doc_file.tobytes()
    assert a.norms.tobytes() == b.norms.tobytes()
    assert a.terms.tobytes() == b.terms.tobytes()
    df = a.terms["doc_freq"]
    target = np.maximum(1, 20000 // (np.arange(3000) + 2))
    big = target >= 200
    assert np.all(np.abs(df[big] - target[big]) < 6 * np.sqrt(target[big]))
    assert a.sum_doc_freq == int(df.sum()) and a.doc_count == 20000
    ix = helpers.oracle_index([a])
    tot = 0
    for t in range(0, 3000, 37):
        docs, freqs = ix.postings(0, t, int(df[t]) + 1)
        assert len(docs) == df[t]
        if len(docs):
            assert np.all(np.diff(docs) > 0) and docs[-1] < 20000 and freqs.min() >= 1
            assert int(freqs.sum()) == a.terms["total_term_freq"][t]
        tot += len(docs)
    assert tot > 0
    hist = np.bincount(a.norms, minlength=256)
    assert (hist > 0).sum() >= 8  # several distinct norm classes


def _brute_force(seg_list, postings_list, ix, spec, k, stats=None):
    """Independent numpy model: returns the (doc,score) stream in collection order."""
    stream_d, stream_s = [], []
    doc_base = 0
    for seg, postings in zip(seg_list, postings_list):
        kind = spec[0]
        if kind == "dismax":
            clauses = [(ob.SHOULD,) + tuple(cl) for cl in spec[1]]
        else:
            clauses = [(ob.SHOULD, spec[1], 1.0)] if kind == "term" else spec[1]
        per = []
        for occ, t, *rest in clauses:
            boost = rest[0] if rest else 1.0
            w, _idf, _avgdl, cache = ix.term_weight(t, boost)
            docs, freqs = postings[t]
            sc = helpers.bm25_scores_numpy(w, 1.2, freqs, seg.norms[docs], cache)
            per.append((occ, docs, sc))
        musts = [p for p in per if p[0] == ob.MUST]
        shoulds = [p for p in per if p[0] == ob.SHOULD]
        if musts:
            if any(len(p[1]) == 0 for p in musts):
                doc_base += seg.max_doc
                continue
            order = sorted(range(len(musts)), key=lambda i: len(musts[i][1]))  # stable by cost
            docs = musts[order[0]][1]
            for i in order[1:]:
                docs = np.intersect1d(docs, musts[i][1])
            score = np.zeros(len(docs), np.float32)
            first = True
            for i in order:
                d, s = musts[i][1], musts[i][2]
                part = s[np.searchsorted(d, docs)]
                score = part.copy() if first else (score + part).astype(np.float32)
                first = False
        else:
            shoulds = [p for p in shoulds if len(p[1])]
            if not shoulds:
                doc_base += seg.max_doc
                continue
            docs = np.unique(np.concatenate([p[1] for p in shoulds]))
            score = np.zeros(len(docs), np.float32)
            count = np.zeros(len(docs), np.int32)
            for _occ, d, s in shoulds:  # clause order, starting from 0.0f
                pos = np.searchsorted(docs, d)
                score[pos] = (score[pos] + s).astype(np.float32)
                count[pos] += 1
            msm = spec[2] if kind == "bool" and len(per) > 1 else 0
            if msm > 1:  # disjunction_scorer.rs:317-329
                docs, score = docs[count >= msm], score[count >= msm]
            if kind == "dismax" and len(shoulds) > 1:  # score_max, disjunction_scorer.rs:241-263
                mx = np.full(len(docs), -np.inf, np.float32)
                for _occ, d, sc in shoulds:
                    pos = np.searchsorted(docs, d)
                    mx[pos] = np.maximum(mx[pos], sc)
                tie = np.float32(spec[2])
                score = (mx + ((score - mx).astype(np.float32) * tie).astype(np.float32)).astype(np.float32)
        if seg.live_docs is not None:
            live = (seg.live_docs[docs >> 6] >> (docs & 63).astype(np.uint64)) & np.uint64(1)
            docs, score = docs[live == 1], score[live == 1]
        shoulds = [p for p in shoulds if len(p[1])]
        if musts and shoulds:
            # ReqOptScorer (req_opt_scorer.rs:43-64): sequential (scores_sum, scores_num) chain over
            # the collected docs of this leaf; optional side = clause-order sum from 0.0f
            opt_docs = np.unique(np.concatenate([p[1] for p in shoulds]))
            opt = np.zeros(len(opt_docs), np.float32)
            for _occ, d, s in shoulds:
                pos = np.searchsorted(opt_docs, d)
                opt[pos] = (opt[pos] + s).astype(np.float32)
            pos = np.minimum(np.searchsorted(opt_docs, docs), len(opt_docs) - 1)
            has_opt = opt_docs[pos] == docs
            ssum, snum = np.float32(0.0), 0
            for i in range(len(docs)):
                req = score[i]
                if snum > 100 and np.float32(2.0) * req < ssum / np.float32(snum):
                    if stats is not None:
                        stats["skipped"] = stats.get("skipped", 0) + 1
                    continue
                ssum = np.float32(ssum + req)
                snum += 1
                if has_opt[i]:
                    score[i] = np.float32(req + opt[pos[i]])
        stream_d.append(docs + doc_base)
        stream_s.append(score)
        doc_base += seg.max_doc
    if not stream_d:
        return np.zeros(0, np.int32), np.zeros(0, np.float32)
    return np.concatenate(stream_d).astype(np.int32), np.concatenate(stream_s)


@pytest.mark.parametrize("version,live,ef", [(1, None, 0), (0, None, 0), (1, 0.7, 0), (1, None, 1), (0, 0.8, 2)])
def test_oracle_search_matches_brute_force(version, live, ef):
    rng = np.random.default_rng(300 + version)
    dfs = [0, 1, 3, 100, 128, 129, 500, 2000, 9000, 30000, 45000]
    segs, posts = [], []
    for s in range(2):
        cnt = []
        seg, p = helpers.build_segment(rng, 50000 + 1000 * s, dfs, doc_version=version, live_fraction=live,
                                       use_ef=ef > 0, with_pf=ef != 2, counts=cnt)
        assert (cnt[0][1] + cnt[0][2] > 0) == (ef > 0)
        segs.append(seg)
        posts.append(p)
    ix = helpers.oracle_index(segs)
    specs = [("term", 7), ("term", 1), ("term", 0), ("term", 10),
             ("bool", [(ob.MUST, 9), (ob.MUST, 10)], 0),
             ("bool", [(ob.MUST, 10), (ob.MUST, 6), (ob.MUST, 8)], 0),
             ("bool", [(ob.MUST, 10), (ob.MUST, 0)], 0),
             ("bool", [(ob.SHOULD, 3), (ob.SHOULD, 9), (ob.SHOULD, 7)], 0),
             ("bool", [(ob.SHOULD, 10), (ob.SHOULD, 9), (ob.SHOULD, 8), (ob.SHOULD, 1), (ob.SHOULD, 0)], 0),
             ("bool", [(ob.SHOULD, 5, 2.5)], 0)]
    q, c = ob.make_queries(specs)
    for k in (1, 10, 100):
        hits, counts, total = ix.search_batch(q, c, k)
        for i, spec in enumerate(specs):
            d, s = _brute_force(segs, posts, ix, spec, k)
            want, _ = ob.topk_stream(d, s, k)
            assert total[i] == len(d), spec
            assert counts[i] == len(want)
            got = hits[i][:counts[i]]
            assert np.array_equal(got["doc"], want["doc"]), spec
            assert np.array_equal(got["score"].view(np.uint32), want["score"].view(np.uint32)), spec


@pytest.mark.parametrize("live", [None, 0.8])
def test_oracle_req_opt_matches_brute_force(live):
    """MUST + SHOULD in one query: ReqOptScorer's running-mean skip (req_opt_scorer.rs:19,46-53) as
    a plain numpy chain — and the skip does fire on this data."""
    rng = np.random.default_rng(410)
    dfs = [0, 2, 90, 700, 5000, 14000, 26000, 33000]
    segs, posts = [], []
    for s in range(2):
        seg, p = helpers.build_segment(rng, 36000 + 500 * s, dfs, live_fraction=live)
        segs.append(seg)
        posts.append(p)
    ix = helpers.oracle_index(segs)
    specs = [("bool", [(ob.MUST, 7), (ob.SHOULD, 6)], 0),
             ("bool", [(ob.MUST, 6), (ob.MUST, 7), (ob.SHOULD, 5), (ob.SHOULD, 3)], 0),
             ("bool", [(ob.SHOULD, 4), (ob.MUST, 5), (ob.SHOULD, 7), (ob.SHOULD, 2)], 0),
             ("bool", [(ob.MUST, 5), (ob.SHOULD, 0)], 0),      # SHOULD absent everywhere: plain MUST
             ("bool", [(ob.MUST, 0), (ob.SHOULD, 5)], 0),      # MUST absent: no scorer
             ("bool", [(ob.MUST, 3), (ob.SHOULD, 7, 3.0), (ob.SHOULD, 1)], 1)]
    q, c = ob.make_queries(specs)
    stats = {}
    for k in (10, 100):
        hits, counts, total = ix.search_batch(q, c, k)
        for i, spec in enumerate(specs):
            d, s = _brute_force(segs, posts, ix, spec, k, stats)
            want, _ = ob.topk_stream(d, s, k)
            assert total[i] == len(d), spec
            got = hits[i][:counts[i]]
            assert np.array_equal(got["doc"], want["doc"]), spec
            assert np.array_equal(got["score"].view(np.uint32), want["score"].view(np.uint32)), spec
    assert stats.get("skipped", 0) > 100


def test_oracle_min_should_match_matches_brute_force():
    rng = np.random.default_rng(420)
    dfs = [0, 2, 90, 700, 5000, 14000, 26000]
    segs, posts = [], []
    for s in range(2):
        seg, p = helpers.build_segment(rng, 30000 + 500 * s, dfs, live_fraction=0.9 if s else None)
        segs.append(seg)
        posts.append(p)
    ix = helpers.oracle_index(segs)
    specs = [("bool", [(ob.SHOULD, 6), (ob.SHOULD, 5)], 2),
             ("bool", [(ob.SHOULD, 6), (ob.SHOULD, 5), (ob.SHOULD, 4), (ob.SHOULD, 3)], 3),
             ("bool", [(ob.SHOULD, 4), (ob.SHOULD, 0), (ob.SHOULD, 6)], 2),
             ("bool", [(ob.SHOULD, 3), (ob.SHOULD, 2)], 3)]
    q, c = ob.make_queries(specs)
    hits, counts, total = ix.search_batch(q, c, 20)
    for i, spec in enumerate(specs):
        d, s = _brute_force(segs, posts, ix, spec, 20)
        want, _ = ob.topk_stream(d, s, 20)
        assert total[i] == len(d), spec
        got = hits[i][:counts[i]]
        assert np.array_equal(got["doc"], want["doc"]), spec
        assert np.array_equal(got["score"].view(np.uint32), want["score"].view(np.uint32)), spec
    assert total[3] == 0 and 0 < total[1] < total[0]


def test_elias_fano_encoder_reference_vectors():
    """The reference's own unit vectors, util/packed/elias_fano_encoder.rs:398-448."""
    import ctypes as C
    L = codec.lib()
    L.rc_ef_num_longs_for_bits.restype = C.c_int64
    L.rc_ef_num_longs_for_bits.argtypes = [C.c_int64]
    for n, want in ((5, 1), (31, 1), (32, 1), (33, 1), (65, 2), (128, 2), (129, 3)):   # :399-407
        assert L.rc_ef_num_longs_for_bits(n) == want
    L.rc_ef_pack_value.argtypes = [C.c_int64, C.c_void_p, C.c_int, C.c_int, C.c_int64]
    L.rc_ef_pack_value.restype = None
    lv = np.zeros(2, np.int64)                                                          # :414-425
    L.rc_ef_pack_value(2, lv.ctypes.data, 2, 2, 31)
    assert lv[0] == np.int64(-0x8000000000000000)
    lv[:] = 0
    L.rc_ef_pack_value(0b11111, lv.ctypes.data, 2, 5, 12)
    assert lv.view(np.uint64)[0] == 0xF000000000000000 and lv[1] == 1
    # :427-447 encode_upper: EliasFanoEncoder::new(7, 24, 256), high values 0,0,1,1,2 -> 1,3,11,27,91
    L.rc_ef_encode.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int, C.c_void_p]
    vals = np.array([0, 1, 2, 3, 4, 24, 24], np.int64)
    out = np.zeros(16, np.int64)
    geom = np.zeros(4, np.int32)
    assert L.rc_ef_encode(vals.ctypes.data, 7, 24, out.ctypes.data, 16, geom.ctypes.data) == 0
    assert geom[0] == 1                       # floor(log2(24 / 7)) — "different from lucene version" (:72-77)
    assert int(out[0]) & 0x7f == 91
    assert int(out[0]) == 91 | (1 << (5 + 12)) | (1 << (6 + 12))
    # :408-412 get_encoder(128, 510901): geometry of a block-sized encoder
    vals = np.sort(np.random.default_rng(3).choice(510901, 128, replace=False)).astype(np.int64)
    vals[-1] = 510901
    out = np.zeros(64, np.int64)
    assert L.rc_ef_encode(vals.ctypes.data, 128, 510901, out.ctypes.data, 64, geom.ctypes.data) == 0
    assert list(geom) == [11, 6, 22, 0]       # 510901/128 = 3991 -> 11 low bits; (249 + 128) bits -> 6 longs


def test_oracle_disjunction_max_matches_brute_force():
    """DisjunctionMaxQuery over TermQuerys (query/disjunction_max_query.rs:51-155, scorer
    disjunction_scorer.rs:106-186,241-263): max + (sum - max) * tie_breaker."""
    rng = np.random.default_rng(430)
    dfs = [0, 2, 90, 700, 5000, 14000, 26000]
    segs, posts = [], []
    for s in range(2):
        d = list(dfs)
        if s == 1:
            d[5] = 0
        seg, p = helpers.build_segment(rng, 30000 + 500 * s, d, live_fraction=0.9 if s else None)
        segs.append(seg)
        posts.append(p)
    ix = helpers.oracle_index(segs)
    specs = [("dismax", [(6,), (5,)], 0.0), ("dismax", [(6,), (5,), (4, 2.0), (3,)], 0.3),
             ("dismax", [(4,), (0,), (6,)], 1.0), ("dismax", [(3,)], 0.5), ("dismax", [(5,), (2,)], 0.1)]
    q, c = ob.make_queries(specs)
    hits, counts, total = ix.search_batch(q, c, 20)
    for i, spec in enumerate(specs):
        d, s = _brute_force(segs, posts, ix, spec, 20)
        want, _ = ob.topk_stream(d, s, 20)
        assert total[i] == len(d), spec
        got = hits[i][:counts[i]]
        assert np.array_equal(got["doc"], want["doc"]), spec
        assert np.array_equal(got["score"].view(np.uint32), want["score"].view(np.uint32)), spec
