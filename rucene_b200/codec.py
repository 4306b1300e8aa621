"""ctypes binding of librucene_codec.so (include/rucene_codec.h): Lucene50 postings writer for a
DocsAndFreqs field, BM25 host math, synthetic Zipfian segments and block streams."""
import ctypes as C

import numpy as np

from . import _build


class TermState(C.Structure):
    """rg_term_state == BlockTermState (codec/postings/blocktree/mod.rs:33-59)."""
    _fields_ = [("doc_freq", C.c_int32), ("singleton_doc_id", C.c_int32),
                ("total_term_freq", C.c_int64), ("doc_start_fp", C.c_int64),
                ("skip_offset", C.c_int64)]


TERM_STATE_DTYPE = np.dtype([("doc_freq", "<i4"), ("singleton_doc_id", "<i4"),
                             ("total_term_freq", "<i8"), ("doc_start_fp", "<i8"),
                             ("skip_offset", "<i8")])


class SynthConfig(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("max_doc", C.c_int32), ("n_terms", C.c_uint32),
                ("doc_version", C.c_int32), ("n_threads", C.c_int32)]


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    L = C.CDLL(_build.build_codec())
    vp, u8p, i32p = C.c_void_p, C.POINTER(C.c_uint8), C.POINTER(C.c_int32)
    L.rc_last_error.restype = C.c_char_p
    L.rc_writer_create.restype = vp
    L.rc_writer_create.argtypes = [C.c_int, C.c_int32, u8p, C.c_char_p]
    L.rc_writer_add_term.argtypes = [vp, vp, vp, C.c_int32, C.POINTER(TermState)]
    L.rc_writer_finish.argtypes = [vp]
    L.rc_writer_set_ef.argtypes = [vp, C.c_int, C.c_int]
    L.rc_writer_block_counts.argtypes = [vp, vp]
    L.rc_writer_block_counts.restype = None
    L.rc_writer_data.restype = vp
    L.rc_writer_data.argtypes = [vp, C.POINTER(C.c_size_t)]
    L.rc_writer_forutil_table.argtypes = [vp, i32p]
    L.rc_writer_destroy.argtypes = [vp]
    L.rc_forutil_write_block.argtypes = [vp, C.c_int, vp]
    L.rc_float_to_byte315.restype = C.c_uint8
    L.rc_float_to_byte315.argtypes = [C.c_float]
    L.rc_byte315_to_float.restype = C.c_float
    L.rc_byte315_to_float.argtypes = [C.c_uint8]
    L.rc_encode_norm_value.restype = C.c_uint8
    L.rc_encode_norm_value.argtypes = [C.c_float, C.c_int32]
    L.rc_bm25_idf.restype = C.c_float
    L.rc_bm25_idf.argtypes = [C.c_int64, C.c_int64]
    L.rc_bm25_avg_field_length.restype = C.c_float
    L.rc_bm25_avg_field_length.argtypes = [C.c_int64, C.c_int64, C.c_int64]
    L.rc_bm25_norm_cache.argtypes = [C.c_float, C.c_float, C.c_float, vp]
    L.rc_synth_segment.restype = vp
    L.rc_synth_segment.argtypes = [C.POINTER(SynthConfig)]
    L.rc_segment_destroy.argtypes = [vp]
    L.rc_segment_doc_file.restype = vp
    L.rc_segment_doc_file.argtypes = [vp, C.POINTER(C.c_size_t)]
    L.rc_segment_norms.restype = vp
    L.rc_segment_norms.argtypes = [vp]
    L.rc_segment_terms.restype = vp
    L.rc_segment_terms.argtypes = [vp, C.POINTER(C.c_uint32)]
    L.rc_segment_stats.argtypes = [vp, C.POINTER(C.c_int64)]
    L.rc_segment_forutil_table.argtypes = [vp, i32p]
    L.rc_synth_blocks.restype = vp
    L.rc_synth_blocks.argtypes = [C.c_uint64, C.c_uint32, C.c_int, C.c_int, C.c_int]
    L.rc_blocks_stream.restype = vp
    L.rc_blocks_stream.argtypes = [vp, C.POINTER(C.c_size_t)]
    L.rc_blocks_offsets.restype = vp
    L.rc_blocks_offsets.argtypes = [vp, C.POINTER(C.c_uint32)]
    L.rc_blocks_values.restype = vp
    L.rc_blocks_values.argtypes = [vp]
    L.rc_blocks_forutil_table.argtypes = [i32p]
    L.rc_blocks_destroy.argtypes = [vp]
    L.rc_hardware_threads.restype = C.c_int
    _lib = L
    return L


def _err():
    return lib().rc_last_error().decode()


def _view(ptr, nbytes, dtype=np.uint8):
    """numpy view over native memory (no copy); the owner object must outlive it."""
    if nbytes == 0:
        return np.zeros(0, dtype=dtype)
    buf = (C.c_uint8 * nbytes).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype)


def forutil_table():
    t = (C.c_int32 * 32)()
    lib().rc_blocks_forutil_table(t)
    return np.array(t, dtype=np.int32)


class Segment:
    """One segment as the hot path sees it: `.doc` bytes, norms, per-term BlockTermState and the
    field statistics CollectionStatistics needs (search/searcher.rs:311-351)."""

    def __init__(self, doc_file, norms, terms, doc_count, sum_total_term_freq, sum_doc_freq,
                 max_doc, live_docs=None, owner=None, n_blocks=0):
        self.doc_file = doc_file          # np.uint8
        self.norms = norms                # np.uint8[max_doc] or None
        self.terms = terms                # np structured TERM_STATE_DTYPE
        self.doc_count = int(doc_count)
        self.sum_total_term_freq = int(sum_total_term_freq)
        self.sum_doc_freq = int(sum_doc_freq)
        self.max_doc = int(max_doc)
        self.live_docs = live_docs        # np.uint64 words or None
        self.n_blocks = int(n_blocks)
        self._owner = owner


class _NativeSegment:
    def __init__(self, h):
        self.h = h

    def __del__(self):
        if self.h:
            lib().rc_segment_destroy(self.h)
            self.h = None


def synth_segment(seed, max_doc, n_terms, doc_version=1, n_threads=0):
    L = lib()
    cfg = SynthConfig(seed, max_doc, n_terms, doc_version, n_threads)
    h = L.rc_synth_segment(C.byref(cfg))
    if not h:
        raise RuntimeError("rc_synth_segment: " + _err())
    owner = _NativeSegment(h)
    n = C.c_size_t()
    p = L.rc_segment_doc_file(h, C.byref(n))
    doc_file = _view(p, n.value)
    norms = _view(L.rc_segment_norms(h), max_doc)
    nt = C.c_uint32()
    tp = L.rc_segment_terms(h, C.byref(nt))
    terms = _view(tp, nt.value * TERM_STATE_DTYPE.itemsize, TERM_STATE_DTYPE)
    st = (C.c_int64 * 8)()
    L.rc_segment_stats(h, st)
    return Segment(doc_file, norms, terms, st[0], st[1], st[2], st[3], owner=owner, n_blocks=st[4])


class PostingsWriter:
    """Lucene50PostingsWriter for one DocsAndFreqs field (codec/postings/posting_writer.rs)."""

    def __init__(self, doc_version=1, max_doc=1 << 20, segment_id=None, suffix="", use_ef=False,
                 with_pf=True):
        """use_ef / with_pf: EfWriterMeta (posting_writer.rs:33-57) — the reference's EF / BITSET doc-block
        encodings, dormant in the open-source writer (off by default here too)."""
        L = lib()
        sid = (C.c_uint8 * 16)(*(segment_id or bytes(range(16))))
        self.h = L.rc_writer_create(doc_version, max_doc, sid, suffix.encode())
        if not self.h:
            raise RuntimeError("rc_writer_create: " + _err())
        if use_ef:
            L.rc_writer_set_ef(self.h, 1, 1 if with_pf else 0)
        self.max_doc = max_doc
        self.states = []
        self.sum_ttf = 0
        self.sum_df = 0

    def add_term(self, docs, freqs):
        docs = np.ascontiguousarray(docs, dtype=np.int32)
        freqs = np.ascontiguousarray(freqs, dtype=np.int32)
        if len(docs) == 0:
            self.states.append((0, -1, 0, 0, -1))
            return self.states[-1]
        st = TermState()
        rc = lib().rc_writer_add_term(self.h, docs.ctypes.data, freqs.ctypes.data, len(docs),
                                      C.byref(st))
        if rc != 0:
            raise ValueError("rc_writer_add_term: " + _err())
        t = (st.doc_freq, st.singleton_doc_id, st.total_term_freq, st.doc_start_fp, st.skip_offset)
        self.states.append(t)
        self.sum_ttf += st.total_term_freq
        self.sum_df += st.doc_freq
        return t

    def block_counts(self):
        """(full blocks written, of them EF, of them BITSET)"""
        out = np.zeros(3, np.uint64)
        lib().rc_writer_block_counts(self.h, out.ctypes.data)
        return tuple(int(x) for x in out)

    def finish(self, norms=None, doc_count=None, live_docs=None):
        L = lib()
        L.rc_writer_finish(self.h)
        n = C.c_size_t()
        p = L.rc_writer_data(self.h, C.byref(n))
        doc_file = _view(p, n.value).copy()
        terms = np.array(self.states, dtype=TERM_STATE_DTYPE)
        return Segment(doc_file, norms, terms, self.max_doc if doc_count is None else doc_count,
                       self.sum_ttf, self.sum_df, self.max_doc, live_docs=live_docs)

    def __del__(self):
        if getattr(self, "h", None):
            lib().rc_writer_destroy(self.h)
            self.h = None


def write_block(values, doc_version=1):
    """ForUtil::write_block for one 128-int block -> bytes."""
    v = np.ascontiguousarray(values, dtype=np.int32)
    assert v.shape == (128,)
    out = np.zeros(1 + 512, dtype=np.uint8)
    n = lib().rc_forutil_write_block(v.ctypes.data, doc_version, out.ctypes.data)
    return out[:n].copy()


class BlockStream:
    def __init__(self, stream, offsets, values, owner=None):
        self.stream, self.offsets, self.values, self._owner = stream, offsets, values, owner


class _NativeBlocks:
    def __init__(self, h):
        self.h = h

    def __del__(self):
        if self.h:
            lib().rc_blocks_destroy(self.h)
            self.h = None


def synth_blocks(seed, n_blocks, mode=0, param=0, doc_version=1):
    L = lib()
    h = L.rc_synth_blocks(seed, n_blocks, mode, param, doc_version)
    if not h:
        raise RuntimeError("rc_synth_blocks: " + _err())
    owner = _NativeBlocks(h)
    n = C.c_size_t()
    p = L.rc_blocks_stream(h, C.byref(n))
    stream = _view(p, n.value)
    nb = C.c_uint32()
    op = L.rc_blocks_offsets(h, C.byref(nb))
    offsets = _view(op, nb.value * 8, np.uint64)
    values = _view(L.rc_blocks_values(h), nb.value * 128 * 4, np.int32)
    return BlockStream(stream, offsets, values, owner)


def encode_norm_value(boost, field_length):
    return int(lib().rc_encode_norm_value(boost, field_length))


def bm25_idf(doc_freq, doc_count):
    return float(lib().rc_bm25_idf(doc_freq, doc_count))


def bm25_avg_field_length(sum_ttf, doc_count, max_doc):
    return float(lib().rc_bm25_avg_field_length(sum_ttf, doc_count, max_doc))


def bm25_norm_cache(k1, b, avgdl):
    out = np.zeros(256, dtype=np.float32)
    lib().rc_bm25_norm_cache(k1, b, avgdl, out.ctypes.data)
    return out
