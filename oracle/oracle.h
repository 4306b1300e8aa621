/*
 * oracle.h — C API of the CPU oracle.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a CPU restatement of zhihu/rucene's
 * IndexSearcher hot path (ForUtil block decode -> BlockDocIterator ->
 * Conjunction/Disjunction scorers -> BM25 -> TopDocsCollector).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load it.  The product (rucene_b200/, librucene_gpu.so) never links,
 * imports or executes anything under oracle/.
 *
 * Parity status: the reference cannot be compiled here (Rust nightly-2020-03-12
 * + un-vendored crates, no rustc/cargo).  The restatement is pinned against
 * every known-answer vector the reference's own tests hold for this path
 * (tests/test_oracle_kat.py lists them with file:line); whole-block byte
 * streams, skip data and end-to-end BM25 TopDocs are NOT pinned by any vector
 * in the reference ("parity unpinned" for those).
 */
#ifndef RUCENE_ORACLE_H
#define RUCENE_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* Per-term handle of one segment = BlockTermState
 * (reference: src/core/codec/postings/blocktree/mod.rs:33-59). */
typedef struct {
    int32_t doc_freq;        /* 0 = term absent from this segment */
    int32_t singleton_doc_id;/* docid when doc_freq==1, else -1 */
    int64_t total_term_freq;
    int64_t doc_start_fp;    /* into the .doc file */
    int64_t skip_offset;     /* relative to doc_start_fp, -1 if doc_freq<=128 */
} orc_term_state;

enum { ORC_MUST = 0, ORC_SHOULD = 1, ORC_MUST_NOT = 2, ORC_FILTER = 3 };

typedef struct {
    int32_t occur;
    uint32_t term_id;
    float boost;
} orc_clause;

typedef struct {
    uint32_t clause_begin;
    uint32_t n_clauses;
    int32_t min_should_match; /* as passed to BooleanQuery::build; kind 2: the bits of the f32
                                 tie_breaker_multiplier instead */
    int32_t is_boolean;       /* 0: bare TermQuery (n_clauses==1), 1: BooleanQuery::build,
                                 2: DisjunctionMaxQuery::build over the clauses' TermQuerys */
} orc_query;

typedef struct {
    int32_t doc;
    float score;
} orc_hit;

typedef struct orc_index orc_index;

const char* orc_last_error(void);

/* ---- index ---- */
orc_index* orc_index_create(float k1, float b);
void orc_index_destroy(orc_index*);
/* doc_file: whole .doc file (IndexHeader + ForUtil table + postings + footer).
 * The oracle keeps pointers to doc_file/norms/live_docs: caller keeps them alive.
 * force_scalar: decode version-1 files as if SSE3 were absent is NOT a thing in
 * the reference (layout is fixed by version), so there is no such switch. */
int orc_index_add_segment(orc_index*, const uint8_t* doc_file, size_t doc_len,
                          int32_t max_doc, const uint8_t* norms /*max_doc bytes or NULL*/,
                          const uint64_t* live_docs /*bitset words or NULL=all live*/,
                          const orc_term_state* terms, uint32_t n_terms,
                          int64_t field_doc_count, int64_t sum_total_term_freq,
                          int64_t sum_doc_freq);
/* IndexSearcher::search for a batch, one query per thread (n_threads>=1).
 * out_hits: n_queries*k, out_counts: hits returned per query, out_total: total_hits.
 * parallel_mode=0: searcher.rs:487-525 (sequential leaves, one collector);
 * parallel_mode=1: search_parallel semantics with leaves merged in leaf order
 *                  (searcher.rs:527-630, top_docs.rs:157-172). */
int orc_search_batch(orc_index*, const orc_query* queries, uint32_t n_queries,
                     const orc_clause* clauses, uint32_t k, int parallel_mode,
                     int n_threads, orc_hit* out_hits, uint32_t* out_counts,
                     uint64_t* out_total);
/* One segment's TopDocsLeafCollector result per query in the engine's leaf-record layout
 * {u32 n; u32 pad; u64 total_hits; orc_hit heap[k]} (heap-array order, top_docs.rs:203-213). */
int orc_search_leaf_records(orc_index*, uint32_t seg, const orc_query* queries, uint32_t n_queries,
                            const orc_clause* clauses, uint32_t k, int n_threads, uint8_t* out_records);
/* BM25 weight pieces, for host-side cross checks. */
int orc_term_weight(orc_index*, uint32_t term_id, float boost, float* out_weight,
                    float* out_idf, float* out_avgdl, float out_cache[256]);
/* Iterate a term's postings on segment `seg` with BlockDocIterator::next();
 * returns number written (<= cap). */
int64_t orc_postings(orc_index*, uint32_t seg, uint32_t term_id, int32_t* docs,
                     int32_t* freqs, int64_t cap);
/* Drive BlockDocIterator::advance(target) for each target in order; out[i]=doc, freq. */
/* one DocIterator over the term: advance(targets[i]) per entry, or next() where targets[i] == -1 */
int orc_advance_seq(orc_index*, uint32_t seg, uint32_t term_id, const int32_t* targets,
                    uint32_t n, int32_t* out_docs, int32_t* out_freqs);

/* ---- block codec ---- */
/* ForUtil::read_block on a raw block stream.  forutil_table: the 32 header
 * vints' values ((format_id<<5)|(bpv-1)); doc_version selects the SIMD path
 * (version>0) exactly like posting_reader.rs:103-107.  Decodes n_blocks
 * consecutive blocks starting at byte offsets[i] into out[i*128..]. */
int orc_forutil_decode(const uint8_t* stream, size_t len, const uint64_t* offsets,
                       uint32_t n_blocks, int doc_version, const int32_t forutil_table[32],
                       int32_t* out, int n_threads);
void orc_simd_pack(const uint32_t* data128, uint8_t* encoded, int bits);
void orc_simd_unpack(const uint8_t* encoded, uint32_t* data128, int bits);
void orc_simd_delta_pack(const uint32_t* data128, uint8_t* encoded, uint32_t base, int bits);
void orc_simd_delta_unpack(const uint8_t* encoded, uint32_t* data128, uint32_t base, int bits);
int orc_simd_max_bits(const uint32_t* data128);
/* format_id 0 = Packed, 1 = PackedSingleBlock; decodes `iterations` rounds
 * (BulkOperation::decode_byte_to_int). Returns number of values written. */
int orc_packed_decode(int format_id, int bpv, const uint8_t* blocks, size_t n_bytes,
                      int32_t* values, int iterations);
int orc_packed_encode(int format_id, int bpv, const int32_t* values, uint8_t* blocks,
                      int iterations);
int orc_packed_iterations(int format_id, int bpv);   /* compute_iterations */
int orc_packed_encoded_size(int format_id, int bpv); /* Format::byte_count(128,bpv) */
int orc_max_data_size(void);                         /* for_util.rs:64-97 */
int orc_fastest_format(int bpv, float overhead, int* out_bpv); /* FormatAndBits::fastest(128,..) */
/* first index i with data[i] >= target in a sorted 128 block (simd_block_decoder.rs:100-128) */
int orc_block_advance(const int32_t* sorted128, int32_t target);

/* ---- scorers / collector on mock scorers (reference search/mod.rs:209-367) ---- */
/* Conjunction of mock iterators (score=docid as f32). lists: concatenated docids,
 * lens[i] = length of list i.  Emits (doc,score) for next() until exhaustion. */
int orc_mock_conjunction(const int32_t* lists, const uint32_t* lens, uint32_t n_lists,
                         int32_t* out_docs, float* out_scores, uint32_t cap);
int orc_mock_disjunction(const int32_t* lists, const uint32_t* lens, uint32_t n_lists,
                         int32_t min_should_match, int32_t* out_docs, float* out_scores,
                         uint32_t cap);
/* ReqOpt(req=list0, opt=list1) / ReqNot(req=list0, not=list1) on mock scorers. */
int orc_mock_req_opt(const int32_t* req, uint32_t n_req, const int32_t* opt, uint32_t n_opt,
                     int32_t* out_docs, float* out_scores, uint32_t cap);
int orc_mock_req_not(const int32_t* req, uint32_t n_req, const int32_t* nots, uint32_t n_not,
                     int32_t* out_docs, float* out_scores, uint32_t cap);
/* Generic mock scorer tree + op script (see oracle.cpp: parse_mock / orc_mock_run). */
int orc_mock_run(const int32_t* spec, uint32_t n_spec, const int32_t* ops, uint32_t n_ops,
                 int32_t* out_docs, float* out_scores);
/* TopDocsCollector over an explicit (doc,score) stream: returns hits (desc) and the raw
 * heap-array order (what BinaryHeap::into_vec yields). */
int orc_topk_stream(const int32_t* docs, const float* scores, uint64_t n, uint32_t k,
                    orc_hit* out_sorted, orc_hit* out_heap_order, uint32_t* out_count);
/* finish_parallel: replay add_doc over per-leaf heap arrays in leaf order. */
int orc_topk_merge(const orc_hit* leaf_hits, const uint32_t* leaf_counts, uint32_t n_leaves,
                   uint32_t k, orc_hit* out_sorted, uint32_t* out_count);

/* ---- BM25 / SmallFloat ---- */
uint8_t orc_float_to_byte315(float f);
float orc_byte315_to_float(uint8_t b);
float orc_norm_table(int i);
float orc_bm25_idf(int64_t doc_freq, int64_t doc_count);
float orc_bm25_avgdl(int64_t sum_total_term_freq, int64_t doc_count, int64_t max_doc);
float orc_bm25_score(float weight, float k1, float freq, float norm_cache_value);
void orc_bm25_cache(float k1, float b, float avgdl, float out_cache[256]);
uint8_t orc_encode_norm(float boost, int32_t field_length);

int orc_hardware_threads(void);

#ifdef __cplusplus
}
#endif
#endif
