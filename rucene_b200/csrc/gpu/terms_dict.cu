// terms_dict.cu — exact term lookups on the device for a whole batch.
//
// In the reference every TermQuery pays one SegmentTermIterator::seek_exact per segment
// (codec/postings/blocktree/blocktree_reader.rs:1364: FST walk over `.tip`, block scan over `.tim`,
// term_iter_frame.rs:436) followed by Lucene50PostingsReader::decode_term (posting_reader.rs:264-306) to obtain
// the BlockTermState.  Here the dictionary of a segment is uploaded once as its terms in dictionary order (the
// BlockTree iterates them sorted, unsigned bytewise) with the engine-wide term id of each; a batch of query terms
// is resolved by one kernel — a binary search per (query term, segment) over the sorted byte strings — and the
// BlockTermState is the row rg_segment_upload already holds for that id.  (Parsing `.tim`/`.tip` themselves on the
// device — FST arcs, floor blocks, suffix/stats/metadata sections — is not done: see DESIGN.md.)
#include <algorithm>
#include <cstring>

#include "engine.hpp"

namespace rg {

struct DictDev {
    const uint8_t* bytes;
    const uint64_t* off;   // n + 1
    const uint32_t* ids;   // engine-wide term id of dictionary entry i
    uint32_t n;
    uint32_t pad;
};

// unsigned bytewise comparison, shorter string first on a common prefix (BytesRef::cmp)
__device__ __forceinline__ int cmp_bytes(const uint8_t* a, uint32_t la, const uint8_t* b, uint32_t lb) {
    const uint32_t n = min(la, lb);
    for (uint32_t i = 0; i < n; i++) {
        const int d = (int)a[i] - (int)b[i];
        if (d) return d;
    }
    return (int)la - (int)lb;
}

__global__ void __launch_bounds__(128)
k_terms_lookup(const DictDev* __restrict__ dicts, uint32_t n_dicts, const uint8_t* __restrict__ qbytes,
               const uint64_t* __restrict__ qoff, uint32_t n_terms, uint32_t* __restrict__ out /* [n_dicts][n_terms] */) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_terms * n_dicts) return;
    const uint32_t d = i / n_terms, t = i - d * n_terms;
    const DictDev dict = dicts[d];
    const uint8_t* key = qbytes + qoff[t];
    const uint32_t klen = (uint32_t)(qoff[t + 1] - qoff[t]);
    uint32_t lo = 0, hi = dict.n;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        const uint64_t o = dict.off[mid];
        if (cmp_bytes(dict.bytes + o, (uint32_t)(dict.off[mid + 1] - o), key, klen) < 0) lo = mid + 1;
        else hi = mid;
    }
    uint32_t id = 0xffffffffu;
    if (lo < dict.n) {
        const uint64_t o = dict.off[lo];
        if (cmp_bytes(dict.bytes + o, (uint32_t)(dict.off[lo + 1] - o), key, klen) == 0) id = dict.ids[lo];
    }
    out[i] = id;
}

}  // namespace rg

using namespace rg;

#define RG_TRY try {
#define RG_CATCH \
    }            \
    catch (...) { return translate_exception(); }

extern "C" {

int rg_terms_upload(rg_engine* e, uint32_t seg_ord, const uint8_t* bytes, const uint64_t* offsets,
                    const uint32_t* term_ids, uint32_t n_terms) {
    RG_TRY
    if (!e || !offsets || (n_terms && !bytes && offsets[n_terms] != 0)) throw ArgError("null argument");
    if (seg_ord >= e->segs.size()) throw ArgError("no such segment");
    Segment& seg = e->segs[seg_ord];
    if (offsets[0] != 0) throw ArgError("offsets[0] must be 0");
    for (uint32_t i = 0; i < n_terms; i++) {
        if (offsets[i + 1] < offsets[i]) throw ArgError("term offsets must not decrease");
        if (term_ids && term_ids[i] >= seg.host_terms.size()) throw ArgError("dictionary refers to a term id outside the segment's term table");
        if (i) {  // dictionary order: strictly increasing, unsigned bytewise
            const uint64_t la = offsets[i] - offsets[i - 1], lb = offsets[i + 1] - offsets[i];
            const int c = memcmp(bytes + offsets[i - 1], bytes + offsets[i], (size_t)std::min(la, lb));
            if (c > 0 || (c == 0 && la >= lb)) throw ArgError("dictionary terms must be sorted (unsigned bytewise) and unique");
        }
    }
    if (!term_ids && n_terms > seg.host_terms.size()) throw ArgError("dictionary is larger than the segment's term table");
    RG_CUDA_CHECK(cudaSetDevice(e->device));
    cudaStream_t st = e->stream;
    const uint64_t total = offsets[n_terms];
    seg.dict_bytes.alloc(std::max<uint64_t>(1, total));
    seg.dict_off.alloc((size_t)n_terms + 1);
    seg.dict_ids.alloc(std::max<uint32_t>(1, n_terms));
    if (total) RG_CUDA_CHECK(cudaMemcpyAsync(seg.dict_bytes.p, bytes, total, cudaMemcpyHostToDevice, st));
    RG_CUDA_CHECK(cudaMemcpyAsync(seg.dict_off.p, offsets, ((size_t)n_terms + 1) * 8, cudaMemcpyHostToDevice, st));
    std::vector<uint32_t> ids(n_terms);
    for (uint32_t i = 0; i < n_terms; i++) ids[i] = term_ids ? term_ids[i] : i;
    if (n_terms) RG_CUDA_CHECK(cudaMemcpyAsync(seg.dict_ids.p, ids.data(), (size_t)n_terms * 4, cudaMemcpyHostToDevice, st));
    RG_CUDA_CHECK(cudaStreamSynchronize(st));
    seg.dict_n = n_terms;
    seg.has_dict = true;
    return RG_OK;
    RG_CATCH
}

int rg_terms_lookup(rg_engine* e, const uint8_t* bytes, const uint64_t* offsets, uint32_t n, uint32_t* out_term_ids,
                    int32_t* out_doc_freq) {
    RG_TRY
    if (!e || !offsets || !out_term_ids || (n && offsets[n] && !bytes)) throw ArgError("null argument");
    const uint32_t n_segs = (uint32_t)e->segs.size();
    for (const Segment& s : e->segs)
        if (!s.has_dict) throw ArgError("rg_terms_lookup needs rg_terms_upload for every segment");
    for (uint32_t i = 0; i < n; i++)
        if (offsets[i + 1] < offsets[i]) throw ArgError("term offsets must not decrease");
    for (uint32_t i = 0; i < n; i++) out_term_ids[i] = 0xffffffffu;
    if (out_doc_freq) std::fill(out_doc_freq, out_doc_freq + (size_t)n_segs * n, 0);
    if (!n || !n_segs) return RG_OK;
    RG_CUDA_CHECK(cudaSetDevice(e->device));
    cudaStream_t st = e->stream;
    std::vector<DictDev> h(n_segs);
    for (uint32_t s = 0; s < n_segs; s++)
        h[s] = DictDev{e->segs[s].dict_bytes.p, e->segs[s].dict_off.p, e->segs[s].dict_ids.p, e->segs[s].dict_n, 0u};
    DevBuf<DictDev> d_dicts;
    DevBuf<uint8_t> d_bytes;
    DevBuf<uint64_t> d_off;
    DevBuf<uint32_t> d_out;
    d_dicts.alloc(n_segs);
    d_bytes.alloc(std::max<uint64_t>(1, offsets[n]));
    d_off.alloc((size_t)n + 1);
    d_out.alloc((size_t)n * n_segs);
    RG_CUDA_CHECK(cudaMemcpyAsync(d_dicts.p, h.data(), n_segs * sizeof(DictDev), cudaMemcpyHostToDevice, st));
    if (offsets[n]) RG_CUDA_CHECK(cudaMemcpyAsync(d_bytes.p, bytes, offsets[n], cudaMemcpyHostToDevice, st));
    RG_CUDA_CHECK(cudaMemcpyAsync(d_off.p, offsets, ((size_t)n + 1) * 8, cudaMemcpyHostToDevice, st));
    const uint32_t threads = n * n_segs;
    k_terms_lookup<<<(threads + 127) / 128, 128, 0, st>>>(d_dicts.p, n_segs, d_bytes.p, d_off.p, n, d_out.p);
    RG_CUDA_CHECK(cudaGetLastError());
    e->launches++;
    std::vector<uint32_t> found((size_t)n * n_segs);
    RG_CUDA_CHECK(cudaMemcpyAsync(found.data(), d_out.p, found.size() * 4, cudaMemcpyDeviceToHost, st));
    RG_CUDA_CHECK(cudaStreamSynchronize(st));
    for (uint32_t s = 0; s < n_segs; s++) {
        for (uint32_t i = 0; i < n; i++) {
            const uint32_t id = found[(size_t)s * n + i];
            if (id == 0xffffffffu) continue;
            if (out_term_ids[i] != 0xffffffffu && out_term_ids[i] != id)
                throw ArgError("dictionaries disagree on the engine-wide id of a term");
            out_term_ids[i] = id;
            if (out_doc_freq) out_doc_freq[(size_t)s * n + i] = e->segs[s].host_terms[id].doc_freq;
        }
    }
    return RG_OK;
    RG_CATCH
}

}  // extern "C"
