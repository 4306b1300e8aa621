// EXPERIMENT (not compiled into librucene_gpu.so): k_eval_or variant "presence bitmaps +
// MaxScore essential-clause scoring".  Functionally correct — it passed tests/test_gpu_search.py and
// tests/test_golden.py on B200 (round 1) — but measured 2x SLOWER than the shipped cached-block-
// stream kernel on the scaled C4 workload (104 ms vs 50 ms for 1024 queries on 10M docs):
//   * only ~2/3 of the postings are pruned (theta rises slowly with k=100), and
//   * the per-(doc, clause) random-access scoring runs at ~25% lane utilisation (ncu: 8.5 active
//     lanes in the clause bodies), 68+85 instructions per scored posting;
//   * smem ATOMS bit setting + a 131-instruction docid-only refill keep phase 1 at 1.7 instr/posting;
//   * the x9-unrolled clause loops overflow the instruction cache (no_instruction stall 2.07).
// Next-round leads: compact (doc, clause) pairs before scoring, set bits from registers at
// refill, keep the clause loops rolled, per-block score upper bounds for tighter pruning.
// To try it: replace the k_eval_or section of query_kernels.cu with this text.
#if 0
// ------------------------------------------------------------------------------------------
// k_eval_or  — one WARP per work item; presence bitmaps + essential-clause scoring.
// ------------------------------------------------------------------------------------------
// A work item is a (query, segment, docid range).  The warp walks it in windows of kWw docids that
// always start at a real posting.  Per window:
//   1. bitmaps: every clause's doc-delta blocks are unpacked and prefix-summed ONCE into a
//      128-docid stream cache (freq blocks are not touched); postings below the window end set a
//      bit in the clause's 1024-bit window bitmap.  Lane w owns word w (docs 32w..32w+31).
//   2. counts : U = OR of the bitmaps (& live docs) -> popcount = matches of the window, exactly
//      what BulkScorer would have collected (total_hits).
//   3. scoring: only docs that can still enter the top-k heap are scored.  With theta = proven
//      lower bound of the heap root, clauses are split MaxScore-style: the cheapest-weight
//      clauses whose clause-order f32 sum of w*(k1+1) stays <= theta are "non-essential" — a doc
//      matching only those scores <= theta and can never be collected into the heap (BM25's
//      tf-norm factor is <= 1 and f32 addition is monotone).  Every doc matching an essential
//      clause is scored exactly: for each clause present (clause order, from 0.0f — the order of
//      DisjunctionSumScorer::score_sum) the posting's ordinal = rank of the doc in the clause's
//      bitmap, its freq is extracted by random access into the freq block (unpack.cuh extract1),
//      then BM25.  theta == -inf (cold start) makes every clause essential.
//   4. emit   : scored docs with score > theta are appended in docid order to the candidate run.
constexpr int kOrWarps = 4;
constexpr int kOrThreads = kOrWarps * 32;
constexpr int kWw = 1024;            // docids per window
constexpr int kNewcW = 64;

struct WTerm {
    const int32_t* blk_last;
    const BlockDesc* blk_desc;
    const float* cache;
    uint32_t nb;        // full blocks
    uint32_t cblk;      // block held by the stream cache (nb = vint tail); nb+1 = exhausted
    uint32_t n;         // valid entries in the stream cache (docs < hi)
    uint32_t pos;       // next unconsumed entry
    uint32_t term_id;
    uint32_t last;      // no further block after the cached one
    uint32_t ord0;      // ordinal (block*128+index) of the first posting of the current window
    float w1;           // weight * (k1 + 1)
};

struct alignas(16) WarpShared {  // followed by topk[kcap] floats, then cdocs[T][128]
    uint32_t bits[kMaxTerms][32];
    float acc[kWw];
    WTerm term[kMaxTerms];
    float newc[kNewcW];
};

struct WEmit {
    float* topk;       // shared memory, kcap floats
    uint32_t topk_n;
    float theta_local;
    uint32_t theta_in;
    uint32_t run_slot, run_cap, run_cnt;
    uint32_t matches;
    bool overflow;
};

__device__ __forceinline__ void wtheta_recompute(const WEmit& em, uint32_t k, int lane, float& theta, int& argmin) {
    float m = INFINITY;
    int mi = 0;
    for (uint32_t j = lane; j < k; j += 32) {
        const float v = em.topk[j];
        if (v < m) {
            m = v;
            mi = (int)j;
        }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        const float om = __shfl_xor_sync(0xffffffffu, m, o);
        const int oi = __shfl_xor_sync(0xffffffffu, mi, o);
        if (om < m || (om == m && oi < mi)) {
            m = om;
            mi = oi;
        }
    }
    theta = m;
    argmin = mi;
}

__device__ __forceinline__ void wtheta_update(WEmit& em, uint32_t k, uint32_t kcap, int lane,
                                              const float* newc, uint32_t newc_n, uint32_t* theta_out) {
    const uint32_t n_new = min(newc_n, (uint32_t)kNewcW);
    if (n_new == 0 || k > kcap) return;
    __syncwarp();
    float theta = em.theta_local;
    int argmin = 0;
    uint32_t n = em.topk_n;
    if (n == k) wtheta_recompute(em, k, lane, theta, argmin);
    for (uint32_t i = 0; i < n_new; i++) {
        const float x = newc[i];
        if (n < k) {
            if (lane == 0) em.topk[n] = x;
            n++;
            __syncwarp();
            if (n == k) wtheta_recompute(em, k, lane, theta, argmin);
        } else if (x > theta) {
            if (lane == 0) em.topk[argmin] = x;
            __syncwarp();
            wtheta_recompute(em, k, lane, theta, argmin);
        }
    }
    em.topk_n = n;
    em.theta_local = n == k ? theta : -INFINITY;
    if (lane == 0) {
        uint32_t ord = em.theta_in;
        if (em.theta_local != -INFINITY) ord = max(ord, float_to_ordered(em.theta_local));
        if (ord > kOrderedNegInf) atomicMax(theta_out, ord);
    }
}

// freq of tail entry j (posting_reader.rs:308-333): sequential vint walk; singleton: total_term_freq
__device__ int tail_freq_at(const SegDev& seg, const TermDev& td, uint32_t j) {
    if (td.doc_freq == 1) return td.singleton_freq;
    const uint8_t* p = seg.tails + td.tail_off;
    uint32_t pos = 0;
    int f = 1;
    for (uint32_t i = 0; i <= j; i++) {
        const uint32_t code = (uint32_t)read_vint(p, pos);
        f = (code & 1u) ? 1 : read_vint(p, pos);
    }
    return f;
}

// Fill clause t's stream cache with the docids of block `b` (b == nb: vint tail / singleton).
// Doc deltas only: unpack + warp scan.  Entries >= hi are cut off; pos skips docs < lo.
__device__ __forceinline__ void stream_fill(const SegDev& seg, WTerm& tc, int32_t* cd, uint32_t b, int lo, int hi,
                                            int lane) {
    int4 docs;
    uint32_t n_in = kBlock;
    if (b < tc.nb) {
        const BlockDesc bd = tc.blk_desc[b];
        const int base = b == 0 ? 0 : __ldg(tc.blk_last + b - 1);
        const int4 dl = unpack4(seg.arena + bd.off16, (int)(bd.bits & 0xff), lane, seg.version, seg.sb_mask);
        docs = deltas_to_docs(dl, base);
        reinterpret_cast<int4*>(cd)[lane] = docs;
    } else {
        const TermDev td = seg.terms[tc.term_id];
        n_in = td.tail_n;
        if (lane == 0) {
            if (td.doc_freq == 1) {
                cd[0] = td.singleton_doc;
            } else {
                const uint8_t* p = seg.tails + td.tail_off;
                uint32_t pos = 0;
                int32_t acc = td.tail_base;
                for (uint32_t i = 0; i < td.tail_n; i++) {
                    const uint32_t code = (uint32_t)read_vint(p, pos);
                    acc += (int32_t)(code >> 1);
                    cd[i] = acc;
                    if (!(code & 1u)) read_vint(p, pos);
                }
            }
        }
        __syncwarp();
        const int i0 = 4 * lane;
        docs = make_int4(i0 < (int)n_in ? cd[i0] : kNoMoreDocs, i0 + 1 < (int)n_in ? cd[i0 + 1] : kNoMoreDocs,
                         i0 + 2 < (int)n_in ? cd[i0 + 2] : kNoMoreDocs, i0 + 3 < (int)n_in ? cd[i0 + 3] : kNoMoreDocs);
    }
    const uint32_t below = __reduce_add_sync(0xffffffffu, (uint32_t)((docs.x < lo) + (docs.y < lo) + (docs.z < lo) + (docs.w < lo)));
    const uint32_t under = __reduce_add_sync(0xffffffffu, (uint32_t)((docs.x < hi) + (docs.y < hi) + (docs.z < hi) + (docs.w < hi)));
    if (lane == 0) {
        tc.cblk = b;
        tc.pos = below;
        tc.n = under;
        tc.last = (under < n_in || b >= tc.nb || (b + 1 == tc.nb && seg.terms[tc.term_id].tail_n == 0)) ? 1u : 0u;
    }
    __syncwarp();
}

__global__ void __launch_bounds__(kOrThreads, 4)
k_eval_or(EvalParams p, const uint32_t* __restrict__ item_ids, uint32_t n_ids, uint32_t warp_bytes,
          uint32_t kcap) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = lane_id(), warp = threadIdx.x >> 5;
    const uint32_t wid = blockIdx.x * kOrWarps + warp;
    if (wid >= n_ids) return;
    unsigned char* base = smem_raw + (size_t)warp * warp_bytes;
    WarpShared& sh = *reinterpret_cast<WarpShared*>(base);
    float* topk = reinterpret_cast<float*>(base + sizeof(WarpShared));
    int32_t* cdocs = reinterpret_cast<int32_t*>(topk + kcap);
    const uint32_t item_idx = item_ids[wid];
    const WorkItem it = p.items[item_idx];
    const SegDev seg = p.segs[it.seg];
    const int T = it.n_terms;
    const int lo = it.lo, hi = it.hi;

#pragma unroll
    for (int t = 0; t < kMaxTerms; t++) sh.bits[t][lane] = 0;
    bool positive = true;  // MaxScore pruning needs non-negative clause weights
    if (lane < T) {
        const ItemClause c = p.clauses[it.clause_begin + lane];
        const TermDev td = seg.terms[c.term_id];
        WTerm& tc = sh.term[lane];
        tc.blk_last = seg.blk_last + td.blk_begin;
        tc.blk_desc = seg.blk_desc + td.blk_begin;
        tc.cache = p.caches + (size_t)c.cache_id * 256;
        tc.nb = td.n_blocks;
        tc.cblk = lower_bound_i32(tc.blk_last, 0, td.n_blocks, lo);
        tc.n = 0;
        tc.pos = 0;
        tc.term_id = c.term_id;
        tc.last = 0;
        tc.ord0 = 0;
        tc.w1 = __fmul_rn(c.weight, __fadd_rn(p.k1, 1.0f));
        positive = tc.w1 >= 0.0f;
    }
    positive = __all_sync(0xffffffffu, positive);
    __syncwarp();
    // prime the streams
    long long w0 = kNoMoreDocs;
    for (int t = 0; t < T; t++) {
        WTerm& tc = sh.term[t];
        int32_t* cd = cdocs + t * kBlock;
        uint32_t b = tc.cblk;
        for (;;) {
            const bool has_tail = seg.terms[tc.term_id].tail_n > 0;
            if (b > tc.nb || (b == tc.nb && !has_tail)) {
                if (lane == 0) {
                    tc.cblk = tc.nb + 1;
                    tc.n = tc.pos = 0;
                    tc.last = 1;
                }
                __syncwarp();
                break;
            }
            stream_fill(seg, tc, cd, b, lo, hi, lane);
            if (tc.pos < tc.n || tc.last) break;
            b++;
        }
        if (tc.pos < tc.n) w0 = min(w0, (long long)cd[tc.pos]);
    }

    WEmit em;
    em.topk = topk;
    em.topk_n = 0;
    em.theta_local = -INFINITY;
    em.theta_in = 0;
    em.run_slot = kNone;
    em.run_cap = 0;
    em.run_cnt = 0;
    em.matches = 0;
    em.overflow = false;
    const bool lb_ok = (uint32_t)lane < it.chain_pos;
    const uint32_t* theta_lb = p.item_theta + item_idx - 1 - (lb_ok ? lane : 0);
    uint32_t win_no = 0;
    uint32_t ess = (1u << T) - 1u;
    float ess_te = -INFINITY;

    while (w0 < hi) {
        const int win0 = (int)w0;
        const int win1 = (int)min((long long)hi, w0 + kWw);
        if ((win_no++ & 7u) == 0 && it.chain_pos) {
            uint32_t inh = lb_ok ? ld_volatile_u32(theta_lb) : 0u;
            inh = __reduce_max_sync(0xffffffffu, inh);
            em.theta_in = max(em.theta_in, inh);
        }
        int next_doc = kNoMoreDocs;
        // ---- 1. bitmaps: drain every clause's docid stream up to the window end
        for (int t = 0; t < T; t++) {
            WTerm& tc = sh.term[t];
            int32_t* cd = cdocs + t * kBlock;
            uint32_t pos = tc.pos, n = tc.n;
            if (lane == 0) tc.ord0 = tc.cblk * kBlock + pos;
            for (;;) {
                if (pos >= n) {
                    if (tc.last) break;
                    stream_fill(seg, tc, cd, tc.cblk + 1, lo, hi, lane);
                    pos = tc.pos;
                    n = tc.n;
                    if (pos >= n) continue;  // (empty after trimming) -> last is set
                }
                const uint32_t i = pos + lane;
                const int d = i < n ? cd[i] : kNoMoreDocs;
                const bool in_win = d < win1;
                const uint32_t c = __popc(__ballot_sync(0xffffffffu, in_win));  // sorted: a prefix
                if (in_win) atomicOr(&sh.bits[t][(d - win0) >> 5], 1u << ((d - win0) & 31));
                pos += c;
                if (c < 32 && pos < n) break;  // next cached doc is beyond this window
            }
            if (lane == 0) tc.pos = pos;
            if (pos < n) next_doc = min(next_doc, cd[pos]);
            __syncwarp();
        }
        // ---- 2. counts + essential split
        float te = em.theta_local;
        if (em.theta_in > kOrderedNegInf) te = fmaxf(te, ordered_to_float(em.theta_in));
        const bool open = te == -INFINITY;
        if (te != ess_te) {  // theta moved: redo the split (O(T^2), T <= 9)
        ess_te = te;
        ess = (1u << T) - 1u;  // bit t set: clause t is essential
        if (!open && positive) {
            // move clauses to the non-essential set in ascending w1 while the clause-order f32 sum of
            // their w1 stays <= theta
            uint32_t non = 0;
            for (int round = 0; round < T; round++) {
                int best = -1;
                float bw = INFINITY;
                for (int t = 0; t < T; t++)
                    if (!((non >> t) & 1u) && sh.term[t].w1 < bw) {
                        bw = sh.term[t].w1;
                        best = t;
                    }
                const uint32_t trial = non | (1u << best);
                float ub = 0.0f;
                for (int t = 0; t < T; t++)
                    if ((trial >> t) & 1u) ub = __fadd_rn(ub, sh.term[t].w1);
                if (ub <= te) non = trial;
                else break;
            }
            ess &= ~non;
        }
        }
        uint32_t bw_[kMaxTerms];
        uint32_t U = 0, E = 0;
#pragma unroll
        for (int t = 0; t < kMaxTerms; t++) {
            bw_[t] = 0;
            if (t < T) {
                bw_[t] = sh.bits[t][lane];
                sh.bits[t][lane] = 0;
                U |= bw_[t];
                if ((ess >> t) & 1u) E |= bw_[t];
            }
        }
        if (seg.live) {
            uint32_t lv = 0;
            const int d0 = win0 + 32 * lane;
            for (int j = 0; j < 32; j++) {
                const int d = d0 + j;
                if (d < win1 && ((seg.live[d >> 6] >> (d & 63)) & 1ull)) lv |= 1u << j;
            }
            U &= lv;
            E &= lv;
        }
        em.matches += __reduce_add_sync(0xffffffffu, (uint32_t)__popc(U));
        // ---- 3. exact scores of the docs that match an essential clause
        uint32_t C = 0;  // candidates of this lane's word
        if (__any_sync(0xffffffffu, E != 0)) {
            uint32_t pre[kMaxTerms];
#pragma unroll
            for (int t = 0; t < kMaxTerms; t++) {
                pre[t] = 0;
                if (t < T) {
                    const uint32_t cnt = __popc(bw_[t]);
                    uint32_t incl = cnt;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
                        if (lane >= o) incl += v;
                    }
                    pre[t] = incl - cnt + sh.term[t].ord0;  // ordinal of this word's first posting
                }
            }
            uint32_t todo = E;
            while (todo) {
                const int j = __ffs(todo) - 1;
                todo &= todo - 1;
                const int d = win0 + 32 * lane + j;
                const uint32_t below = (1u << j) - 1u;
                float score = 0.0f;
                const float nrm_byte = 0.f;
                (void)nrm_byte;
                const uint32_t nb8 = seg.norms ? (uint32_t)__ldg(seg.norms + d) : 0u;
#pragma unroll
                for (int t = 0; t < kMaxTerms; t++) {
                    if (t < T && ((bw_[t] >> j) & 1u)) {
                        const WTerm& tc = sh.term[t];
                        const uint32_t ord = pre[t] + __popc(bw_[t] & below);
                        const uint32_t b = ord >> 7, idx = ord & 127u;
                        int f;
                        if (b < tc.nb) {
                            const BlockDesc bd = tc.blk_desc[b];
                            f = extract1(seg.arena + bd.off16 + ((bd.bits >> 16) & 0xff), (int)((bd.bits >> 8) & 0xff),
                                         (int)idx, seg.version, seg.sb_mask);
                        } else {
                            f = tail_freq_at(seg, seg.terms[tc.term_id], idx);
                        }
                        const float nrm = seg.norms ? __ldg(tc.cache + nb8) : p.k1;
                        score = __fadd_rn(score, bm25_score(tc.w1, (float)f, nrm));
                    }
                }
                sh.acc[32 * lane + j] = score;
                if (open || score > te) C |= 1u << j;
            }
        }
        // ---- 4. emit candidates in docid order (lane = word order, bit order inside)
        const uint32_t cc = __popc(C);
        uint32_t incl = cc;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        const uint32_t total_c = __shfl_sync(0xffffffffu, incl, 31);
        if (total_c && !em.overflow) {
            CandRun* hdr = reinterpret_cast<CandRun*>(p.cand_arena);
            if (em.run_slot == kNone || em.run_cnt + total_c > em.run_cap) {
                uint32_t slot = 0;
                const uint32_t cap = max(total_c, em.run_slot == kNone ? kRunFirst : kRunMin);
                if (lane == 0) {
                    const unsigned long long s64 = atomicAdd(p.arena_next, (unsigned long long)cap + 1ull);
                    slot = (s64 + cap + 1ull > (unsigned long long)p.arena_slots) ? kNone : (uint32_t)s64;
                    if (slot == kNone) atomicOr(p.error_flag, 1u);
                    else if (em.run_slot == kNone) p.item_head[item_idx] = slot;
                    else hdr[em.run_slot] = CandRun{slot, em.run_cnt};
                }
                slot = __shfl_sync(0xffffffffu, slot, 0);
                if (slot == kNone) {
                    em.overflow = true;
                } else {
                    em.run_slot = slot;
                    em.run_cap = cap;
                    em.run_cnt = 0;
                }
            }
            if (!em.overflow) {
                uint32_t o = incl - cc;  // exclusive prefix
                uint32_t cbits = C;
                while (cbits) {
                    const int j = __ffs(cbits) - 1;
                    cbits &= cbits - 1;
                    const float sc = sh.acc[32 * lane + j];
                    p.cand_arena[em.run_slot + 1 + em.run_cnt + o] = rg_hit{win0 + 32 * lane + j + seg.doc_base, sc};
                    if (o < (uint32_t)kNewcW) sh.newc[o] = sc;
                    o++;
                }
                em.run_cnt += total_c;
                if (lane == 0) hdr[em.run_slot] = CandRun{kNone, em.run_cnt};
                wtheta_update(em, p.k, kcap, lane, sh.newc, total_c, p.item_theta + item_idx);
            }
        }
        __syncwarp();
        next_doc = __reduce_min_sync(0xffffffffu, next_doc);
        if (next_doc == kNoMoreDocs) break;
        w0 = next_doc;
    }
    if (lane == 0) p.item_matches[item_idx] = em.matches;
}

#endif
