// EXPERIMENT (not compiled into librucene_gpu.so): k_eval_or "presence bitmaps + MaxScore
// essential-clause scoring", second iteration (bits set from registers at refill, essential docs
// compacted over the lanes in docid order, no accumulator window).  Correct — passed
// tests/test_gpu_search.py + tests/test_golden.py on B200 — but 76.8 ms vs 43.3 ms for the shipped
// cached-block-stream kernel on the scaled C4 workload (1024 queries, 10M docs).  ncu (or_v8):
//   * pruning works: only ~14 % of the postings belong to docs that need a score (simulation of the
//     ideal: 9-20 %), but scoring them costs ~20 warp instructions per doc: 8.3M compacted chunks
//     average 16 of 32 lanes, and the per-clause bodies (70 instr) diverge across clauses;
//   * docid-only refill: 186 instr/block (1.14/posting); per-window fixed work (essential split,
//     per-clause ordinal prefixes, bitmap clears) ~450 instr x 1.8M windows = 12 %;
//   * total 7.15 warp instr/posting vs 4.67 for the shipped kernel.
// Leads: iterate each lane's own present clauses (bit mask) instead of all clauses, term-major
// compaction of (doc, clause) pairs, bigger windows to fill the chunks, per-block score bounds.
#if 0
// ------------------------------------------------------------------------------------------
// k_eval_or  — one WARP per work item; presence bitmaps + essential-clause scoring.
// ------------------------------------------------------------------------------------------
// A work item is a (query, segment, docid range).  The warp walks it in windows of kWw docids that
// always start at a real posting.  Per window:
//   1. bitmaps: each clause's doc-delta blocks are unpacked and prefix-summed ONCE (freq blocks
//      are not touched); postings below the window end set a bit in the clause's 1024-bit window
//      bitmap straight from registers, the rest of the block waits in a 128-docid stream cache.
//      Lane w owns word w (docs 32w..32w+31) of every bitmap.
//   2. counts : U = OR of the bitmaps (& live docs); popcount(U) = the docs BulkScorer would have
//      collected in this window (total_hits), exactly.
//   3. scoring: only docs that can still enter the top-k heap are scored.  theta is a proven lower
//      bound of the heap root; clauses are split MaxScore-style: the lowest-weight clauses whose
//      clause-order f32 sum of w*(k1+1) stays <= theta are non-essential — a doc matching only
//      those scores <= theta (BM25's tf-norm factor is <= 1, f32 addition is monotone) and could
//      never replace the heap root (top_docs.rs:72 needs root.score < score).  Docs matching an
//      essential clause (bitmap E) are compacted over the lanes in docid order and scored exactly:
//      for each clause present, in clause order from 0.0f (DisjunctionSumScorer::score_sum), the
//      posting's ordinal = clause cursor + rank of the doc inside the clause's bitmap, its freq is
//      read by random access into the freq block (extract1), then BM25 with IEEE ops.
//      theta == -inf (cold start) makes every clause essential, i.e. every doc is scored.
//   4. emit   : scored docs with score > theta go, in docid order, to the candidate run.
constexpr int kOrWarps = 4;
constexpr int kOrThreads = kOrWarps * 32;
constexpr int kWw = 1024;            // docids per window
constexpr int kNewcW = 64;

struct WTerm {
    const int32_t* blk_last;
    const BlockDesc* blk_desc;
    const float* cache;
    uint32_t nb;        // full blocks
    uint32_t cblk;      // block held by the stream cache (nb = vint tail)
    uint32_t n;         // valid entries in the stream cache (docs < hi)
    uint32_t pos;       // next unconsumed entry
    uint32_t term_id;
    uint32_t last;      // no further block after the cached one
    uint32_t ord0;      // ordinal (block*128+index) of the first posting of the current window
    float w1;           // weight * (k1 + 1)
};

struct alignas(16) WarpShared {  // followed by topk[kcap] floats, then cdocs[T][128]
    uint32_t bits[kMaxTerms][32];
    uint32_t pre[kMaxTerms][32];   // ordinal of the first posting of word w, per clause
    WTerm term[kMaxTerms];
    float newc[kNewcW];
};

struct WEmit {
    float* topk;       // shared memory, kcap floats
    uint32_t topk_n;
    float theta_local;
    uint32_t theta_in;
    uint32_t run_slot, run_cap, run_cnt;
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

__device__ __noinline__ void wtheta_update(WEmit& em, uint32_t k, uint32_t kcap, int lane, const float* newc,
                                           uint32_t newc_n, uint32_t* theta_out) {
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
__device__ __noinline__ int tail_freq_at(const SegDev& seg, uint32_t term_id, uint32_t j) {
    const TermDev td = seg.terms[term_id];
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

// Decode the docids of clause block `b` (b == nb: vint tail / singleton) into the stream cache and
// set the window bits of its postings below win1 straight from registers.  Doc deltas only.
// Entries >= hi are cut off; entries < lo (first block of the range only) are skipped.
__device__ __noinline__ void stream_fill(const SegDev& seg, WTerm& tc, int32_t* cd, uint32_t* bits, uint32_t b,
                                         int lo, int hi, int win0, int win1, int lane) {
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
    const int d[4] = {docs.x, docs.y, docs.z, docs.w};
    // window bits: consecutive postings of a lane usually share a word -> merge before the atomic
    uint32_t below = 0, under = 0, inwin = 0;
    int cw = -1;
    uint32_t cm = 0;
#pragma unroll
    for (int q = 0; q < 4; q++) {
        below += d[q] < lo;
        under += d[q] < hi;
        if (d[q] >= lo && d[q] < win1) {
            inwin++;
            const int r = d[q] - win0;
            const int w = r >> 5;
            if (w != cw) {
                if (cm) atomicOr(bits + cw, cm);
                cw = w;
                cm = 0;
            }
            cm |= 1u << (r & 31);
        }
    }
    if (cm) atomicOr(bits + cw, cm);
    below = __reduce_add_sync(0xffffffffu, below);
    under = __reduce_add_sync(0xffffffffu, under);
    inwin = __reduce_add_sync(0xffffffffu, inwin);
    if (lane == 0) {
        tc.cblk = b;
        tc.pos = below + inwin;
        tc.n = under;
        tc.last = (under < n_in || b >= tc.nb || (b + 1 == tc.nb && seg.terms[tc.term_id].tail_n == 0)) ? 1u : 0u;
    }
    __syncwarp();
}

__global__ void __launch_bounds__(kOrThreads, 6)
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

    for (int t = 0; t < T; t++) sh.bits[t][lane] = 0;
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
        tc.last = (tc.cblk > td.n_blocks || (tc.cblk == td.n_blocks && td.tail_n == 0)) ? 1u : 0u;
        tc.ord0 = 0;
        tc.w1 = __fmul_rn(c.weight, __fadd_rn(p.k1, 1.0f));
        positive = tc.w1 >= 0.0f;
    }
    positive = __all_sync(0xffffffffu, positive);
    __syncwarp();
    // prime the streams (no window yet: win1 = lo sets no bits)
    long long w0 = kNoMoreDocs;
    for (int t = 0; t < T; t++) {
        WTerm& tc = sh.term[t];
        int32_t* cd = cdocs + t * kBlock;
        uint32_t b = tc.cblk;
        while (!tc.last || b == tc.cblk) {
            if (tc.last && tc.n == 0 && (b > tc.nb || (b == tc.nb && seg.terms[tc.term_id].tail_n == 0))) break;
            stream_fill(seg, tc, cd, sh.bits[t], b, lo, hi, lo, lo, lane);
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
    em.overflow = false;
    const bool lb_ok = (uint32_t)lane < it.chain_pos;
    const uint32_t* theta_lb = p.item_theta + item_idx - 1 - (lb_ok ? lane : 0);
    uint32_t win_no = 0, my_matches = 0;
    uint32_t ess = (1u << T) - 1u;
    float ess_te = -INFINITY;
    const bool has_norms = seg.norms != nullptr;

    while (w0 < hi) {
        const int win0 = (int)w0;
        const int win1 = (int)min((long long)hi, w0 + kWw);
        if ((win_no++ & 7u) == 0 && it.chain_pos) {
            uint32_t inh = lb_ok ? ld_volatile_u32(theta_lb) : 0u;
            inh = __reduce_max_sync(0xffffffffu, inh);
            em.theta_in = max(em.theta_in, inh);
        }
        int next_doc = kNoMoreDocs;
        // ---- 1. bitmaps
        for (int t = 0; t < T; t++) {
            WTerm& tc = sh.term[t];
            int32_t* cd = cdocs + t * kBlock;
            uint32_t* bits = sh.bits[t];
            uint32_t pos = tc.pos, n = tc.n;
            if (lane == 0) tc.ord0 = tc.cblk * kBlock + pos;
            // leftovers of the cached block
            while (pos < n) {
                const uint32_t i = pos + lane;
                const int d = i < n ? cd[i] : kNoMoreDocs;
                const bool in_win = d < win1;
                const uint32_t c = __popc(__ballot_sync(0xffffffffu, in_win));  // sorted: a prefix
                if (in_win) atomicOr(bits + ((d - win0) >> 5), 1u << ((d - win0) & 31));
                pos += c;
                if (c < 32) break;
            }
            // further blocks that start inside the window
            if (pos >= n) {
                if (lane == 0) tc.pos = pos;
                __syncwarp();
                while (!tc.last) {
                    stream_fill(seg, tc, cd, bits, tc.cblk + 1, lo, hi, win0, win1, lane);
                    if (tc.pos < tc.n) break;
                }
                pos = tc.pos;
                n = tc.n;
            } else if (lane == 0) {
                tc.pos = pos;
            }
            if (pos < n) next_doc = min(next_doc, cd[pos]);
            __syncwarp();
        }
        // ---- 2. theta, essential split, counts
        float te = em.theta_local;
        if (em.theta_in > kOrderedNegInf) te = fmaxf(te, ordered_to_float(em.theta_in));
        const bool open = te == -INFINITY;
        if (te != ess_te) {  // theta moved: redo the split (O(T^2), T <= 9)
            ess_te = te;
            ess = (1u << T) - 1u;
            if (!open && positive) {
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
        uint32_t U = 0, E = 0;
        for (int t = 0; t < T; t++) {
            const uint32_t bw = sh.bits[t][lane];
            U |= bw;
            if ((ess >> t) & 1u) E |= bw;
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
        my_matches += __popc(U);
        // ---- 3+4. exact scores of the E docs, compacted over the lanes in docid order
        const uint32_t ecnt = __popc(E);
        uint32_t eincl = ecnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, eincl, o);
            if (lane >= o) eincl += v;
        }
        const uint32_t total_e = __shfl_sync(0xffffffffu, eincl, 31);
        uint32_t newc_n = 0;
        if (total_e) {
            // ordinal of the first posting of each word, per clause
            for (int t = 0; t < T; t++) {
                const uint32_t cnt = __popc(sh.bits[t][lane]);
                uint32_t incl = cnt;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += v;
                }
                sh.pre[t][lane] = incl - cnt + sh.term[t].ord0;
            }
            __syncwarp();
            for (uint32_t r0 = 0; r0 < total_e; r0 += 32) {
                const uint32_t r = r0 + lane;
                const bool act = r < total_e;
                // word holding the r-th E doc: smallest w with eincl[w] > r
                int w = 0;
#pragma unroll
                for (int step = 16; step; step >>= 1) {
                    const uint32_t v = __shfl_sync(0xffffffffu, eincl, w + step - 1);
                    if (v <= r) w += step;
                }
                w = min(w, 31);
                const uint32_t Ew = __shfl_sync(0xffffffffu, E, w);
                const uint32_t ebase = __shfl_sync(0xffffffffu, eincl - ecnt, w);
                float score = 0.0f;
                int doc = 0;
                if (act) {
                    // (r - ebase)-th set bit of Ew
                    uint32_t m = Ew;
                    for (uint32_t s = r - ebase; s; s--) m &= m - 1;
                    const int j = __ffs(m) - 1;
                    doc = win0 + 32 * w + j;
                    const uint32_t below = (1u << j) - 1u;
                    const uint32_t nb8 = has_norms ? (uint32_t)__ldg(seg.norms + doc) : 0u;
                    for (int t = 0; t < T; t++) {
                        const uint32_t bw = sh.bits[t][w];
                        if ((bw >> j) & 1u) {
                            const WTerm& tc = sh.term[t];
                            const uint32_t ord = sh.pre[t][w] + __popc(bw & below);
                            const uint32_t b = ord >> 7, idx = ord & 127u;
                            int f;
                            if (b < tc.nb) {
                                const BlockDesc bd = tc.blk_desc[b];
                                f = extract1(seg.arena + bd.off16 + ((bd.bits >> 16) & 0xff),
                                             (int)((bd.bits >> 8) & 0xff), (int)idx, seg.version, seg.sb_mask);
                            } else {
                                f = tail_freq_at(seg, tc.term_id, idx);
                            }
                            const float nrm = has_norms ? __ldg(tc.cache + nb8) : p.k1;
                            score = __fadd_rn(score, bm25_score(tc.w1, (float)f, nrm));
                        }
                    }
                }
                const bool cand = act && (open || score > te);
                const uint32_t cmask = __ballot_sync(0xffffffffu, cand);
                if (!cmask || em.overflow) continue;
                const uint32_t c = __popc(cmask);
                CandRun* hdr = reinterpret_cast<CandRun*>(p.cand_arena);
                if (em.run_slot == kNone || em.run_cnt + c > em.run_cap) {
                    uint32_t slot = 0;
                    const uint32_t cap = em.run_slot == kNone ? kRunFirst : kRunMin;
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
                        continue;
                    }
                    em.run_slot = slot;
                    em.run_cap = cap;
                    em.run_cnt = 0;
                }
                if (cand) {
                    const uint32_t rr = __popc(cmask & ((1u << lane) - 1u));
                    p.cand_arena[em.run_slot + 1 + em.run_cnt + rr] = rg_hit{doc + seg.doc_base, score};
                    if (newc_n + rr < (uint32_t)kNewcW) sh.newc[newc_n + rr] = score;
                }
                em.run_cnt += c;
                newc_n += c;
                if (lane == 0) hdr[em.run_slot] = CandRun{kNone, em.run_cnt};
            }
        }
        __syncwarp();
        for (int t = 0; t < T; t++) sh.bits[t][lane] = 0;
        if (newc_n) wtheta_update(em, p.k, kcap, lane, sh.newc, newc_n, p.item_theta + item_idx);
        __syncwarp();
        next_doc = __reduce_min_sync(0xffffffffu, next_doc);
        if (next_doc == kNoMoreDocs) break;
        w0 = next_doc;
    }
    my_matches = __reduce_add_sync(0xffffffffu, my_matches);
    if (lane == 0) p.item_matches[item_idx] = my_matches;
}

#endif
