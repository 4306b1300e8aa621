// The round-1 "cached block streams" k_eval_or (every block unpacked, prefix-summed and BM25-scored
// exactly once into a per-clause shared-memory stream cache; warp-private f32 accumulator window).
// Kept for reference/fallback: 79 G postings/s on the scaled C4 workload, 4.7 warp instructions per
// posting (profiles/r1_eval_or_hotspots.txt).  To use it: replace the k_eval_or section of
// query_kernels.cu with this text.
#if 0
// ------------------------------------------------------------------------------------------
// k_eval_or  — one WARP per work item, no block-level synchronisation at all.
// ------------------------------------------------------------------------------------------
// A work item is a (query, segment, docid range) of ~32K postings.  Every clause is a *cached
// block stream*: its current 128-posting block lives decoded AND scored in shared memory
// (docids + BM25 scores), so each block is unpacked, prefix-summed and scored exactly once.
// The warp walks the range in windows of kWw docids that always start at a real posting:
//   for clause t = 0..T-1 (clause order): consume the stream's postings < window end, 32 per
//       step, "acc[d] = acc[d] + s" in the warp-private accumulator window — pair order ==
//       clause order == DisjunctionSumScorer::score_sum's f32 order; refill the stream (decode the
//       next block / the vint tail) whenever it runs dry;
//   scan the touched 32-doc steps in docid order -> total_hits, theta filter, candidates;
//   next window start = min over clauses of their next cached docid (exact).
constexpr int kOrWarps = 4;
constexpr int kOrThreads = kOrWarps * 32;
constexpr int kWw = 1024;            // docids per window
constexpr int kNewcW = 64;

struct WTerm {
    const int32_t* blk_last;
    const BlockDesc* blk_desc;
    const float* cache;
    uint32_t nb;        // full blocks
    uint32_t cur;       // next block to decode (nb = vint tail, nb+1 = exhausted)
    uint32_t n;         // valid entries in the stream cache
    uint32_t pos;       // next unconsumed entry
    uint32_t term_id;
    float w1;           // weight * (k1 + 1)
};

struct WarpShared {            // followed by topk[kcap] floats, then cdocs[T][128], cscores[T][128]
    uint32_t acc[kWw];
    WTerm term[kMaxTerms];
    float newc[kNewcW];
};

// warp-level candidate emitter state (registers, uniform across lanes)
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

// Emit one 32-doc step (docid order).  `newc`/`newc_n`: this window's candidate scores for the
// theta tracker.
__device__ __forceinline__ void wemit_step(WEmit& em, const EvalParams& p, uint32_t item_idx, int lane,
                                           bool present, int gdoc, float score, float te, bool open,
                                           float* newc, uint32_t& newc_n) {
    const uint32_t pm = __ballot_sync(0xffffffffu, present);
    if (!pm) return;
    em.matches += __popc(pm);
    const uint32_t cm = __ballot_sync(0xffffffffu, present && (open || score > te));
    if (!cm || em.overflow) return;
    const uint32_t c = __popc(cm);
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
            return;
        }
        em.run_slot = slot;
        em.run_cap = cap;
        em.run_cnt = 0;
    }
    if ((cm >> lane) & 1u) {
        const uint32_t r = __popc(cm & ((1u << lane) - 1u));
        p.cand_arena[em.run_slot + 1 + em.run_cnt + r] = rg_hit{gdoc, score};
        if (newc_n + r < (uint32_t)kNewcW) newc[newc_n + r] = score;
    }
    em.run_cnt += c;
    newc_n += c;
    if (lane == 0) hdr[em.run_slot] = CandRun{kNone, em.run_cnt};
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

// Refill clause t's stream cache with its next block (or vint tail): unpack, docid scan, norm
// gather, BM25 — once per block.  Entries outside [lo, hi) are trimmed.  Returns false when the
// list is exhausted.  Warp-cooperative; all lanes must call it.
// When called while clause t is being drained into the window [win0, win1) the new block's
// postings below win1 are accumulated straight from registers (no round trip through the cache).
__device__ __forceinline__ bool stream_refill(const SegDev& seg, const EvalParams& p, WTerm& tc, int32_t* cd,
                                           float* cs, int lo, int hi, int lane, int win0, int win1,
                                           uint32_t* acc, uint32_t& touched, uint32_t& hot, uint32_t& my_matches,
                                           float te) {
    for (;;) {
        const uint32_t b = tc.cur;
        if (b > tc.nb) return false;
        int4 docs, freqs;
        uint32_t n_in = kBlock;
        if (b < tc.nb) {
            const BlockDesc bd = tc.blk_desc[b];
            const int base = b == 0 ? 0 : __ldg(tc.blk_last + b - 1);
            const uint4* part = seg.arena + bd.off16;
            const int4 dl = unpack4(part, (int)(bd.bits & 0xff), lane, seg.version, seg.sb_mask);
            freqs = unpack4(part + ((bd.bits >> 16) & 0xff), (int)((bd.bits >> 8) & 0xff), lane, seg.version,
                            seg.sb_mask);
            docs = deltas_to_docs(dl, base);
        } else {  // vint tail / singleton (posting_reader.rs:308-333, :545-547): lane 0 decodes
            const TermDev td = seg.terms[tc.term_id];
            n_in = td.tail_n;
            if (n_in == 0) {
                if (lane == 0) tc.cur = tc.nb + 1;
                __syncwarp();
                return false;
            }
            if (lane == 0) {
                int32_t* fq = reinterpret_cast<int32_t*>(cs);
                decode_tail(seg, td, cd, fq);
            }
            __syncwarp();
            const int i0 = 4 * lane;
            const int32_t* fq = reinterpret_cast<const int32_t*>(cs);
            docs = make_int4(i0 < (int)n_in ? cd[i0] : kNoMoreDocs, i0 + 1 < (int)n_in ? cd[i0 + 1] : kNoMoreDocs,
                             i0 + 2 < (int)n_in ? cd[i0 + 2] : kNoMoreDocs, i0 + 3 < (int)n_in ? cd[i0 + 3] : kNoMoreDocs);
            freqs = make_int4(i0 < (int)n_in ? fq[i0] : 1, i0 + 1 < (int)n_in ? fq[i0 + 1] : 1,
                              i0 + 2 < (int)n_in ? fq[i0 + 2] : 1, i0 + 3 < (int)n_in ? fq[i0 + 3] : 1);
            __syncwarp();
        }
        const int d[4] = {docs.x, docs.y, docs.z, docs.w};
        const int f[4] = {freqs.x, freqs.y, freqs.z, freqs.w};
        float sc[4];
        uint32_t below = 0, inside = 0, direct = 0;
        const float w1 = tc.w1;
        const float* cache = tc.cache;
        const uint8_t* norms = seg.norms;
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const bool ok = d[q] >= lo && d[q] < hi;
            below += d[q] < lo;
            inside += ok;
            float s = 0.f;
            if (ok) {
                const float nrm = norms ? __ldg(cache + __ldg(norms + d[q])) : p.k1;
                s = bm25_score(w1, (float)f[q], nrm);
                if (d[q] < win1) {  // still inside the window being drained: accumulate now
                    const int idx = d[q] - win0;
                    const uint32_t old = acc[idx];
                    const float sum = __fadd_rn(old == kSent ? 0.0f : __uint_as_float(old), s);
                    acc[idx] = __float_as_uint(sum);
                    touched |= 1u << (idx >> 5);
                    if (old == kSent && is_live(seg, d[q])) my_matches++;  // first clause on this doc
                    if (sum > te) hot |= 1u << (idx >> 5);                // may still enter the heap
                    direct++;
                }
            }
            sc[q] = s;
        }
        reinterpret_cast<int4*>(cd)[lane] = docs;
        reinterpret_cast<float4*>(cs)[lane] = make_float4(sc[0], sc[1], sc[2], sc[3]);
        below = __reduce_add_sync(0xffffffffu, below);
        inside = __reduce_add_sync(0xffffffffu, inside);
        direct = __reduce_add_sync(0xffffffffu, direct);
        const bool past_end = below + inside < n_in;  // some posting >= hi: nothing further in range
        below += direct;
        inside -= direct;
        if (lane == 0) {
            tc.pos = below;
            tc.n = below + inside;
            tc.cur = past_end ? tc.nb + 1 : b + 1;
        }
        __syncwarp();
        if (inside > 0) return true;
        if (past_end) return false;
        // whole block consumed (all below lo, or all accumulated directly): decode the next one
    }
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
    float* cscores = reinterpret_cast<float*>(cdocs + T * kBlock);

    for (int i = lane; i < kWw; i += 32) sh.acc[i] = kSent;
    if (lane < T) {
        const ItemClause c = p.clauses[it.clause_begin + lane];
        const TermDev td = seg.terms[c.term_id];
        WTerm& tc = sh.term[lane];
        tc.blk_last = seg.blk_last + td.blk_begin;
        tc.blk_desc = seg.blk_desc + td.blk_begin;
        tc.cache = p.caches + (size_t)c.cache_id * 256;
        tc.nb = td.n_blocks;
        tc.cur = lower_bound_i32(tc.blk_last, 0, td.n_blocks, lo);
        tc.n = 0;
        tc.pos = 0;
        tc.term_id = c.term_id;
        tc.w1 = __fmul_rn(c.weight, __fadd_rn(p.k1, 1.0f));
    }
    __syncwarp();
    long long w0 = kNoMoreDocs;
    uint32_t touched = 0, hot = 0, my_matches = 0;
    for (int t = 0; t < T; t++) {
        if (stream_refill(seg, p, sh.term[t], cdocs + t * kBlock, cscores + t * kBlock, lo, hi, lane, 0,
                          -2147483647 - 1, sh.acc, touched, hot, my_matches, INFINITY))
            w0 = min(w0, (long long)cdocs[t * kBlock + sh.term[t].pos]);
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
    // theta look-back: the up-to-32 preceding items of this heap chain (each publishes
    // max(own, inherited)), re-read every 8 windows
    const bool lb_ok = (uint32_t)lane < it.chain_pos;
    const uint32_t* theta_lb = p.item_theta + item_idx - 1 - (lb_ok ? lane : 0);
    uint32_t win_no = 0;

    while (w0 < hi) {
        const int win0 = (int)w0;
        const int win1 = (int)min((long long)hi, w0 + kWw);
        uint32_t inherited = 0;
        if ((win_no++ & 7u) == 0 && it.chain_pos) {
            inherited = lb_ok ? ld_volatile_u32(theta_lb) : 0u;
            inherited = __reduce_max_sync(0xffffffffu, inherited);
        }
        int next_doc = kNoMoreDocs;
        touched = 0;
        hot = 0;
        em.theta_in = max(em.theta_in, inherited);
        float te = em.theta_local;
        if (em.theta_in > kOrderedNegInf) te = fmaxf(te, ordered_to_float(em.theta_in));
        const bool open = te == -INFINITY;
        // ---- clauses in order: drain each stream up to the window end
        for (int t = 0; t < T; t++) {
            WTerm& tc = sh.term[t];
            const int32_t* cd = cdocs + t * kBlock;
            const float* cs = cscores + t * kBlock;
            uint32_t pos = tc.pos, n = tc.n;
            for (;;) {
                if (pos >= n) {
                    if (tc.cur > tc.nb) break;  // exhausted
                    if (lane == 0) tc.pos = pos;
                    __syncwarp();
                    if (!stream_refill(seg, p, tc, cdocs + t * kBlock, cscores + t * kBlock, lo, hi, lane, win0,
                                       win1, sh.acc, touched, hot, my_matches, te)) {
                        pos = n = 0;
                        break;
                    }
                    pos = tc.pos;
                    n = tc.n;
                }
                const uint32_t i = pos + lane;
                const int d = i < n ? cd[i] : kNoMoreDocs;
                const bool in_win = d < win1;
                const uint32_t c = __popc(__ballot_sync(0xffffffffu, in_win));  // sorted: a prefix
                if (in_win) {
                    const int idx = d - win0;
                    const uint32_t old = sh.acc[idx];
                    const float sum = __fadd_rn(old == kSent ? 0.0f : __uint_as_float(old), cs[i]);
                    sh.acc[idx] = __float_as_uint(sum);
                    touched |= 1u << (idx >> 5);
                    if (old == kSent && is_live(seg, d)) my_matches++;
                    if (sum > te) hot |= 1u << (idx >> 5);
                }
                pos += c;
                if (c < 32 && pos < n) break;  // next cached doc is beyond this window
            }
            if (lane == 0) {
                tc.pos = pos;
                tc.n = n;
            }
            if (pos < n) next_doc = min(next_doc, cd[pos]);
            __syncwarp();
        }
        touched = __reduce_or_sync(0xffffffffu, touched);
        hot = __reduce_or_sync(0xffffffffu, hot);
        // ---- window epilogue.  Matches were counted when a doc was first touched; only 32-doc
        // steps holding a doc whose (partial) sum exceeded theta are scanned for candidates, the
        // rest of the touched steps are just re-armed.
        {
            uint32_t cold = touched & ~hot;
            while (cold) {
                const int s = __ffs(cold) - 1;
                cold &= cold - 1;
                sh.acc[s * 32 + lane] = kSent;
            }
            uint32_t newc_n = 0;
            while (hot) {
                const int s = __ffs(hot) - 1;
                hot &= hot - 1;
                const int idx = s * 32 + lane;
                const uint32_t v = sh.acc[idx];
                sh.acc[idx] = kSent;
                const float sc = __uint_as_float(v);
                const bool cand = v != kSent && (open || sc > te) && is_live(seg, win0 + idx);
                const uint32_t cm = __ballot_sync(0xffffffffu, cand);
                if (!cm || em.overflow) continue;
                const uint32_t c = __popc(cm);
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
                    const uint32_t r = __popc(cm & ((1u << lane) - 1u));
                    p.cand_arena[em.run_slot + 1 + em.run_cnt + r] = rg_hit{win0 + idx + seg.doc_base, sc};
                    if (newc_n + r < (uint32_t)kNewcW) sh.newc[newc_n + r] = sc;
                }
                em.run_cnt += c;
                newc_n += c;
                if (lane == 0) hdr[em.run_slot] = CandRun{kNone, em.run_cnt};
            }
            wtheta_update(em, p.k, kcap, lane, sh.newc, newc_n, p.item_theta + item_idx);
            __syncwarp();
        }
        if (next_doc == kNoMoreDocs) break;
        w0 = next_doc;
    }
    my_matches = __reduce_add_sync(0xffffffffu, my_matches);
    if (lane == 0) p.item_matches[item_idx] = my_matches;
}

#endif
