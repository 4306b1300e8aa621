// unpack.cuh — ForUtil block decode on the device (pure integer work, no tensor cores).
//
// One warp decodes one 128-value block; lane q produces values n = 4q .. 4q+3, so the decoded
// block leaves the warp as one coalesced 512-byte store (or stays in registers for the fused
// query kernels).  Payload parts are 16-byte aligned in the HBM image (common.cuh), which makes
// every load a 128-bit (SIMD128 layout) or aligned 32-bit (big-endian layouts) access.
//
// Reference semantics (paths relative to /root/reference/src/core/):
//   SIMD128 (".doc" version 1): util/packed/packed_simd.rs:126-163 unpack_bits!
//       value n -> lane l = n%4, lane-stream slot q = n/4; v = (W[j][l] >> s | W[j+1][l] << (32-s))
//       & mask with j = q*b/32, s = q*b%32 over little-endian words.
//   Packed (version 0): util/packed/packed_misc.rs:2655-2680 — one MSB-first big-endian stream,
//       value n at bit n*b.
//   PackedSingleBlock (version 0, b in {1,2,4} under COMPACT): packed_misc.rs:2829-2841,
//       2739-2755 — 64/b values per big-endian long, value i of a long at bits [i*b, i*b+b).
//   all-equal (code 0): codec/postings/for_util.rs:210-221.
#pragma once
#include "common.cuh"

namespace rg {

__device__ __forceinline__ uint32_t bswap32(uint32_t x) { return __byte_perm(x, 0, 0x0123); }

// ---- SIMD128 ----------------------------------------------------------------------------
__device__ __forceinline__ int4 unpack4_simd128(const uint4* __restrict__ part, int b, int q) {
    if (b == 32) {
        uint4 v = ldg16(part + q);
        return make_int4((int)v.x, (int)v.y, (int)v.z, (int)v.w);
    }
    const int bit = q * b;
    const int j = bit >> 5, s = bit & 31;
    const int j2 = min(j + 1, b - 1);  // stays inside the payload; masked out when unused
    const uint4 A = ldg16(part + j);
    const uint4 B = ldg16(part + j2);
    const uint32_t mask = (1u << b) - 1u;
    int4 r;
    r.x = (int)(__funnelshift_r(A.x, B.x, s) & mask);
    r.y = (int)(__funnelshift_r(A.y, B.y, s) & mask);
    r.z = (int)(__funnelshift_r(A.z, B.z, s) & mask);
    r.w = (int)(__funnelshift_r(A.w, B.w, s) & mask);
    return r;
}
// one value (random access inside a block, used for lazy freq lookups)
__device__ __forceinline__ int extract_simd128(const uint4* __restrict__ part, int b, int n) {
    const uint32_t* w = reinterpret_cast<const uint32_t*>(part);
    const int lane = n & 3, q = n >> 2;
    if (b == 32) return (int)__ldg(w + q * 4 + lane);
    const int bit = q * b;
    const int j = bit >> 5, s = bit & 31;
    const int j2 = min(j + 1, b - 1);
    const uint32_t lo = __ldg(w + j * 4 + lane), hi = __ldg(w + j2 * 4 + lane);
    return (int)(__funnelshift_r(lo, hi, s) & ((1u << b) - 1u));
}

// ---- Packed (big-endian bit stream) ------------------------------------------------------
__device__ __forceinline__ int extract_packed_be(const uint32_t* __restrict__ w, int b, int n) {
    const int bit = n * b;
    const int wi = bit >> 5, s = bit & 31;
    const int wi2 = min(wi + 1, 4 * b - 1);
    const uint32_t hi = bswap32(__ldg(w + wi)), lo = bswap32(__ldg(w + wi2));
    return (int)(__funnelshift_l(lo, hi, s) >> (32 - b));
}
// ---- PackedSingleBlock -------------------------------------------------------------------
__device__ __forceinline__ int extract_single_block(const uint32_t* __restrict__ w, int b, int n) {
    const int per = 64 / b;
    const int L = n / per, i = n - L * per;
    const uint32_t hi = bswap32(__ldg(w + 2 * L)), lo = bswap32(__ldg(w + 2 * L + 1));
    const int sh = i * b;
    const uint32_t mask = b == 32 ? 0xffffffffu : ((1u << b) - 1u);
    const uint32_t v = sh < 32 ? __funnelshift_r(lo, hi, sh) : (hi >> (sh - 32));
    return (int)(v & mask);
}

// ---- generic entry points -----------------------------------------------------------------
// `part`: 16-byte aligned block part; b: num_bits from the block header (0 = all equal).
__device__ __forceinline__ int4 unpack4(const uint4* __restrict__ part, int b, int q, int version,
                                        uint32_t sb_mask) {
    if (b == 0) {
        const int v = (int)__ldg(reinterpret_cast<const uint32_t*>(part));
        return make_int4(v, v, v, v);
    }
    if (version > 0) return unpack4_simd128(part, b, q);
    const uint32_t* w = reinterpret_cast<const uint32_t*>(part);
    int4 r;
    if ((sb_mask >> (b - 1)) & 1u) {
        r.x = extract_single_block(w, b, 4 * q);
        r.y = extract_single_block(w, b, 4 * q + 1);
        r.z = extract_single_block(w, b, 4 * q + 2);
        r.w = extract_single_block(w, b, 4 * q + 3);
    } else {
        r.x = extract_packed_be(w, b, 4 * q);
        r.y = extract_packed_be(w, b, 4 * q + 1);
        r.z = extract_packed_be(w, b, 4 * q + 2);
        r.w = extract_packed_be(w, b, 4 * q + 3);
    }
    return r;
}
__device__ __forceinline__ int extract1(const uint4* __restrict__ part, int b, int n, int version,
                                        uint32_t sb_mask) {
    if (b == 0) return (int)__ldg(reinterpret_cast<const uint32_t*>(part));
    if (version > 0) return extract_simd128(part, b, n);
    const uint32_t* w = reinterpret_cast<const uint32_t*>(part);
    return ((sb_mask >> (b - 1)) & 1u) ? extract_single_block(w, b, n) : extract_packed_be(w, b, n);
}

// Doc-delta block -> absolute docids: the reference accumulates one delta per next()
// (codec/postings/posting_reader.rs:622-640); here a warp-wide inclusive scan, base = last doc
// of the previous block (0 for a term's first block).
__device__ __forceinline__ int4 deltas_to_docs(int4 d, int base) {
    int4 p;
    p.x = d.x;
    p.y = p.x + d.y;
    p.z = p.y + d.z;
    p.w = p.z + d.w;
    int tot = p.w;
    const int lane = lane_id();
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, tot, o);
        if (lane >= o) tot += t;
    }
    const int excl = tot - p.w + base;
    p.x += excl;
    p.y += excl;
    p.z += excl;
    p.w += excl;
    return p;
}

// vint / vlong readers over a byte pointer (store/io/data_input.rs:78-111)
__device__ __forceinline__ int read_vint(const uint8_t* __restrict__ p, uint32_t& pos) {
    uint32_t b = p[pos++];
    uint32_t v = b & 0x7f;
    int shift = 7;
    while (b & 0x80) {
        b = p[pos++];
        v |= (b & 0x7f) << shift;
        shift += 7;
    }
    return (int)v;
}

}  // namespace rg
