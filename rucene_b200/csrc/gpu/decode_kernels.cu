// decode_kernels.cu — ForUtil 128-int block decode microbench kernels (BASELINE config 2).
//
// k_decode_staged : blocks resident in HBM in 16-byte aligned parts (the engine's index image
//                   layout).  HBM-bound: reads 16*b bytes, writes 512 bytes per block.
// k_decode_raw    : blocks decoded in place from the raw codec stream (1 header byte + payload
//                   at an arbitrary byte alignment) — what rg_forutil_decode() runs on the bytes
//                   a caller hands over; realigns with funnel shifts over aligned 32-bit loads.
#include "engine.hpp"
#include "unpack.cuh"

namespace rg {

constexpr int kDecodeUnroll = 4;   // blocks per warp iteration: 8 x 16-byte loads in flight per lane
constexpr int kDecodeThreads = 256;

__global__ void __launch_bounds__(kDecodeThreads)
k_decode_staged(const uint4* __restrict__ arena, const BlockDesc* __restrict__ desc,
                uint32_t n_blocks, int4* __restrict__ out, int version, uint32_t sb_mask) {
    const int lane = lane_id();
    const uint32_t warp = (blockIdx.x * kDecodeThreads + threadIdx.x) >> 5;
    const uint32_t first = warp * kDecodeUnroll;
    if (first >= n_blocks) return;
    BlockDesc d[kDecodeUnroll];
#pragma unroll
    for (int u = 0; u < kDecodeUnroll; u++) {
        uint32_t i = min(first + u, n_blocks - 1);
        d[u] = desc[i];
    }
    int4 v[kDecodeUnroll];
#pragma unroll
    for (int u = 0; u < kDecodeUnroll; u++)
        v[u] = unpack4(arena + d[u].off16, (int)(d[u].bits & 0xff), lane, version, sb_mask);
#pragma unroll
    for (int u = 0; u < kDecodeUnroll; u++)
        if (first + u < n_blocks) stg16_streaming(out + (size_t)(first + u) * 32 + lane, v[u]);
}

// Every block pair [doc-delta block][freq block] of an uploaded segment, in file order (BASELINE config 2,
// "realistic" value set): what ForUtil::read_block + read_block returns per pair (for_util.rs:187-243;
// the doc part stays deltas).  out: n_blocks x (128 deltas, 128 freqs).  EF / BITSET doc parts are skipped
// (they are not ForUtil-packed); their 128 output slots are left untouched.
__global__ void __launch_bounds__(kDecodeThreads)
k_decode_segment(const uint4* __restrict__ arena, const BlockDesc* __restrict__ desc, uint32_t first,
                 uint32_t n_blocks, int4* __restrict__ out, int version, uint32_t sb_mask) {
    const int lane = lane_id();
    const uint32_t warp = (blockIdx.x * kDecodeThreads + threadIdx.x) >> 5;
    const uint32_t b0 = warp * 2;
    if (b0 >= n_blocks) return;
    BlockDesc d[2];
#pragma unroll
    for (int u = 0; u < 2; u++) d[u] = desc[first + min(b0 + u, n_blocks - 1)];
    int4 dv[2], fv[2];
#pragma unroll
    for (int u = 0; u < 2; u++) {
        const uint4* part = arena + d[u].off16;
        dv[u] = (d[u].bits >> 24) ? make_int4(0, 0, 0, 0) : unpack4(part, (int)(d[u].bits & 0xff), lane, version, sb_mask);
        fv[u] = unpack4(part + ((d[u].bits >> 16) & 0xff), (int)((d[u].bits >> 8) & 0xff), lane, version, sb_mask);
    }
#pragma unroll
    for (int u = 0; u < 2; u++) {
        if (b0 + u >= n_blocks) break;
        int4* o = out + (size_t)(b0 + u) * 64 + lane;
        if (!(d[u].bits >> 24)) stg16_streaming(o, dv[u]);
        stg16_streaming(o + 32, fv[u]);
    }
}

// ---- in-place decode of the raw stream ----------------------------------------------------
// 32-bit little-endian word at an arbitrary byte address, from two aligned loads.
__device__ __forceinline__ uint32_t load_u32_unaligned(const uint8_t* __restrict__ p) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const uint32_t* w = reinterpret_cast<const uint32_t*>(a & ~uintptr_t(3));
    const int sh = (int)(a & 3) * 8;
    const uint32_t lo = __ldg(w);
    if (sh == 0) return lo;
    return __funnelshift_r(lo, __ldg(w + 1), sh);
}

__device__ __forceinline__ int4 unpack4_raw(const uint8_t* __restrict__ payload, int b, int q,
                                            int version, uint32_t sb_mask, const int* enc_size) {
    int4 r;
    if (version > 0) {
        if (b == 32) {
            r.x = (int)load_u32_unaligned(payload + 16 * q);
            r.y = (int)load_u32_unaligned(payload + 16 * q + 4);
            r.z = (int)load_u32_unaligned(payload + 16 * q + 8);
            r.w = (int)load_u32_unaligned(payload + 16 * q + 12);
            return r;
        }
        const int bit = q * b;
        const int j = bit >> 5, s = bit & 31;
        const int j2 = min(j + 1, b - 1);
        const uint32_t mask = (1u << b) - 1u;
        const uint8_t* pa = payload + 16 * j;
        const uint8_t* pb = payload + 16 * j2;
        r.x = (int)(__funnelshift_r(load_u32_unaligned(pa), load_u32_unaligned(pb), s) & mask);
        r.y = (int)(__funnelshift_r(load_u32_unaligned(pa + 4), load_u32_unaligned(pb + 4), s) & mask);
        r.z = (int)(__funnelshift_r(load_u32_unaligned(pa + 8), load_u32_unaligned(pb + 8), s) & mask);
        r.w = (int)(__funnelshift_r(load_u32_unaligned(pa + 12), load_u32_unaligned(pb + 12), s) & mask);
        return r;
    }
    (void)enc_size;
    int vals[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int n = 4 * q + k;
        uint32_t hi, lo;
        int v;
        if ((sb_mask >> (b - 1)) & 1u) {
            const int per = 64 / b;
            const int L = n / per, i = n - L * per;
            hi = bswap32(load_u32_unaligned(payload + 8 * L));
            lo = bswap32(load_u32_unaligned(payload + 8 * L + 4));
            const int sh = i * b;
            const uint32_t mask = b == 32 ? 0xffffffffu : ((1u << b) - 1u);
            v = (int)((sh < 32 ? __funnelshift_r(lo, hi, sh) : (hi >> (sh - 32))) & mask);
        } else {
            const int bit = n * b;
            const int wi = bit >> 5, s = bit & 31;
            const int wi2 = min(wi + 1, 4 * b - 1);
            hi = bswap32(load_u32_unaligned(payload + 4 * wi));
            lo = bswap32(load_u32_unaligned(payload + 4 * wi2));
            v = (int)(__funnelshift_l(lo, hi, s) >> (32 - b));
        }
        vals[k] = v;
    }
    return make_int4(vals[0], vals[1], vals[2], vals[3]);
}

__global__ void __launch_bounds__(kDecodeThreads)
k_decode_raw(const uint8_t* __restrict__ stream, const uint64_t* __restrict__ offsets,
             uint32_t n_blocks, int4* __restrict__ out, int version, uint32_t sb_mask) {
    const int lane = lane_id();
    const uint32_t blk = (blockIdx.x * kDecodeThreads + threadIdx.x) >> 5;
    if (blk >= n_blocks) return;
    const uint8_t* p = stream + offsets[blk];
    const int b = p[0] & 0x3f;  // ForUtil::read_block: low six bits = num_bits
    int4 v;
    if (b == 0) {
        uint32_t pos = 1;
        const int x = read_vint(p, pos);
        v = make_int4(x, x, x, x);
    } else {
        v = unpack4_raw(p + 1, b, lane, version, sb_mask, nullptr);
    }
    stg16_streaming(out + (size_t)blk * 32 + lane, v);
}

void launch_decode_staged(cudaStream_t st, const uint4* arena, const BlockDesc* desc,
                          uint32_t n_blocks, int32_t* out, int version, uint32_t sb_mask) {
    if (n_blocks == 0) return;
    const uint32_t warps = (n_blocks + kDecodeUnroll - 1) / kDecodeUnroll;
    const uint32_t ctas = (warps + (kDecodeThreads / 32) - 1) / (kDecodeThreads / 32);
    k_decode_staged<<<ctas, kDecodeThreads, 0, st>>>(arena, desc, n_blocks,
                                                     reinterpret_cast<int4*>(out), version, sb_mask);
}

void launch_decode_segment(cudaStream_t st, const uint4* arena, const BlockDesc* desc, uint32_t first,
                           uint32_t n_blocks, int32_t* out, int version, uint32_t sb_mask) {
    if (n_blocks == 0) return;
    const uint32_t warps = (n_blocks + 1) / 2;
    const uint32_t ctas = (warps + (kDecodeThreads / 32) - 1) / (kDecodeThreads / 32);
    k_decode_segment<<<ctas, kDecodeThreads, 0, st>>>(arena, desc, first, n_blocks, reinterpret_cast<int4*>(out),
                                                      version, sb_mask);
}

void launch_decode_raw(cudaStream_t st, const uint8_t* stream, const uint64_t* offsets,
                       uint32_t n_blocks, int32_t* out, int version, uint32_t sb_mask) {
    if (n_blocks == 0) return;
    const uint32_t ctas = (n_blocks + (kDecodeThreads / 32) - 1) / (kDecodeThreads / 32);
    k_decode_raw<<<ctas, kDecodeThreads, 0, st>>>(stream, offsets, n_blocks,
                                                  reinterpret_cast<int4*>(out), version, sb_mask);
}

}  // namespace rg
