// Shared pieces of the fused SwinUNet token kernels (swin_mlp_sm100.cu, swin_attn_sm100.cu): small shared-memory / conversion helpers and
// the LayerNorm producer that turns token rows into the K-major SWIZZLE_64B A operand of a tcgen05 GEMM.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdint>

#include "launch.h"
#include "sm100_common.cuh"

namespace w2x {
namespace swintok {
using namespace sm100;

constexpr int kRows = 128;                  // tokens per tile = UMMA M
constexpr uint32_t kAChunk = kRows * 64;    // one 32-channel K chunk of the A operand: [128 rows][32 k] fp16, SWIZZLE_64B

__device__ __forceinline__ void stsF32(uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }
__device__ __forceinline__ void unpack8(const uint4& raw, float (&v)[8]) {
    const __half2* h = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 f = __half22float2(h[i]);
        v[2 * i] = f.x;
        v[2 * i + 1] = f.y;
    }
}
__device__ __forceinline__ void loadF8(uint32_t addr, float (&v)[8]) {
    const uint4 a = ldsV4(addr), b = ldsV4(addr + 16u);
    v[0] = __uint_as_float(a.x); v[1] = __uint_as_float(a.y); v[2] = __uint_as_float(a.z); v[3] = __uint_as_float(a.w);
    v[4] = __uint_as_float(b.x); v[5] = __uint_as_float(b.y); v[6] = __uint_as_float(b.z); v[7] = __uint_as_float(b.w);
}

// LayerNorm producer warps (groups of 4 warps, one token row per thread): rows of tile k -> fp16 A-operand rows (SWIZZLE_64B,
// 32-channel K chunks of kAChunk bytes) in buffer k % kABufs at offA; tokenOf(k, row) names the token a row holds (the fused attention
// kernel gathers window-ordered tokens this way: roll + window partition are its address math); barAFull / barAEmpty are the shared-memory addresses of buffer
// 0's barriers.  beforeTile(k) / afterTile(k) run before tile k's rows are loaded / after they have been handed over; kFirst / kStep let several groups of four warps share the tiles (measured: a second group does not pay, the 96
// registers per thread it leaves cost more than the overlap gains).
struct NoTileHook {
    __device__ __forceinline__ void operator()(int) const {}
};
template <int C, int kABufs, typename TokenOf, typename AfterTile = NoTileHook, typename BeforeTile = NoTileHook>
__device__ __forceinline__ void lnProducerLoop(const __half* x, float eps, TokenOf tokenOf, uint32_t base, uint32_t offA, uint32_t offGamma, uint32_t offBeta,
                                               uint32_t barAFull, uint32_t barAEmpty, int nMine, int kFirst = 0, int kStep = 1, AfterTile afterTile = AfterTile(),
                                               BeforeTile beforeTile = BeforeTile()) {
    const int lane = threadIdx.x & 31;
    pdlWait();  // x is written by the preceding kernel
    const int row = threadIdx.x & 127;
    const uint32_t sw = (uint32_t)(row >> 1) & 3u;
    // normalise 12 pieces (96 channels) held in registers and store them as A-operand rows; J0 = index of the first piece
    auto emit = [&](const uint4 (&raw)[12], int J0, float mean, float rstd, bool valid, uint32_t rowAddr) {
#pragma unroll
        for (int j = 0; j < 12; ++j) {
            const int J = J0 + j;
            float v[8], gm[8], bt[8];
            unpack8(raw[j], v);
            loadF8(base + offGamma + 32u * J, gm);
            loadF8(base + offBeta + 32u * J, bt);
            uint4 o;
            __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
            for (int i = 0; i < 4; ++i)
                oh[i] = __floats2half2_rn((v[2 * i] - mean) * rstd * gm[2 * i] + bt[2 * i], (v[2 * i + 1] - mean) * rstd * gm[2 * i + 1] + bt[2 * i + 1]);
            if (!valid) o = make_uint4(0, 0, 0, 0);
            stsV4(rowAddr + (uint32_t)(J >> 2) * kAChunk + ((((uint32_t)J & 3u) ^ sw) << 4), o);
        }
    };
    for (int k = kFirst; k < nMine; k += kStep) {
        beforeTile(k);   // per-tile work that must not sit behind this tile's wait for a free A buffer
        const long long g = tokenOf(k, row);   // token index of row `row` of this CTA's k-th tile, or < 0 for a padding row
        const bool valid = g >= 0;
        const uint4* src = reinterpret_cast<const uint4*>(x + (valid ? g : 0) * C);
        const int buf = kABufs == 2 ? (k & 1) : 0;
        const uint32_t use = kABufs == 2 ? (uint32_t)(k >> 1) : (uint32_t)k;
        const uint32_t rowAddr = base + offA + (uint32_t)buf * (uint32_t)(C / 32) * kAChunk + (uint32_t)row * 64u;
        uint4 raw[12];
        if constexpr (C == 96) {
            // the whole row stays in registers: two-pass statistics
#pragma unroll
            for (int j = 0; j < 12; ++j) raw[j] = make_uint4(0, 0, 0, 0);
            if (valid) {
#pragma unroll
                for (int j = 0; j < 12; ++j) raw[j] = src[j];
            }
            float ps[4] = {0.f, 0.f, 0.f, 0.f};   // four partial sums: the reductions are not one 96-long dependent chain
#pragma unroll
            for (int j = 0; j < 12; ++j) {
                float v[8];
                unpack8(raw[j], v);
#pragma unroll
                for (int i = 0; i < 8; ++i) ps[i & 3] += v[i];
            }
            const float mean = ((ps[0] + ps[1]) + (ps[2] + ps[3])) * (1.f / C);
            float pq[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int j = 0; j < 12; ++j) {
                float v[8];
                unpack8(raw[j], v);
#pragma unroll
                for (int i = 0; i < 8; ++i) { const float d = v[i] - mean; pq[i & 3] = fmaf(d, d, pq[i & 3]); }
            }
            const float rstd = rsqrtf(((pq[0] + pq[1]) + (pq[2] + pq[3])) * (1.f / C) + eps);
            mbarWait(barAEmpty + 8u * buf, (use & 1u) ^ 1u);   // the last fc1 chunk that read this buffer has completed
            emit(raw, 0, mean, rstd, valid, rowAddr);
        } else {
            // 384-byte rows do not fit the register budget next to the epilogue warps: statistics in one streaming pass (sums of
            // x - x0 and (x - x0)^2 with x0 = the row's first element, so a large common offset cannot cancel), then the row
            // is read again (an L2 hit) 96 channels at a time for the normalisation
            float x0 = 0.f, p1[4] = {0.f, 0.f, 0.f, 0.f}, p2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int h = 0; h < (C / 8) / 12; ++h) {
#pragma unroll
                for (int j = 0; j < 12; ++j) raw[j] = make_uint4(0, 0, 0, 0);
                if (valid) {
#pragma unroll
                    for (int j = 0; j < 12; ++j) raw[j] = src[12 * h + j];
                }
                if (h == 0) x0 = __low2float(*reinterpret_cast<const __half2*>(&raw[0]));
#pragma unroll
                for (int j = 0; j < 12; ++j) {
                    float v[8];
                    unpack8(raw[j], v);
#pragma unroll
                    for (int i = 0; i < 8; ++i) { const float d = v[i] - x0; p1[i & 3] += d; p2[i & 3] = fmaf(d, d, p2[i & 3]); }
                }
            }
            const float s1 = (p1[0] + p1[1]) + (p1[2] + p1[3]), s2 = (p2[0] + p2[1]) + (p2[2] + p2[3]);
            const float m1 = s1 * (1.f / C);
            const float mean = x0 + m1;
            const float rstd = rsqrtf(fmaxf(s2 * (1.f / C) - m1 * m1, 0.f) + eps);
            if (valid) {
#pragma unroll
                for (int j = 0; j < 12; ++j) raw[j] = src[j];   // in flight while the buffer is still being read by fc1
            }
            mbarWait(barAEmpty + 8u * buf, (use & 1u) ^ 1u);
#pragma unroll
            for (int h = 0; h < (C / 8) / 12; ++h) {
                uint4 nxt[12];
                if (h + 1 < (C / 8) / 12 && valid) {
#pragma unroll
                    for (int j = 0; j < 12; ++j) nxt[j] = src[12 * (h + 1) + j];
                }
                emit(raw, 12 * h, mean, rstd, valid, rowAddr);
                if (h + 1 < (C / 8) / 12) {
#pragma unroll
                    for (int j = 0; j < 12; ++j) raw[j] = nxt[j];
                }
            }
        }
        fenceProxyAsync();
        __syncwarp();
        if (lane == 0) mbarArrive(barAFull + 8u * buf);
        afterTile(k);   // further per-tile work of the producer warps (the fused attention kernel writes back the previous tile here)
    }
}

}  // namespace swintok
}  // namespace w2x
