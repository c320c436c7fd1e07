// Fused MLP half of a SwinUNet block (torchvision SwinTransformerBlock: x = x + fc2(GELU(fc1(LayerNorm(x))))), one kernel
// instead of layernorm_kernel + two igemm_kernel launches.  Per 128-token tile the token rows are read from HBM once and written
// once; the normalised rows, the hidden tensor (2C wide) and both weight matrices never leave the SM:
//
//   warps 0-3   producers: one token row per thread (192 B, twelve 16-byte loads in flight), LayerNorm in registers (two-pass, fp32),
//               fp16 rows written into shared memory in the SWIZZLE_64B K-major layout of a UMMA A operand (three 32-channel chunks)
//   warp  4     MMA issuer (one elected lane): fc1 = 6 x tcgen05.mma M128 N192 K16 into a double-buffered TMEM accumulator,
//               fc2 = 12 x M128 N96 K16 whose A operand is the hidden tile the epilogue warps wrote; fc1 of tile k+1 is queued
//               before fc2 of tile k, so the tensor pipe works while the CUDA cores evaluate GELU
//   warps 5-12  epilogue (two warps per TMEM lane quarter, splitting the columns): phase 1 = tcgen05.ld, + bias, erf-GELU, fp16,
//               written as the SWIZZLE_128B A operand of fc2; phase 2 = tcgen05.ld, + bias + residual (prefetched from x), 16-byte
//               stores to x in place.  Phase 1 of tile k+1 runs before phase 2 of tile k (fc2 of tile k executes meanwhile).
//
// Both weight matrices (2 x 36 KB for C = 96) are TMA-loaded once per CTA and stay resident.  The kernel is bound by the GELU
// evaluation on the CUDA cores (128 x 192 values per tile), not by HBM or the tensor pipe.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <string>

#include "conv_params.h"
#include "launch.h"
#include "sm100_common.cuh"
#include "swin_token.cuh"

namespace w2x {
using namespace sm100;
using namespace swintok;
namespace {

constexpr int kC = 96;             // token width
constexpr int kHid = 192;          // hidden width (mlp_ratio 2)
constexpr int kProdWarps = 4;
constexpr int kMmaWarp = 4;
constexpr int kEpiWarp0 = 5;
constexpr int kEpiWarps = 8;
constexpr int kMlpThreads = 32 * (kEpiWarp0 + kEpiWarps);   // 416

// shared-memory map (byte offsets from the 1024-aligned base)
constexpr uint32_t kBarW = 0, kBarAFull = 8, kBarAEmpty = 24, kBarD1Full = 40, kBarD1Empty = 56, kBarHFull = 72, kBarHEmpty = 88, kBarD2Full = 104,
                   kBarD2Empty = 112, kTmemSlot = 120;
constexpr uint32_t kOffB1 = 256, kOffB2 = kOffB1 + kHid * 4, kOffGamma = kOffB2 + kC * 4, kOffBeta = kOffGamma + kC * 4;
constexpr uint32_t kW1Chunk = kHid * 64;          // [192 rows][32 k] fp16, SWIZZLE_64B
constexpr uint32_t kW2Chunk = kC * 128;           // [96 rows][64 k] fp16, SWIZZLE_128B
constexpr uint32_t kHChunk = kRows * 128;         // [128 rows][64 k]
constexpr uint32_t kOffW1 = 4096;
constexpr uint32_t kOffW2 = kOffW1 + 3 * kW1Chunk;
constexpr uint32_t kOffA = kOffW2 + 3 * kW2Chunk;
constexpr uint32_t kOffH = kOffA + 2 * 3 * kAChunk;
constexpr uint32_t kMlpSmem = kOffH + 2 * 3 * kHChunk + 1024;   // + alignment slack
static_assert(kOffBeta + kC * 4 <= kOffW1, "constants overflow the header");
static_assert(kOffW2 % 1024 == 0 && kOffA % 1024 == 0 && kOffH % 1024 == 0, "swizzled operands need 1024-byte alignment");
static_assert(kMlpSmem <= 227 * 1024, "shared memory budget");
constexpr uint32_t kTmemD2 = 2 * kHid;            // columns: D1 buffers at 0 and 192, D2 at 384
constexpr uint32_t kTmemCols = 512;

struct MlpArgs {
    CUtensorMap tmW1, tmW2;
    __half* x;                 // [tokens][C] fp16, updated in place
    const float* b1;           // [2C]
    const float* b2;           // [C]
    const float* gamma;        // [C]
    const float* beta;         // [C]
    float eps;
    long long tokens;
};

// Every CTA owns one contiguous, balanced range of tokens [start, end) and walks it in tiles of 128; only its last tile is ragged.  (Dealing
// whole tiles round-robin left the CTAs that drew one tile more as the critical path: 450 tiles on 148 SMs cost four tile times instead of
// 3.04.)  The epilogue warps skip the GELU of a lane quarter that holds no valid row, so the ragged tile costs a fraction of a full one.
struct LinearTokens {
    long long start, end;
    __device__ __forceinline__ long long operator()(int k, int row) const {
        const long long g = start + (long long)k * kRows + row;
        return g < end ? g : -1;
    }
    __device__ __forceinline__ int tiles() const { return (int)((end - start + kRows - 1) / kRows); }
    __device__ __forceinline__ int validRows(int k) const {   // rows of tile k that hold a token
        const long long left = end - start - (long long)k * kRows;
        return left >= kRows ? kRows : (left > 0 ? (int)left : 0);
    }
};
__device__ __forceinline__ LinearTokens linearTokens(long long tokens) {
    const long long per = (tokens + gridDim.x - 1) / gridDim.x;
    const long long start = per * blockIdx.x < tokens ? per * blockIdx.x : tokens;
    return LinearTokens{start, start + per < tokens ? start + per : tokens};
}

__global__ void __launch_bounds__(kMlpThreads, 1) swin_mlp_kernel(const __grid_constant__ MlpArgs a) {
    extern __shared__ uint8_t smemRaw[];
    const uint32_t base = (smemU32(smemRaw) + 1023u) & ~1023u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    pdlLaunchDependents();

    // ---- prologue: constants only (weights, biases, LayerNorm parameters) ----
    if (threadIdx.x == 0) {
        mbarInit(base + kBarW, 1);
        for (int i = 0; i < 2; ++i) {
            mbarInit(base + kBarAFull + 8u * i, 4);
            mbarInit(base + kBarAEmpty + 8u * i, 1);
            mbarInit(base + kBarD1Full + 8u * i, 1);
            mbarInit(base + kBarD1Empty + 8u * i, kEpiWarps);
            mbarInit(base + kBarHFull + 8u * i, kEpiWarps);
            mbarInit(base + kBarHEmpty + 8u * i, 1);
        }
        mbarInit(base + kBarD2Full, 1);
        mbarInit(base + kBarD2Empty, kEpiWarps);
        mbarInitFence();
        tmaPrefetchDesc(&a.tmW1);
        tmaPrefetchDesc(&a.tmW2);
    }
    for (int i = threadIdx.x; i < kHid; i += kMlpThreads) stsF32(base + kOffB1 + 4u * i, a.b1[i]);
    for (int i = threadIdx.x; i < kC; i += kMlpThreads) {
        stsF32(base + kOffB2 + 4u * i, a.b2[i]);
        stsF32(base + kOffGamma + 4u * i, a.gamma[i]);
        stsF32(base + kOffBeta + 4u * i, a.beta[i]);
    }
    if (warp == kMmaWarp) tmemAlloc(base + kTmemSlot, kTmemCols);
    tcFenceBefore();
    __syncthreads();
    tcFenceAfter();
    uint32_t tmemBase;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmemBase) : "r"(base + kTmemSlot));
    if (warp == kMmaWarp && lane == 0) {
        mbarExpectTx(base + kBarW, 3 * kW1Chunk + 3 * kW2Chunk);
        for (int kc = 0; kc < 3; ++kc) {
            tmaLoad2d(base + kOffW1 + kc * kW1Chunk, &a.tmW1, base + kBarW, kc * 32, 0);
            tmaLoad2d(base + kOffW2 + kc * kW2Chunk, &a.tmW2, base + kBarW, kc * 64, 0);
        }
    }
    const LinearTokens tok = linearTokens(a.tokens);
    const int nMine = tok.tiles();
    pdlWait();  // x is written by the preceding kernel

    if (warp < kProdWarps) {
        // ---- LayerNorm producers: thread = token row ----
        lnProducerLoop<kC, 2>(a.x, a.eps, tok, base, kOffA, kOffGamma, kOffBeta, base + kBarAFull, base + kBarAEmpty, nMine);
    } else if (warp == kMmaWarp) {
        // ---- MMA issuer: whole warp converged, one elected lane issues ----
        const uint32_t hi64 = descHi(512u, 4u), hi128 = descHi(1024u, 2u);
        const uint32_t idesc1 = instrDescF16(kRows, kHid), idesc2 = instrDescF16(kRows, kC);
        auto fc1 = [&](int t) {
            const int buf = t & 1;
            const uint32_t ph = (uint32_t)(t >> 1) & 1u;
            mbarWait(base + kBarAFull + 8u * buf, ph);
            mbarWait(base + kBarD1Empty + 8u * buf, ph ^ 1u);
            tcFenceAfter();
            if (electOne()) {
                const uint32_t aBase = base + kOffA + (uint32_t)buf * 3u * kAChunk;
#pragma unroll
                for (int kc = 0; kc < 3; ++kc)
#pragma unroll
                    for (int ks = 0; ks < 2; ++ks)
                        ummaLoHi(tmemBase + (uint32_t)buf * kHid, descLo(aBase + kc * kAChunk + ks * 32u), hi64, descLo(base + kOffW1 + kc * kW1Chunk + ks * 32u), hi64,
                                 idesc1, (kc | ks) != 0 ? 1u : 0u);
                tcCommit(base + kBarD1Full + 8u * buf);
                tcCommit(base + kBarAEmpty + 8u * buf);
            }
            __syncwarp();
        };
        auto fc2 = [&](int t) {
            const int buf = t & 1;
            mbarWait(base + kBarHFull + 8u * buf, (uint32_t)(t >> 1) & 1u);
            mbarWait(base + kBarD2Empty, ((uint32_t)t & 1u) ^ 1u);
            tcFenceAfter();
            if (electOne()) {
                const uint32_t hBase = base + kOffH + (uint32_t)buf * 3u * kHChunk;
#pragma unroll
                for (int kc = 0; kc < 3; ++kc)
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)
                        ummaLoHi(tmemBase + kTmemD2, descLo(hBase + kc * kHChunk + ks * 32u), hi128, descLo(base + kOffW2 + kc * kW2Chunk + ks * 32u), hi128, idesc2,
                                 (kc | ks) != 0 ? 1u : 0u);
                tcCommit(base + kBarD2Full);
                tcCommit(base + kBarHEmpty + 8u * buf);
            }
            __syncwarp();
        };
        mbarWait(base + kBarW, 0);
        if (nMine > 0) fc1(0);
        for (int k = 0; k < nMine; ++k) {
            if (k + 1 < nMine) fc1(k + 1);
            fc2(k);
        }
    } else {
        // ---- epilogue warps ----
        const int quarter = warp & 3;                // TMEM lane quarter this warp may read: its index in the CTA modulo 4
        const int half = (warp - kEpiWarp0) >> 2;    // which half of the accumulator columns
        const int row = quarter * 32 + lane;
        const uint32_t taddrLane = tmemBase + ((uint32_t)(quarter * 32) << 16);
        const uint32_t sw = (uint32_t)row & 7u;
        uint32_t r[32];
        auto phase1 = [&](int t) {   // hidden = GELU(fc1 + b1) -> shared memory (A operand of fc2)
            const int buf = t & 1;
            const uint32_t ph = (uint32_t)(t >> 1) & 1u;
            mbarWait(base + kBarD1Full + 8u * buf, ph);
            tcFenceAfter();
            mbarWait(base + kBarHEmpty + 8u * buf, ph ^ 1u);   // fc2 of tile t-2 has consumed this hidden buffer
            const uint32_t hRow = base + kOffH + (uint32_t)buf * 3u * kHChunk + (uint32_t)row * 128u;
            const int nc = quarter * 32 < tok.validRows(t) ? 3 : 0;   // a lane quarter without a valid row (ragged last tile): nothing to convert
#pragma unroll 1
            for (int c = 0; c < nc; ++c) {
                const int col0 = half * (kHid / 2) + c * 32;
                tmemLd32(taddrLane + (uint32_t)(buf * kHid + col0), r);
                tmemLdWait();
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int h0 = col0 + 8 * q;
                    float bias[8];
                    loadF8(base + kOffB1 + 4u * h0, bias);
                    uint4 o;
                    __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        oh[i] = __floats2half2_rn(geluErf(__uint_as_float(r[8 * q + 2 * i]) + bias[2 * i]), geluErf(__uint_as_float(r[8 * q + 2 * i + 1]) + bias[2 * i + 1]));
                    stsV4(hRow + (uint32_t)(h0 >> 6) * kHChunk + (((uint32_t)((h0 & 63) >> 3) ^ sw) << 4), o);
                }
            }
            fenceProxyAsync();
            tcFenceBefore();
            __syncwarp();
            if (lane == 0) {
                mbarArrive(base + kBarHFull + 8u * buf);
                mbarArrive(base + kBarD1Empty + 8u * buf);
            }
        };
        uint4 res[6];
        auto prefetchResidual = [&](int t) {   // issued one GELU phase ahead of its use: the L2 latency is hidden behind phase 1
            const long long g = tok.start + (long long)t * kRows + row;
#pragma unroll
            for (int j = 0; j < 6; ++j) res[j] = make_uint4(0, 0, 0, 0);
            if (g < tok.end) {
                const uint4* xr = reinterpret_cast<const uint4*>(a.x + g * kC + half * (kC / 2));
#pragma unroll
                for (int j = 0; j < 6; ++j) res[j] = xr[j];
            }
        };
        auto phase2 = [&](int t) {   // x += fc2 + b2
            const long long g = tok.start + (long long)t * kRows + row;
            const bool valid = g < tok.end;
            const int col0 = half * (kC / 2);
            __half* xrow = a.x + g * kC + col0;
            mbarWait(base + kBarD2Full, (uint32_t)t & 1u);
            tcFenceAfter();
            uint32_t r2[32];
            tmemLd32(taddrLane + kTmemD2 + (uint32_t)col0, r);
            tmemLd16(taddrLane + kTmemD2 + (uint32_t)col0 + 32u, r2);
            tmemLdWait();
            tcFenceBefore();
            __syncwarp();
            if (lane == 0) mbarArrive(base + kBarD2Empty);
            if (valid) {
#pragma unroll
                for (int j = 0; j < 6; ++j) {
                    float bias[8], rv[8];
                    loadF8(base + kOffB2 + 4u * (col0 + 8 * j), bias);
                    unpack8(res[j], rv);
                    uint4 o;
                    __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float a0 = __uint_as_float(j < 4 ? r[8 * j + 2 * i] : r2[8 * (j - 4) + 2 * i]);
                        const float a1 = __uint_as_float(j < 4 ? r[8 * j + 2 * i + 1] : r2[8 * (j - 4) + 2 * i + 1]);
                        oh[i] = __floats2half2_rn(a0 + bias[2 * i] + rv[2 * i], a1 + bias[2 * i + 1] + rv[2 * i + 1]);
                    }
                    reinterpret_cast<uint4*>(xrow)[j] = o;
                }
            }
        };
        if (nMine > 0) phase1(0);
        for (int k = 0; k < nMine; ++k) {
            prefetchResidual(k);
            if (k + 1 < nMine) phase1(k + 1);
            phase2(k);
        }
    }

    tcFenceBefore();
    __syncthreads();
    if (warp == kMmaWarp) {
        tcFenceAfter();
        tmemDealloc(tmemBase, kTmemCols);
    }
}


// ------------------------------------------------------------------------------------------------------------------------------
// Streaming variant for wider levels (C = 192: the two weight matrices are 288 KB and cannot stay resident).  Same roles, but the
// hidden dimension is processed in chunks of 64 units: a TMA warp streams, per chunk, the 64 fc1 rows W1[c*64 .. +64][C] and the 64
// fc2 columns W2[C][c*64 .. +64] from L2 through a two-stage ring; fc1(chunk) -> D1 (64 TMEM columns, double-buffered) -> GELU ->
// hidden chunk in shared memory (double-buffered) -> fc2 accumulates the chunk into D2[128][C].  fc1 of chunk g+1 is queued before
// fc2 of chunk g.  With C = 192 a row is normalised by two producer threads (96 channels each, statistics exchanged by shuffle).
// ------------------------------------------------------------------------------------------------------------------------------
template <int C>
struct StreamCfg {
    static constexpr int kHidden = 2 * C;
    static constexpr int kChunk = 64;                         // hidden units per chunk
    static constexpr int kChunks = kHidden / kChunk;
    static constexpr int kKA = C / 32;                        // 32-channel K chunks of the A operand (SWIZZLE_64B)
    static constexpr int kPieces = C / 8;                     // 16-byte pieces of a token row
    static constexpr int kProd = 4;                           // producer warps: one token row per thread
    static constexpr int kTmaW = kProd, kMmaW = kProd + 1, kEpi0 = kProd + 2;
    static constexpr int kThreadsS = 32 * (kEpi0 + kEpiWarps);
    static constexpr int kABufs = C == 96 ? 2 : 1;
    static constexpr uint32_t kW1Bytes = kKA * kChunk * 64;   // kKA boxes [64 rows][32 k]
    static constexpr uint32_t kW2Bytes = C * 128;             // one box [C rows][64 k], SWIZZLE_128B
    static constexpr uint32_t kStage = kW1Bytes + kW2Bytes;
    static constexpr uint32_t kABytes = kKA * kAChunk;
    static constexpr uint32_t kOffRing = 4096;
    static constexpr uint32_t kOffAS = kOffRing + 2 * kStage;
    static constexpr uint32_t kOffHS = kOffAS + kABufs * kABytes;
    static constexpr uint32_t kSmem = kOffHS + 2 * kHChunk + 1024;
    static constexpr uint32_t kD2Col = 2 * kChunk;
    static constexpr uint32_t kTmem = C == 96 ? 256 : 512;
    // header: barriers, then b1[2C], b2[C], gamma[C], beta[C] as f32
    static constexpr uint32_t kB1 = 256, kB2 = kB1 + 4 * kHidden, kGamma = kB2 + 4 * C, kBeta = kGamma + 4 * C;
    static_assert(kBeta + 4 * C <= kOffRing, "constants overflow the header");
    static_assert(kStage % 1024 == 0 && kW1Bytes % 1024 == 0 && kABytes % 1024 == 0, "swizzled operands need 1024-byte alignment");
    static_assert(kSmem <= 227 * 1024, "shared memory budget");
};
// barriers of the streaming kernel (byte offsets)
constexpr uint32_t sBarWFull = 0, sBarWEmpty = 16, sBarAFull = 32, sBarAEmpty = 48, sBarD1Full = 64, sBarD1Empty = 80, sBarHFull = 96, sBarHEmpty = 112,
                   sBarD2Full = 128, sBarD2Empty = 136, sTmemSlot = 144, sBarW2Full = 152, sBarW2Empty = 168;   // sBarW*: the fc1 half of a ring stage

template <int C>
__global__ void __launch_bounds__(StreamCfg<C>::kThreadsS, 1) swin_mlp_stream_kernel(const __grid_constant__ MlpArgs a) {
    using Cfg = StreamCfg<C>;
    extern __shared__ uint8_t smemRaw[];
    const uint32_t base = (smemU32(smemRaw) + 1023u) & ~1023u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    pdlLaunchDependents();
    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) {
            mbarInit(base + sBarWFull + 8u * i, 1);
            mbarInit(base + sBarWEmpty + 8u * i, 1);
            mbarInit(base + sBarW2Full + 8u * i, 1);
            mbarInit(base + sBarW2Empty + 8u * i, 1);
            mbarInit(base + sBarAFull + 8u * i, 4);
            mbarInit(base + sBarAEmpty + 8u * i, 1);
            mbarInit(base + sBarD1Full + 8u * i, 1);
            mbarInit(base + sBarD1Empty + 8u * i, kEpiWarps);
            mbarInit(base + sBarHFull + 8u * i, kEpiWarps);
            mbarInit(base + sBarHEmpty + 8u * i, 1);
        }
        mbarInit(base + sBarD2Full, 1);
        mbarInit(base + sBarD2Empty, kEpiWarps);
        mbarInitFence();
        tmaPrefetchDesc(&a.tmW1);
        tmaPrefetchDesc(&a.tmW2);
    }
    for (int i = threadIdx.x; i < Cfg::kHidden; i += Cfg::kThreadsS) stsF32(base + Cfg::kB1 + 4u * i, a.b1[i]);
    for (int i = threadIdx.x; i < C; i += Cfg::kThreadsS) {
        stsF32(base + Cfg::kB2 + 4u * i, a.b2[i]);
        stsF32(base + Cfg::kGamma + 4u * i, a.gamma[i]);
        stsF32(base + Cfg::kBeta + 4u * i, a.beta[i]);
    }
    if (warp == Cfg::kMmaW) tmemAlloc(base + sTmemSlot, Cfg::kTmem);
    tcFenceBefore();
    __syncthreads();
    tcFenceAfter();
    uint32_t tmemBase;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmemBase) : "r"(base + sTmemSlot));
    const LinearTokens tok = linearTokens(a.tokens);
    const int nMine = tok.tiles();
    const int nChunksMine = nMine * Cfg::kChunks;

    if (warp == Cfg::kTmaW) {
        // ---- weight stream (constants: starts before the predecessor kernel has finished) ----
        if (lane == 0) {
            for (int g = 0; g < nChunksMine; ++g) {
                // the two halves of a stage are released separately: the fc1 rows of chunk g can be fetched as soon as fc1 of chunk
                // g-2 has completed, one GELU period before the fc2 columns of chunk g-2 are done with
                const int st = g & 1, c = g % Cfg::kChunks;
                const uint32_t par = (uint32_t)((g >> 1) & 1) ^ 1u, dst = base + Cfg::kOffRing + (uint32_t)st * Cfg::kStage;
                mbarWait(base + sBarWEmpty + 8u * st, par);
                const uint32_t full1 = base + sBarWFull + 8u * st, full2 = base + sBarW2Full + 8u * st;
                mbarExpectTx(full1, Cfg::kW1Bytes);
                for (int ka = 0; ka < Cfg::kKA; ++ka) tmaLoad2d(dst + (uint32_t)ka * (Cfg::kChunk * 64u), &a.tmW1, full1, ka * 32, c * Cfg::kChunk);
                mbarWait(base + sBarW2Empty + 8u * st, par);
                mbarExpectTx(full2, Cfg::kW2Bytes);
                tmaLoad2d(dst + Cfg::kW1Bytes, &a.tmW2, full2, c * Cfg::kChunk, 0);
            }
        }
    } else if (warp < Cfg::kProd) {
        // ---- LayerNorm producers: one token row per thread ----
        lnProducerLoop<C, Cfg::kABufs>(a.x, a.eps, tok, base, Cfg::kOffAS, Cfg::kGamma, Cfg::kBeta, base + sBarAFull, base + sBarAEmpty, nMine);
    } else if (warp == Cfg::kMmaW) {
        const uint32_t hi64 = descHi(512u, 4u), hi128 = descHi(1024u, 2u);
        const uint32_t idesc1 = instrDescF16(kRows, Cfg::kChunk), idesc2 = instrDescF16(kRows, C);
        auto fc1 = [&](int g) {   // D1[g & 1] = A(tile) x W1 chunk^T
            const int k = g / Cfg::kChunks, c = g - k * Cfg::kChunks, st = g & 1;
            const int buf = Cfg::kABufs == 2 ? (k & 1) : 0;
            const uint32_t use = Cfg::kABufs == 2 ? (uint32_t)(k >> 1) : (uint32_t)k;
            if (c == 0) mbarWait(base + sBarAFull + 8u * buf, use & 1u);
            mbarWait(base + sBarWFull + 8u * st, (uint32_t)(g >> 1) & 1u);
            mbarWait(base + sBarD1Empty + 8u * st, ((uint32_t)(g >> 1) & 1u) ^ 1u);
            tcFenceAfter();
            if (electOne()) {
                const uint32_t aBase = base + Cfg::kOffAS + (uint32_t)buf * Cfg::kABytes, wBase = base + Cfg::kOffRing + (uint32_t)st * Cfg::kStage;
#pragma unroll
                for (int ka = 0; ka < Cfg::kKA; ++ka)
#pragma unroll
                    for (int ks = 0; ks < 2; ++ks)
                        ummaLoHi(tmemBase + (uint32_t)st * Cfg::kChunk, descLo(aBase + ka * kAChunk + ks * 32u), hi64, descLo(wBase + ka * (Cfg::kChunk * 64u) + ks * 32u), hi64,
                                 idesc1, (ka | ks) != 0 ? 1u : 0u);
                tcCommit(base + sBarD1Full + 8u * st);
                tcCommit(base + sBarWEmpty + 8u * st);
                if (c == Cfg::kChunks - 1) tcCommit(base + sBarAEmpty + 8u * buf);   // the tile's normalised rows are no longer needed
            }
            __syncwarp();
        };
        auto fc2 = [&](int g) {   // D2 += hidden chunk x W2 chunk^T
            const int k = g / Cfg::kChunks, c = g - k * Cfg::kChunks, st = g & 1;
            mbarWait(base + sBarHFull + 8u * st, (uint32_t)(g >> 1) & 1u);
            mbarWait(base + sBarW2Full + 8u * st, (uint32_t)(g >> 1) & 1u);
            if (c == 0) mbarWait(base + sBarD2Empty, ((uint32_t)k & 1u) ^ 1u);
            tcFenceAfter();
            if (electOne()) {
                const uint32_t hBase = base + Cfg::kOffHS + (uint32_t)st * kHChunk, wBase = base + Cfg::kOffRing + (uint32_t)st * Cfg::kStage + Cfg::kW1Bytes;
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                    ummaLoHi(tmemBase + Cfg::kD2Col, descLo(hBase + ks * 32u), hi128, descLo(wBase + ks * 32u), hi128, idesc2, (c | ks) != 0 ? 1u : 0u);
                tcCommit(base + sBarHEmpty + 8u * st);
                tcCommit(base + sBarW2Empty + 8u * st);
                if (c == Cfg::kChunks - 1) tcCommit(base + sBarD2Full);
            }
            __syncwarp();
        };
        if (nChunksMine > 0) fc1(0);
        for (int g = 0; g < nChunksMine; ++g) {
            if (g + 1 < nChunksMine) fc1(g + 1);
            fc2(g);
        }
    } else {
        // ---- epilogue warps ----
        pdlWait();  // the residual rows come from the preceding kernel
        const int quarter = warp & 3;
        const int half = (warp - Cfg::kEpi0) >> 2;
        const int row = quarter * 32 + lane;
        const uint32_t taddrLane = tmemBase + ((uint32_t)(quarter * 32) << 16);
        const uint32_t sw = (uint32_t)row & 7u;
        uint32_t r[32];
        auto phase1 = [&](int g) {   // GELU(fc1 chunk + b1) -> hidden chunk buffer g & 1
            const int c = g % Cfg::kChunks, st = g & 1;
            const uint32_t ph = (uint32_t)(g >> 1) & 1u;
            const bool work = quarter * 32 < tok.validRows(g / Cfg::kChunks);   // a lane quarter without a valid row: nothing to convert
            mbarWait(base + sBarD1Full + 8u * st, ph);
            tcFenceAfter();
            if (work) {
                tmemLd32(taddrLane + (uint32_t)(st * Cfg::kChunk + half * 32), r);
                tmemLdWait();
            }
            tcFenceBefore();
            __syncwarp();
            if (lane == 0) mbarArrive(base + sBarD1Empty + 8u * st);   // the accumulator is in registers: fc1 of chunk g+2 may overwrite it
            mbarWait(base + sBarHEmpty + 8u * st, ph ^ 1u);   // fc2 of chunk g-2 has consumed this buffer
            const uint32_t hRow = base + Cfg::kOffHS + (uint32_t)st * kHChunk + (uint32_t)row * 128u;
            if (work) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    float bias[8];
                    loadF8(base + Cfg::kB1 + 4u * (uint32_t)(c * Cfg::kChunk + half * 32 + 8 * q), bias);
                    uint4 o;
                    __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        oh[i] = __floats2half2_rn(geluErf(__uint_as_float(r[8 * q + 2 * i]) + bias[2 * i]), geluErf(__uint_as_float(r[8 * q + 2 * i + 1]) + bias[2 * i + 1]));
                    stsV4(hRow + (((uint32_t)(half * 4 + q) ^ sw) << 4), o);
                }
            }
            fenceProxyAsync();
            __syncwarp();
            if (lane == 0) mbarArrive(base + sBarHFull + 8u * st);
        };
        auto phase2 = [&](int k) {   // x += fc2 + b2 for tile k
            const long long g = tok.start + (long long)k * kRows + row;
            const bool valid = g < tok.end;
            const int col0 = half * (C / 2);
            __half* xrow = a.x + g * C + col0;
            mbarWait(base + sBarD2Full, (uint32_t)k & 1u);
            tcFenceAfter();
#pragma unroll 1
            for (int p = 0; p < C / 64; ++p) {   // 32 columns per round
                uint4 res[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) res[j] = make_uint4(0, 0, 0, 0);
                if (valid) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) res[j] = reinterpret_cast<const uint4*>(xrow)[4 * p + j];
                }
                tmemLd32(taddrLane + Cfg::kD2Col + (uint32_t)(col0 + 32 * p), r);
                tmemLdWait();
                if (valid) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float bias[8], rv[8];
                        loadF8(base + Cfg::kB2 + 4u * (uint32_t)(col0 + 32 * p + 8 * j), bias);
                        unpack8(res[j], rv);
                        uint4 o;
                        __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            oh[i] = __floats2half2_rn(__uint_as_float(r[8 * j + 2 * i]) + bias[2 * i] + rv[2 * i], __uint_as_float(r[8 * j + 2 * i + 1]) + bias[2 * i + 1] + rv[2 * i + 1]);
                        reinterpret_cast<uint4*>(xrow)[4 * p + j] = o;
                    }
                }
            }
            if (C % 64) {   // 16-column tail (C = 96: 48 columns per thread)
                constexpr int p16 = (C / 64) * 32;
                uint4 res[2] = {make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0)};
                if (valid) {
                    res[0] = reinterpret_cast<const uint4*>(xrow)[p16 / 8];
                    res[1] = reinterpret_cast<const uint4*>(xrow)[p16 / 8 + 1];
                }
                tmemLd16(taddrLane + Cfg::kD2Col + (uint32_t)(col0 + p16), r);
                tmemLdWait();
                if (valid) {
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        float bias[8], rv[8];
                        loadF8(base + Cfg::kB2 + 4u * (uint32_t)(col0 + p16 + 8 * j), bias);
                        unpack8(res[j], rv);
                        uint4 o;
                        __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            oh[i] = __floats2half2_rn(__uint_as_float(r[8 * j + 2 * i]) + bias[2 * i] + rv[2 * i], __uint_as_float(r[8 * j + 2 * i + 1]) + bias[2 * i + 1] + rv[2 * i + 1]);
                        reinterpret_cast<uint4*>(xrow)[p16 / 8 + j] = o;
                    }
                }
            }
            tcFenceBefore();
            __syncwarp();
            if (lane == 0) mbarArrive(base + sBarD2Empty);
        };
        // chunk g+1 is converted before the finished tile of chunk g is written back (its fc2 runs meanwhile)
        if (nChunksMine > 0) phase1(0);
        for (int g = 0; g < nChunksMine; ++g) {
            if (g + 1 < nChunksMine) phase1(g + 1);
            if (g % Cfg::kChunks == Cfg::kChunks - 1) phase2(g / Cfg::kChunks);
        }
    }

    tcFenceBefore();
    __syncthreads();
    if (warp == Cfg::kMmaW) {
        tcFenceAfter();
        tmemDealloc(tmemBase, Cfg::kTmem);
    }
}


}  // namespace

struct SwinMlpPlan {
    MlpArgs args;
    int c = 0;
    bool stream = false;   // weights streamed per hidden chunk (swin_mlp_stream_kernel) instead of resident
};

bool swinMlpSupported(int c, int hidden) { return (c == 96 || c == 192) && hidden == 2 * c; }

// variant: 0 = resident weights where they fit (C = 96), streamed otherwise; 1 = always streamed
SwinMlpPlan* swinMlpCreatePlan(__half* x, int c, const float* gamma, const float* beta, float eps, const __half* w1, const float* b1, const __half* w2, const float* b2,
                               int variant) {
    if (!swinMlpSupported(c, 2 * c)) throw Error("swin mlp: unsupported width");
    if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(w1) | reinterpret_cast<uintptr_t>(w2)) & 15) throw Error("swin mlp: operands must be 16-byte aligned");
    SwinMlpPlan* plan = new SwinMlpPlan{};
    plan->c = c;
    plan->stream = variant == 1 || c != kC;
    try {
        // fc1 weights [2C][C] and fc2 weights [C][2C], both K-major
        if (plan->stream) {
            encodeMatrixMap2d(&plan->args.tmW1, w1, c, 2 * c, 32, 64, false);
            encodeMatrixMap2d(&plan->args.tmW2, w2, 2 * c, c, 64, c, true);
        } else {
            encodeMatrixMap2d(&plan->args.tmW1, w1, kC, kHid, 32, kHid, false);
            encodeMatrixMap2d(&plan->args.tmW2, w2, kHid, kC, 64, kC, true);
        }
    } catch (...) {
        delete plan;
        throw;
    }
    plan->args.x = x;
    plan->args.b1 = b1;
    plan->args.b2 = b2;
    plan->args.gamma = gamma;
    plan->args.beta = beta;
    plan->args.eps = eps;
    plan->args.tokens = 0;
    return plan;
}

void swinMlpDestroyPlan(SwinMlpPlan* plan) { delete plan; }

const char* swinMlpDescribe(const SwinMlpPlan* plan, char* buf, int cap) {
    const unsigned smem = !plan->stream ? kMlpSmem : plan->c == 96 ? StreamCfg<96>::kSmem : StreamCfg<192>::kSmem;
    std::snprintf(buf, cap, "swin-mlp fused LN+fc1+GELU+fc2+residual (tcgen05) c=%d hidden=%d rows=%d weights=%s smem=%u", plan->c, 2 * plan->c, kRows,
                  plan->stream ? "streamed" : "resident", smem);
    return buf;
}

void swinMlpLaunch(const SwinMlpPlan* plan, cudaStream_t s, long long tokens) {
    static bool attrSet[64] = {};
    static int sms[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    if (!attrSet[dev]) {
        cudaFuncSetAttribute(swin_mlp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMlpSmem);
        cudaFuncSetAttribute(swin_mlp_stream_kernel<96>, cudaFuncAttributeMaxDynamicSharedMemorySize, StreamCfg<96>::kSmem);
        cudaFuncSetAttribute(swin_mlp_stream_kernel<192>, cudaFuncAttributeMaxDynamicSharedMemorySize, StreamCfg<192>::kSmem);
        cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev);
        if (sms[dev] <= 0) sms[dev] = 148;
        attrSet[dev] = true;
    }
    if (tokens <= 0) return;
    MlpArgs a = plan->args;
    a.tokens = tokens;
    const long long tiles = (tokens + kRows - 1) / kRows;
    const dim3 grid((unsigned)(tiles < sms[dev] ? tiles : sms[dev]));
    cudaError_t e;
    if (!plan->stream) e = launchPdl(swin_mlp_kernel, grid, dim3(kMlpThreads), kMlpSmem, s, a);
    else if (plan->c == 96) e = launchPdl(swin_mlp_stream_kernel<96>, grid, dim3(StreamCfg<96>::kThreadsS), StreamCfg<96>::kSmem, s, a);
    else e = launchPdl(swin_mlp_stream_kernel<192>, grid, dim3(StreamCfg<192>::kThreadsS), StreamCfg<192>::kSmem, s, a);
    if (e != cudaSuccess) throw Error(std::string("swin mlp launch: ") + cudaGetErrorString(e));
}

}  // namespace w2x
