// Fused attention half of a SwinUNet block (torchvision SwinTransformerBlock: x = x + proj(W-MSA / SW-MSA(LayerNorm(x)))), one
// kernel instead of layernorm_kernel + QKV igemm + window_attention_kernel (mma.sync) + proj igemm.  The token rows are read from HBM
// once and written once; the normalised rows, Q, K, V, the scores, the probabilities and the attention output never leave the SM.
//
// A tile is three 6x6 windows (108 tokens, rows 108..127 of the M = 128 UMMA tile are padding).  Row r holds position r % 36 of window
// r / 36; the cyclic shift (torch.roll) and the window partition are the address math of the gather / scatter.  Per tile and per
// 32-channel chunk c of Q / K / V (C = 96: a pair of 16-wide heads):
//
//   QKV(c)   D1[128][96]  = LN(x)[128][C] . W'[c]^T          6 x tcgen05.mma M128 N96 K16   (W' = rows of Wqkv regrouped per chunk,
//                                                                                           q rows pre-multiplied by d^-1/2 log2 e)
//   E1(c)    D1 + bias -> fp16 -> Q[128][32], K[key][32] (K-major SWIZZLE_64B operands), V^T[32][key] (SWIZZLE_128B); the key index of
//            row r is 40 (r / 36) + r % 36, so a window's keys start at a 16-byte chunk of a probability row
//   S(h)     S_h[128][128] = Q_h K_h^T                       1 x M128 N128 K16 per head (block-diagonal: only the 36 columns of a row's
//                                                            own window are ever read back)
//   SM(h)    thread = query row: tcgen05.ld of its window's 40 columns, + relative-position bias (+ shift mask), exp2 softmax,
//            probabilities -> fp16 -> P[128][128] (SWIZZLE_128B A operand; off-diagonal blocks stay zero from the prologue)
//   PV(h)    O[128][16 h .. +16] = P V_h                     8 x M128 N16 K16
//   E2(c)    O -> fp16 -> Oc[128][32] (A operand)
//   proj(c)  D2[128][C] += Oc . Wproj[:, 32 c .. +32]^T      2 x M128 N96 K16
//   E3       D2 + bias + residual -> x (in place, scattered back to the tokens' positions)
//
// Roles (16 warps, four per scheduler; a warp reads the TMEM lane quarter given by its index modulo 4):
//   warps 0-3    LayerNorm producers (swin_token.cuh, one token row per thread) + E3 of the tile two behind (their index is the lane quarter)
//   warps 4-7    row warps: E1 and E2 for the 32 rows of their quarter
//   warp  12     issuer A: QKV and proj MMAs          warp 15   issuer B: S and PV MMAs (one elected lane each)
//   warps 8-11, 13, 14   softmax: quarters 0 / 3 hold rows of one window (one warp), quarters 1 / 2 straddle two windows (one warp per
//                window); every softmax warp runs one pass per head
// Each hand-off (tcgen05.commit -> mbarrier -> waiting warp) then sits between DIFFERENT warps, so the row warps convert chunk g + 1 and
// the issuers queue its MMAs while the softmax warps are still on chunk g.  (The first version ran SM, E1, E2, E3 in sequence on eight
// epilogue warps: 0.18 ms per level-1 block, every warp latency-bound on its own chain of waits; profiles/r02_ncu_swin_attn_v1.txt.)
// At C = 96 all weights (Wqkv' 54 KB, Wproj 18 KB) and the bias tables (6 heads x [36][44] fp32) stay resident; the C = 192 instantiation
// (one 32-wide head per chunk, Wqkv' streamed, no projection) is described at ACfg below.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <string>
#include <vector>

#include "conv_params.h"
#include "launch.h"
#include "sm100_common.cuh"
#include "swin_token.cuh"

namespace w2x {
using namespace sm100;
using namespace swintok;
namespace {

constexpr int kHeads = 6;
constexpr int kWin = 6, kNT = 36;      // window side, tokens per window
constexpr int kWinTile = 3;            // windows per tile
constexpr int kKeyStride = 40;         // key index of window w, position p: 40 w + p (a window starts at a 16-byte chunk of a P row)
constexpr int kBiasPitch = 44;         // floats per bias-table row: 16-byte loads of 8 consecutive rows hit distinct banks
constexpr uint32_t kBiasHead = kNT * kBiasPitch * 4;
constexpr uint32_t kBiasBytes = kHeads * kBiasHead;
constexpr int kRowWarp0 = 4, kRowWarps = 4, kSmWarps = 6, kIssuerA = 12, kIssuerB = 15;
constexpr int kAttnThreads = 32 * 16;
constexpr uint32_t kWqkvChunk = 96 * 64;               // the 96 rows (q32 | k32 | v32) of one channel chunk, one 32-wide K chunk: [96][32 k] fp16
constexpr uint32_t kVBuf = 2 * 32 * 128;               // V^T: two 64-key chunks of [32 dims][64 keys] fp16, SWIZZLE_128B
constexpr uint32_t kPChunk = kRows * 128;              // [128 rows][64 keys]

// barriers (byte offsets from the 1024-aligned base)
constexpr uint32_t bW = 0, bAFull = 8, bAEmpty = 16, bD1Full = 24, bQKFull = 32, bQKEmpty = 40, bVEmpty = 48, bSFull = 64, bSEmpty = 80, bPFull = 96, bPEmpty = 104,
                   bOFull = 112, bOEmpty = 128, bOcFull = 144, bOcEmpty = 152, bD2Full = 160, bD2Empty = 168, kTmemSlot = 176, bWFull = 184, bWEmpty = 200, bD1Full2 = 216, bD1Empty = 224;

// Two instantiations:
//   C = 96  (levels 1 / 5, head dim 16: a 32-channel chunk is a PAIR of heads): everything above, all weights resident;
//   C = 192 (levels 2 - 4, head dim 32: a chunk is one head): the 221 KB of Wqkv' stream through a two-stage TMA ring (one 36 KB chunk per
//           stage, issued by issuer A one chunk ahead), and the kernel ends at the attention output (fp16, written to `out` by the row
//           warps): Wproj and a 192-column accumulator do not fit next to the ring, the output projection stays on the Linear kernel.
template <int C>
struct ACfg {
    static constexpr int kC = C;
    static constexpr int kHD = C / kHeads;            // 16 | 32
    static constexpr int kHPC = 32 / kHD;             // heads per 32-channel chunk
    static constexpr int kChunks = C / 32;            // channel chunks of Q / K / V = K chunks of the LayerNorm rows
    static constexpr bool kProj = C == 96;            // output projection + residual inside the kernel
    static constexpr bool kResident = C == 96;        // Wqkv' resident (else streamed per chunk)
    static constexpr uint32_t kOffBqkv = 256, kOffBproj = kOffBqkv + 3 * C * 4, kOffGamma = kOffBproj + (kProj ? C * 4 : 0), kOffBeta = kOffGamma + C * 4;
    static constexpr uint32_t kWqkvKa = 3 * C * 64;           // resident: one 32-wide K chunk of W': [3C rows][32 k] fp16, SWIZZLE_64B
    static constexpr uint32_t kStage = kChunks * kWqkvChunk;  // streamed: the chunk's 96 rows, all K chunks
    static constexpr uint32_t kWprojKa = C * 64;              // [C rows][32 k]
    static constexpr uint32_t kOffWqkv = 4096;
    static constexpr uint32_t kOffWproj = kOffWqkv + (kResident ? kChunks * kWqkvKa : 2 * kStage);
    static constexpr uint32_t kOffBias = kOffWproj + (kProj ? kChunks * kWprojKa : 0);
    static constexpr uint32_t kOffA = (kOffBias + kBiasBytes + 1023u) & ~1023u;
    static constexpr uint32_t kOffQ = kOffA + kChunks * kAChunk;
    static constexpr uint32_t kOffK = kOffQ + kAChunk;
    static constexpr uint32_t kOffV = kOffK + kAChunk;
    static constexpr uint32_t kOffP = kOffV + 2 * kVBuf;
    static constexpr uint32_t kOffOc = kOffP + 2 * kPChunk;
    static constexpr uint32_t kZeroEnd = kOffOc + (kProj ? kAChunk : 0);
    static constexpr uint32_t kSmem = kZeroEnd + 1024;   // + alignment slack
    // TMEM columns
    static constexpr int kD1Bufs = kProj ? 1 : 2;     // QKV accumulators: the second one takes the columns of the proj accumulator
    static constexpr uint32_t tD1 = 0, tD2 = 96, tS = 192, tO = tS + 256, kTmemCols = 512;
    static_assert(kOffBeta + C * 4 <= kOffWqkv, "constants overflow the header");
    static_assert(kOffWproj % 1024 == 0 && kOffA % 1024 == 0 && kOffV % 1024 == 0 && kOffP % 1024 == 0 && kOffOc % 1024 == 0, "swizzled operands need 1024-byte alignment");
    static_assert(kBiasBytes % 16 == 0 && kOffBias % 16 == 0, "bulk copy granularity");
    static_assert(kSmem <= 227 * 1024, "shared memory budget");
    static_assert(tO + 64 <= kTmemCols, "tensor memory budget");
};

struct AttnArgs {
    CUtensorMap tmWqkv, tmWproj;
    __half* x;               // [n][h][w][C] fp16, updated in place
    const float* bqkv;       // [3C] regrouped like W' (q part pre-scaled)
    const float* bproj;      // [C]
    const float* gamma;      // [C]
    const float* beta;       // [C]
    const float* relpos;     // [heads][36][kBiasPitch] fp32, multiplied by log2 e
    __half* out;             // C = 192 only: attention output [n][h][w][C] (before the output projection)
    float eps;
    int h, w, shiftY, shiftX;   // cyclic shift per dimension (torchvision drops it in a dimension the window covers)
    int nwx, nwy;            // windows per row / column of an image
    long long windows;       // n * nwy * nwx
};

__device__ __forceinline__ void bulkLoad1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void tmemLd8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
}
__device__ __forceinline__ void stsU16(uint32_t addr, unsigned short v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"(v) : "memory"); }
__device__ __forceinline__ float ex2f(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ uint32_t packH2(float a, float b) {
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
}

// token held by row `row` of tile `tile` (window-ordered gather; < 0 for padding rows and windows past the end)
struct WindowTokens {
    int h, w, shiftY, shiftX, nwx, nwy;
    long long windows;
    int first, step;
    __device__ __forceinline__ long long tokenOfTile(long long tile, int row) const {
        if (row >= kWinTile * kNT) return -1;
        const int wi = row / kNT, p = row - wi * kNT;
        long long win = tile * kWinTile + wi;
        if (win >= windows) return -1;
        const int wx = (int)(win % nwx);
        win /= nwx;
        const int wy = (int)(win % nwy);
        const long long img = win / nwy;
        int y = wy * kWin + p / kWin + shiftY, x = wx * kWin + p % kWin + shiftX;   // torch.roll(x, -shift): rolled[p] = x[p + shift]
        if (y >= h) y -= h;
        if (x >= w) x -= w;
        return (img * h + y) * (long long)w + x;
    }
    __device__ __forceinline__ long long operator()(int k, int row) const { return tokenOfTile((long long)first + (long long)k * step, row); }
};

// 16 warps x 128 registers: the whole register file
template <int C>
__global__ void __launch_bounds__(kAttnThreads, 1) swin_attn_kernel(const __grid_constant__ AttnArgs a) {
    using F = ACfg<C>;
    constexpr int kC = F::kC, kHD = F::kHD, kHPC = F::kHPC, kChunks = F::kChunks;
    constexpr uint32_t kOffBqkv = F::kOffBqkv, kOffBproj = F::kOffBproj, kOffGamma = F::kOffGamma, kOffBeta = F::kOffBeta, kWqkvKa = F::kWqkvKa, kWprojKa = F::kWprojKa,
                       kOffWqkv = F::kOffWqkv, kOffWproj = F::kOffWproj, kOffBias = F::kOffBias, kOffA = F::kOffA, kOffQ = F::kOffQ, kOffK = F::kOffK, kOffV = F::kOffV,
                       kOffP = F::kOffP, kOffOc = F::kOffOc, tD1 = F::tD1, tD2 = F::tD2, tS = F::tS, tO = F::tO, kTmemCols = F::kTmemCols, kStage = F::kStage;
    (void)kOffBproj; (void)kWqkvKa; (void)kWprojKa; (void)kOffWproj; (void)kOffOc; (void)tD2; (void)kStage; (void)kHPC;
    extern __shared__ uint8_t smemRaw[];
    const uint32_t base = (smemU32(smemRaw) + 1023u) & ~1023u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    pdlLaunchDependents();

    // ---- prologue: constants only (weights, biases, LayerNorm parameters, bias tables) ----
    if (threadIdx.x == 0) {
        mbarInit(base + bW, 1);
        mbarInit(base + bAFull, 4);
        mbarInit(base + bAEmpty, 1);
        mbarInit(base + bD1Full, 1);
        mbarInit(base + bD1Full2, 1);
        mbarInit(base + bQKFull, kRowWarps);
        mbarInit(base + bQKEmpty, 1);
        mbarInit(base + bPFull, kSmWarps);
        mbarInit(base + bPEmpty, 1);
        mbarInit(base + bOcFull, kRowWarps);
        mbarInit(base + bOcEmpty, 1);
        mbarInit(base + bD2Full, 1);
        mbarInit(base + bD2Empty, kRowWarps);
        for (int i = 0; i < 2; ++i) {
            mbarInit(base + bVEmpty + 8u * i, 1);
            mbarInit(base + bSFull + 8u * i, 1);
            mbarInit(base + bSEmpty + 8u * i, kSmWarps);
            mbarInit(base + bOFull + 8u * i, 1);
            mbarInit(base + bOEmpty + 8u * i, kRowWarps);
            mbarInit(base + bWFull + 8u * i, 1);
            mbarInit(base + bWEmpty + 8u * i, 1);
            mbarInit(base + bD1Empty + 8u * i, kRowWarps);
        }
        mbarInitFence();
        tmaPrefetchDesc(&a.tmWqkv);
        if constexpr (F::kProj) tmaPrefetchDesc(&a.tmWproj);
    }
    for (int i = threadIdx.x; i < 3 * kC; i += kAttnThreads) stsF32(base + kOffBqkv + 4u * i, a.bqkv[i]);
    for (int i = threadIdx.x; i < kC; i += kAttnThreads) {
        if constexpr (F::kProj) stsF32(base + kOffBproj + 4u * i, a.bproj[i]);
        stsF32(base + kOffGamma + 4u * i, a.gamma[i]);
        stsF32(base + kOffBeta + 4u * i, a.beta[i]);
    }
    // Q / K / V / P / Oc start as zeros: padding keys (K rows and V^T columns nobody writes) must be finite, and the off-diagonal
    // blocks of P (rows of one window x keys of another) are never written
    for (uint32_t o = (uint32_t)threadIdx.x * 16u; o < F::kZeroEnd - kOffQ; o += (uint32_t)kAttnThreads * 16u) stsV4(base + kOffQ + o, make_uint4(0, 0, 0, 0));
    fenceProxyAsync();
    if (warp == kIssuerA) tmemAlloc(base + kTmemSlot, kTmemCols);
    tcFenceBefore();
    __syncthreads();
    tcFenceAfter();
    uint32_t tmemBase;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmemBase) : "r"(base + kTmemSlot));
    if (warp == kIssuerA && lane == 0) {
        if constexpr (F::kResident) {
            mbarExpectTx(base + bW, kChunks * kWqkvKa + kChunks * kWprojKa + kBiasBytes);
            for (int ka = 0; ka < kChunks; ++ka) {
                for (int c = 0; c < kChunks; ++c) tmaLoad2d(base + kOffWqkv + ka * kWqkvKa + c * kWqkvChunk, &a.tmWqkv, base + bW, ka * 32, c * 96);
                tmaLoad2d(base + kOffWproj + ka * kWprojKa, &a.tmWproj, base + bW, ka * 32, 0);
            }
        } else {
            mbarExpectTx(base + bW, kBiasBytes);
        }
        bulkLoad1d(base + kOffBias, a.relpos, kBiasBytes, base + bW);
    }
    const long long tiles = (a.windows + kWinTile - 1) / kWinTile;
    const int first = blockIdx.x, step = gridDim.x;
    const int nMine = first < tiles ? (int)((tiles - first + step - 1) / step) : 0;
    const int G = nMine * kChunks;   // this CTA's chunk sequence: g = kChunks k + c; heads j = kHPC g + h
    const WindowTokens tokens{a.h, a.w, a.shiftY, a.shiftX, a.nwx, a.nwy, a.windows, first, step};

    const uint32_t hi64 = descHi(512u, 4u), hi128 = descHi(1024u, 2u);
    if (warp < kRowWarp0) {
        // ---- LayerNorm producers: thread = tile row, gathers its token through the window map; after handing tile k + 1 to the tensor
        // core they write back tile k (E3: their warp index is also the TMEM lane quarter of their rows) ----
        const int quarter = warp;
        const int row = quarter * 32 + lane;
        const uint32_t taddrLane = tmemBase + ((uint32_t)(quarter * 32) << 16);
        uint32_t r[32], r2[32];
        (void)row; (void)taddrLane; (void)r; (void)r2;
        [[maybe_unused]] auto e3 = [&](int k) {   // x += D2 + bproj, 48 columns at a time
            const long long tok = tokens(k, row);
            __half* xrow = a.x + (tok >= 0 ? tok : 0) * kC;
            uint4 res[6];
#pragma unroll
            for (int q = 0; q < 6; ++q) res[q] = make_uint4(0, 0, 0, 0);
            if (tok >= 0) {   // issued before the accumulator wait (an L2 hit: the producers read these rows one tile ago)
#pragma unroll
                for (int q = 0; q < 6; ++q) res[q] = reinterpret_cast<const uint4*>(xrow)[q];
            }
            mbarWait(base + bD2Full, (uint32_t)k & 1u);
            tcFenceAfter();
#pragma unroll 1
            for (int half = 0; half < 2; ++half) {
                const int col0 = half * (kC / 2);
                tmemLd32(taddrLane + tD2 + (uint32_t)col0, r);
                tmemLd16(taddrLane + tD2 + (uint32_t)col0 + 32u, r2);
                tmemLdWait();
                if (half == 1) {
                    tcFenceBefore();
                    __syncwarp();
                    if (lane == 0) mbarArrive(base + bD2Empty);
                }
                uint4 nxt[6];
                if (half == 0 && tok >= 0) {
#pragma unroll
                    for (int q = 0; q < 6; ++q) nxt[q] = reinterpret_cast<const uint4*>(xrow)[6 + q];
                }
                if (tok >= 0) {
#pragma unroll
                    for (int q = 0; q < 6; ++q) {
                        float bias[8], rv[8];
                        loadF8(base + kOffBproj + 4u * (uint32_t)(col0 + 8 * q), bias);
                        unpack8(res[q], rv);
                        uint4 o;
                        __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const float a0 = __uint_as_float(q < 4 ? r[8 * q + 2 * i] : r2[8 * (q - 4) + 2 * i]);
                            const float a1 = __uint_as_float(q < 4 ? r[8 * q + 2 * i + 1] : r2[8 * (q - 4) + 2 * i + 1]);
                            oh[i] = __floats2half2_rn(a0 + bias[2 * i] + rv[2 * i], a1 + bias[2 * i + 1] + rv[2 * i + 1]);
                        }
                        reinterpret_cast<uint4*>(xrow)[6 * half + q] = o;
                    }
                }
                if (half == 0) {
#pragma unroll
                    for (int q = 0; q < 6; ++q) res[q] = nxt[q];
                }
            }
        };
        // tile k's rows are handed over while tile k - 1 is being processed; tile k - 2's accumulator is complete (or about to be) by then,
        // so the write-back never holds up the next tile's loads
        if constexpr (F::kProj) {
            lnProducerLoop<kC, 1>(a.x, a.eps, tokens, base, kOffA, kOffGamma, kOffBeta, base + bAFull, base + bAEmpty, nMine, 0, 1, [&](int k) {
                if (k >= 2) e3(k - 2);
            });
            if (nMine >= 2) e3(nMine - 2);
            if (nMine >= 1) e3(nMine - 1);
        } else {
            lnProducerLoop<kC, 1>(a.x, a.eps, tokens, base, kOffA, kOffGamma, kOffBeta, base + bAFull, base + bAEmpty, nMine);
        }
    } else if (warp == kIssuerA) {
        // ---- issuer A: QKV(g) and proj(g - 3), in the order the row warps produce their inputs ----
        const uint32_t idQkv = instrDescF16(kRows, 96), idProj = instrDescF16(kRows, kC);
        [[maybe_unused]] auto proj = [&](int g) {   // D2 += Oc x Wproj[:, chunk]^T
            const int k = g / kChunks, c = g - k * kChunks;
            mbarWait(base + bOcFull, (uint32_t)g & 1u);
            if (c == 0) mbarWait(base + bD2Empty, ((uint32_t)k & 1u) ^ 1u);
            tcFenceAfter();
            if (electOne()) {
#pragma unroll
                for (int ks = 0; ks < 2; ++ks)
                    ummaLoHi(tmemBase + tD2, descLo(base + kOffOc + ks * 32u), hi64, descLo(base + kOffWproj + c * kWprojKa + ks * 32u), hi64, idProj, (c | ks) != 0 ? 1u : 0u);
                tcCommit(base + bOcEmpty);
                if (c == kChunks - 1) tcCommit(base + bD2Full);
            }
            __syncwarp();
        };
        // streamed weights: the chunk's rows [96][C] of W' (kChunks boxes of 32 k) into ring stage g & 1; the stage was last read by QKV(g - 2)
        auto loadChunk = [&](int g) {
            const int st = g & 1, c = g % kChunks;
            mbarWait(base + bWEmpty + 8u * st, ((uint32_t)(g >> 1) & 1u) ^ 1u);
            if (lane == 0) {
                const uint32_t full = base + bWFull + 8u * st, dst = base + kOffWqkv + (uint32_t)st * kStage;
                mbarExpectTx(full, kStage);
                for (int ka = 0; ka < kChunks; ++ka) tmaLoad2d(dst + (uint32_t)ka * kWqkvChunk, &a.tmWqkv, full, ka * 32, c * 96);
            }
            __syncwarp();
        };
        if constexpr (F::kResident) mbarWait(base + bW, 0);
        else if (G > 0) loadChunk(0);
        for (int g = 0; g < G; ++g) {
            const int k = g / kChunks, c = g - k * kChunks;
            const int d = g % F::kD1Bufs;   // accumulator buffer
            if constexpr (F::kD1Bufs == 1) {
                if (g >= 1) mbarWait(base + bQKFull, (uint32_t)(g - 1) & 1u);   // E1 of the previous chunk has drained D1
            } else {
                mbarWait(base + bD1Empty + 8u * d, ((uint32_t)(g >> 1) & 1u) ^ 1u);   // E1 of chunk g - 2 has drained this accumulator
            }
            if constexpr (!F::kResident) mbarWait(base + bWFull + 8u * (g & 1), (uint32_t)(g >> 1) & 1u);
            if (c == 0) mbarWait(base + bAFull, (uint32_t)k & 1u);
            tcFenceAfter();
            if (electOne()) {
                const uint32_t wBase = F::kResident ? base + kOffWqkv + c * kWqkvChunk : base + kOffWqkv + (uint32_t)(g & 1) * kStage;
                const uint32_t wKa = F::kResident ? kWqkvKa : kWqkvChunk;
#pragma unroll
                for (int ka = 0; ka < kChunks; ++ka)
#pragma unroll
                    for (int ks = 0; ks < 2; ++ks)
                        ummaLoHi(tmemBase + tD1 + 96u * d, descLo(base + kOffA + ka * kAChunk + ks * 32u), hi64, descLo(wBase + ka * wKa + ks * 32u), hi64, idQkv,
                                 (ka | ks) != 0 ? 1u : 0u);
                tcCommit(base + (d == 0 ? bD1Full : bD1Full2));
                if constexpr (!F::kResident) tcCommit(base + bWEmpty + 8u * (g & 1));
                if (c == kChunks - 1) tcCommit(base + bAEmpty);   // the tile's normalised rows are no longer needed
            }
            __syncwarp();
            if constexpr (!F::kResident) {
                if (g + 1 < G) loadChunk(g + 1);   // waits for QKV(g - 1), queued one iteration ago, to release the other ring stage
            }
            if constexpr (F::kProj) {
                if (g >= 3) proj(g - 3);   // its Oc tile was written an iteration ago: never blocks the next QKV
            }
        }
        if constexpr (F::kProj) {
            for (int g = G > 3 ? G - 3 : 0; g < G; ++g) proj(g);
        }
    } else if (warp == kIssuerB) {
        // ---- issuer B: S(2g), S(2g + 1) of the chunk the row warps just converted, interleaved with PV of the previous chunk ----
        const uint32_t idS = instrDescF16(kRows, 128), idPv = instrDescF16(kRows, kHD);
        auto scores = [&](int j) {   // S[j & 1] = Q_h K_h^T for head j (head j % kHPC of the chunk in the Q / K buffers)
            const int sb = j & 1, hh = j % kHPC;
            mbarWait(base + bSEmpty + 8u * sb, ((uint32_t)(j >> 1) & 1u) ^ 1u);
            tcFenceAfter();
            if (electOne()) {
#pragma unroll
                for (int ks = 0; ks < kHD / 16; ++ks)
                    ummaLoHi(tmemBase + tS + 128u * sb, descLo(base + kOffQ + 32u * (hh + ks)), hi64, descLo(base + kOffK + 32u * (hh + ks)), hi64, idS, ks != 0 ? 1u : 0u);
                tcCommit(base + bSFull + 8u * sb);
                if (hh == kHPC - 1) tcCommit(base + bQKEmpty);
            }
            __syncwarp();
        };
        auto pv = [&](int j) {   // O[g & 1][kHD hh .. +kHD] = P V_h
            const int g = j / kHPC, hh = j - g * kHPC, ob = g & 1;
            mbarWait(base + bPFull, (uint32_t)j & 1u);
            if (hh == 0) mbarWait(base + bOEmpty + 8u * ob, ((uint32_t)(g >> 1) & 1u) ^ 1u);
            tcFenceAfter();
            if (electOne()) {
#pragma unroll
                for (int kc = 0; kc < 2; ++kc)
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)
                        ummaLoHi(tmemBase + tO + 32u * ob + (uint32_t)(kHD * hh), descLo(base + kOffP + kc * kPChunk + ks * 32u), hi128,
                                 descLo(base + kOffV + ob * kVBuf + kc * 4096u + (uint32_t)hh * (kHD * 128u) + ks * 32u), hi128, idPv, (kc | ks) != 0 ? 1u : 0u);
                tcCommit(base + bPEmpty);
                if (hh == kHPC - 1) {
                    tcCommit(base + bOFull + 8u * ob);
                    tcCommit(base + bVEmpty + 8u * ob);
                }
            }
            __syncwarp();
        };
        // the scores of chunk g are queued BEFORE the PV MMAs of chunk g - 1: a softmax warp that finishes a head finds the next one ready
        // instead of waiting for PV + S + two hand-offs (each scores() waits only until the previous head of its parity has been LOADED)
        for (int g = 0; g <= G; ++g) {
            if (g < G) {
                mbarWait(base + bQKFull, (uint32_t)g & 1u);
#pragma unroll
                for (int hh = 0; hh < kHPC; ++hh) scores(g * kHPC + hh);
            }
            if (g >= 1) {
#pragma unroll
                for (int hh = 0; hh < kHPC; ++hh) pv((g - 1) * kHPC + hh);
            }
        }
    } else if (warp < kRowWarp0 + kRowWarps) {
        // ---- row warps: thread = tile row ----
        pdlWait();  // the residual rows come from the preceding kernel
        const int quarter = warp & 3;                // = warp - kRowWarp0: the TMEM lane quarter this warp may read
        const int row = quarter * 32 + lane;
        const uint32_t taddrLane = tmemBase + ((uint32_t)(quarter * 32) << 16);
        const bool rowValid = row < kWinTile * kNT;
        const int rw = rowValid ? row / kNT : 0, rp = rowValid ? row - rw * kNT : 0;
        const int kidx = kKeyStride * rw + rp;       // this row's key slot
        const uint32_t swQ = ((uint32_t)row >> 1) & 3u, swKey = ((uint32_t)kidx >> 1) & 3u;
        uint32_t r[32], r2[32];
        auto cvt8 = [&](const uint32_t* acc, uint32_t biasAddr) {
            float bias[8];
            loadF8(biasAddr, bias);
            uint4 o;
            o.x = packH2(__uint_as_float(acc[0]) + bias[0], __uint_as_float(acc[1]) + bias[1]);
            o.y = packH2(__uint_as_float(acc[2]) + bias[2], __uint_as_float(acc[3]) + bias[3]);
            o.z = packH2(__uint_as_float(acc[4]) + bias[4], __uint_as_float(acc[5]) + bias[5]);
            o.w = packH2(__uint_as_float(acc[6]) + bias[6], __uint_as_float(acc[7]) + bias[7]);
            return o;
        };
        auto e1 = [&](int g) {   // D1 (+ bias) -> Q / K / V^T operands
            const int c = g % kChunks, vb = g & 1;
            const uint32_t bOff = base + kOffBqkv + 4u * (uint32_t)(c * 96);
            const int d = g % F::kD1Bufs;
            const uint32_t tAcc = taddrLane + tD1 + 96u * (uint32_t)d;
            mbarWait(base + (d == 0 ? bD1Full : bD1Full2), (uint32_t)(g / F::kD1Bufs) & 1u);
            tcFenceAfter();
            tmemLd32(tAcc, r);          // Q columns
            tmemLd32(tAcc + 32u, r2);   // K columns
            tmemLdWait();
            mbarWait(base + bQKEmpty, ((uint32_t)g & 1u) ^ 1u);   // both S MMAs of the previous chunk have read Q / K
#pragma unroll
            for (int q = 0; q < 4; ++q) stsV4(base + kOffQ + (uint32_t)row * 64u + (((uint32_t)q ^ swQ) << 4), cvt8(&r[8 * q], bOff + 32u * q));
            if (rowValid) {
#pragma unroll
                for (int q = 0; q < 4; ++q) stsV4(base + kOffK + (uint32_t)kidx * 64u + (((uint32_t)q ^ swKey) << 4), cvt8(&r2[8 * q], bOff + 128u + 32u * q));
            }
            tmemLd32(tAcc + 64u, r);    // V columns (both heads)
            tmemLdWait();
            if constexpr (F::kD1Bufs == 2) {
                tcFenceBefore();
                __syncwarp();
                if (lane == 0) mbarArrive(base + bD1Empty + 8u * d);   // the accumulator is in registers
            }
            mbarWait(base + bVEmpty + 8u * vb, ((uint32_t)(g >> 1) & 1u) ^ 1u);   // both PV MMAs of chunk g - 2 have read this V buffer
            if (rowValid) {
                // V^T[dim][key]: 2-byte stores; the lanes of a warp hold consecutive keys, so a store instruction covers 64 contiguous bytes
                const uint32_t vDst = base + kOffV + (uint32_t)vb * kVBuf + ((uint32_t)kidx >> 6) * 4096u + ((uint32_t)kidx & 7u) * 2u;
                const uint32_t kc16 = ((uint32_t)kidx & 63u) >> 3;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    float bias[8];
                    loadF8(bOff + 4u * (uint32_t)(64 + 8 * q), bias);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const uint32_t dd = (uint32_t)(8 * q + i);
                        stsU16(vDst + dd * 128u + ((kc16 ^ (dd & 7u)) << 4), __half_as_ushort(__float2half_rn(__uint_as_float(r[8 * q + i]) + bias[i])));
                    }
                }
            }
            fenceProxyAsync();
            tcFenceBefore();
            __syncwarp();
            if (lane == 0) mbarArrive(base + bQKFull);
        };
        auto e2 = [&](int g) {   // O -> fp16 -> Oc (A operand of proj) or `out`; the probabilities were normalised by the softmax warps
            const int ob = g & 1;
            mbarWait(base + bOFull + 8u * ob, (uint32_t)(g >> 1) & 1u);
            tcFenceAfter();
            tmemLd32(taddrLane + tO + 32u * ob, r2);
            tmemLdWait();
            tcFenceBefore();
            __syncwarp();
            if (lane == 0) mbarArrive(base + bOEmpty + 8u * ob);
            if constexpr (!F::kProj) {
                const int k = g / kChunks, c = g - k * kChunks;
                const long long tok = tokens(k, row);
                if (tok >= 0) {
                    uint4* dst = reinterpret_cast<uint4*>(a.out + tok * kC + 32 * c);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        uint4 o;
                        o.x = packH2(__uint_as_float(r2[8 * q]), __uint_as_float(r2[8 * q + 1]));
                        o.y = packH2(__uint_as_float(r2[8 * q + 2]), __uint_as_float(r2[8 * q + 3]));
                        o.z = packH2(__uint_as_float(r2[8 * q + 4]), __uint_as_float(r2[8 * q + 5]));
                        o.w = packH2(__uint_as_float(r2[8 * q + 6]), __uint_as_float(r2[8 * q + 7]));
                        dst[q] = o;
                    }
                }
                return;
            }
            mbarWait(base + bOcEmpty, ((uint32_t)g & 1u) ^ 1u);   // proj of the previous chunk has read Oc
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                uint4 o;
                o.x = packH2(__uint_as_float(r2[8 * q]), __uint_as_float(r2[8 * q + 1]));
                o.y = packH2(__uint_as_float(r2[8 * q + 2]), __uint_as_float(r2[8 * q + 3]));
                o.z = packH2(__uint_as_float(r2[8 * q + 4]), __uint_as_float(r2[8 * q + 5]));
                o.w = packH2(__uint_as_float(r2[8 * q + 6]), __uint_as_float(r2[8 * q + 7]));
                stsV4(base + kOffOc + (uint32_t)row * 64u + (((uint32_t)q ^ swQ) << 4), o);
            }
            fenceProxyAsync();
            __syncwarp();
            if (lane == 0) mbarArrive(base + bOcFull);
        };
        for (int g = 0; g <= G + 1; ++g) {
            if (g < G) e1(g);
            if (g >= 2) e2(g - 2);
        }
    } else {
        // ---- softmax warps: thread = query row, one pass (one window's 36 keys) per head ----
        const int quarter = warp & 3;
        const int sub = warp >= 12 ? 1 : 0;          // second warp of quarters 1 / 2
        const int row = quarter * 32 + lane;
        const uint32_t taddrLane = tmemBase + ((uint32_t)(quarter * 32) << 16);
        const bool rowValid = row < kWinTile * kNT;
        const int rw = rowValid ? row / kNT : 0, rp = rowValid ? row - rw * kNT : 0;
        const int passWin = quarter == 0 ? 0 : quarter == 3 ? 2 : quarter - 1 + sub;
        const bool active = rowValid && rw == passWin;
        const int pq = active ? rp : 0;
        const uint32_t swP = (uint32_t)row & 7u;
        uint32_t r[32], s2[8];
        // keys a query may attend to in a shifted block (torchvision's attn_mask: tokens of the last window row / column that were rolled
        // in from the opposite image edge form their own regions); bit j = key position j of the pass's window
        const bool shifted = (a.shiftY | a.shiftX) > 0;
        uint32_t allowLo = 0xffffffffu, allowHi = 0xfu;
        bool maskedWin = false;   // warp-uniform: the pass's window lies on the last window row / column of a shifted block
        auto setMask = [&](long long tile) {
            allowLo = 0xffffffffu;
            allowHi = 0xfu;
            maskedWin = false;
            if (shifted) {
                long long win = tile * kWinTile + passWin;
                const int wx = (int)(win % a.nwx);
                win /= a.nwx;
                const int wy = (int)(win % a.nwy);
                constexpr unsigned long long yLo = 0x3ffffull;        // key rows 0..2
                constexpr unsigned long long xLo = 0x1c71c71c7ull;    // key columns 0..2 of every row
                unsigned long long m = 0xfffffffffull;
                if (a.shiftY > 0 && wy == a.nwy - 1) m &= (pq / kWin < kWin / 2) ? yLo : ~yLo;
                if (a.shiftX > 0 && wx == a.nwx - 1) m &= (pq % kWin < kWin / 2) ? xLo : ~xLo;
                allowLo = (uint32_t)m;
                allowHi = (uint32_t)(m >> 32) & 0xfu;
                maskedWin = (a.shiftY > 0 && wy == a.nwy - 1) || (a.shiftX > 0 && wx == a.nwx - 1);
            }
        };
        mbarWait(base + bW, 0);   // bias tables
        for (int j = 0; j < kHPC * G; ++j) {
            const int b = j & 1;   // score buffer
            if (j % kHeads == 0) setMask((long long)first + (long long)(j / kHeads) * step);
            mbarWait(base + bSFull + 8u * b, (uint32_t)(j >> 1) & 1u);
            tcFenceAfter();
            tmemLd32(taddrLane + tS + 128u * b + (uint32_t)(kKeyStride * passWin), r);
            tmemLd8(taddrLane + tS + 128u * b + (uint32_t)(kKeyStride * passWin) + 32u, s2);
            tmemLdWait();
            tcFenceBefore();
            __syncwarp();
            if (lane == 0) mbarArrive(base + bSEmpty + 8u * b);   // the scores are in registers: the next head of this parity may overwrite them
            const int head = j % kHeads;
            const uint32_t bRow = base + kOffBias + (uint32_t)head * kBiasHead + (uint32_t)pq * (kBiasPitch * 4u);
            float v[kNT];
#pragma unroll
            for (int q = 0; q < kNT / 4; ++q) {
                const uint4 bb = ldsV4(bRow + 16u * q);
                const uint32_t* src = q < 8 ? &r[4 * q] : &s2[4 * q - 32];
                v[4 * q] = __uint_as_float(src[0]) + __uint_as_float(bb.x);
                v[4 * q + 1] = __uint_as_float(src[1]) + __uint_as_float(bb.y);
                v[4 * q + 2] = __uint_as_float(src[2]) + __uint_as_float(bb.z);
                v[4 * q + 3] = __uint_as_float(src[3]) + __uint_as_float(bb.w);
            }
            if (maskedWin) {
#pragma unroll
                for (int i = 0; i < kNT; ++i) {
                    const uint32_t bit = i < 32 ? (allowLo >> i) & 1u : (allowHi >> (i - 32)) & 1u;
                    v[i] = bit ? v[i] : v[i] - 144.26950408889634f;   // -100 in natural-log units
                }
            }
            float m4[4] = {v[0], v[1], v[2], v[3]};
#pragma unroll
            for (int i = 4; i < kNT; ++i) m4[i & 3] = fmaxf(m4[i & 3], v[i]);
            const float mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
            float d4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int i = 0; i < kNT; ++i) {
                v[i] = ex2f(v[i] - mx);
                d4[i & 3] += v[i];
            }
            const float inv = __fdividef(1.f, (d4[0] + d4[1]) + (d4[2] + d4[3]));
            mbarWait(base + bPEmpty, ((uint32_t)j & 1u) ^ 1u);   // PV of the previous head has read P
            if (active) {
                const uint32_t pRow = base + kOffP + (uint32_t)row * 128u;
#pragma unroll
                for (int q = 0; q < 5; ++q) {
                    uint4 o;
                    o.x = packH2(v[8 * q] * inv, v[8 * q + 1] * inv);
                    o.y = packH2(v[8 * q + 2] * inv, v[8 * q + 3] * inv);
                    if (q < 4) {
                        o.z = packH2(v[8 * q + 4] * inv, v[8 * q + 5] * inv);
                        o.w = packH2(v[8 * q + 6] * inv, v[8 * q + 7] * inv);
                    } else {
                        o.z = 0u;   // keys 36..39 of the window's slot: padding
                        o.w = 0u;
                    }
                    const uint32_t ci = (uint32_t)(5 * passWin + q);
                    stsV4(pRow + (ci >> 3) * kPChunk + (((ci & 7u) ^ swP) << 4), o);
                }
            }
            fenceProxyAsync();
            __syncwarp();
            if (lane == 0) mbarArrive(base + bPFull);
        }
    }

    tcFenceBefore();
    __syncthreads();
    if (warp == kIssuerA) {
        tcFenceAfter();
        tmemDealloc(tmemBase, kTmemCols);
    }
}

}  // namespace

struct SwinAttnPlan {
    AttnArgs args;
    int n = 0, c = 0;
};

bool swinAttnSupported(int c, int heads, int window, int h, int w) {
    return (c == 96 || c == 192) && heads == kHeads && window == kWin && h > 0 && w > 0 && h % kWin == 0 && w % kWin == 0;
}

// Host-side operand preparation (done once per block at plan time):
//   wOut [3C][C] fp16 bits: the rows of Wqkv regrouped per 32-channel chunk (q32 | k32 | v32), q rows multiplied by d^-1/2 log2(e);
//   bOut [3C]: the bias in the same order and scale;  relOut [heads][36][44]: relative-position bias times log2(e), rows padded.
void swinAttnPrepare(const uint16_t* wqkv, const float* bqkv, const float* relpos, int c, int heads, std::vector<uint16_t>& wOut, std::vector<float>& bOut,
                     std::vector<float>& relOut) {
    if ((c != 96 && c != 192) || heads != kHeads) throw Error("swin attention: unsupported width");
    const float log2e = 1.4426950408889634f;
    const float qScale = log2e / std::sqrt((float)(c / heads));
    wOut.assign((size_t)3 * c * c, 0);
    bOut.assign((size_t)3 * c, 0.f);
    for (int ch = 0; ch < c / 32; ++ch)
        for (int t = 0; t < 3; ++t)
            for (int i = 0; i < 32; ++i) {
                const int src = t * c + 32 * ch + i, dst = ch * 96 + t * 32 + i;
                const float s = t == 0 ? qScale : 1.f;
                for (int k = 0; k < c; ++k) {
                    __half_raw in;
                    in.x = wqkv[(size_t)src * c + k];
                    const __half_raw o = __float2half_rn(__half2float(__half(in)) * s);
                    wOut[(size_t)dst * c + k] = o.x;
                }
                bOut[dst] = bqkv[src] * s;
            }
    relOut.assign((size_t)heads * kNT * kBiasPitch, 0.f);
    for (int hd = 0; hd < heads; ++hd)
        for (int p = 0; p < kNT; ++p)
            for (int j = 0; j < kNT; ++j) relOut[((size_t)hd * kNT + p) * kBiasPitch + j] = relpos[((size_t)hd * kNT + p) * kNT + j] * log2e;
}

// All pointers are device memory: wqkvR / bqkvR / relposR as produced by swinAttnPrepare.  c = 96: wproj [C][C] fp16 K-major and bproj, x is
// updated in place (out unused); c = 192: the attention output goes to out [n][h][w][C] (wproj / bproj unused: swinAttnFusesProj).
bool swinAttnFusesProj(int c) { return c == 96; }

SwinAttnPlan* swinAttnCreatePlan(__half* x, int n, int h, int w, int c, int heads, int window, int shift, const float* gamma, const float* beta, float eps,
                                 const __half* wqkvR, const float* bqkvR, const __half* wproj, const float* bproj, const float* relposR, __half* out) {
    if (!swinAttnSupported(c, heads, window, h, w)) throw Error("swin attention: unsupported geometry");
    const bool fuseProj = swinAttnFusesProj(c);
    if (fuseProj ? (!wproj || !bproj) : !out) throw Error("swin attention: missing operand");
    if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(wqkvR) | reinterpret_cast<uintptr_t>(wproj) | reinterpret_cast<uintptr_t>(relposR) |
         reinterpret_cast<uintptr_t>(out)) & 15)
        throw Error("swin attention: operands must be 16-byte aligned");
    SwinAttnPlan* plan = new SwinAttnPlan{};
    plan->c = c;
    try {
        encodeMatrixMap2d(&plan->args.tmWqkv, wqkvR, c, 3 * c, 32, 96, false);
        if (fuseProj) encodeMatrixMap2d(&plan->args.tmWproj, wproj, c, c, 32, c, false);
        else plan->args.tmWproj = plan->args.tmWqkv;
    } catch (...) {
        delete plan;
        throw;
    }
    AttnArgs& a = plan->args;
    a.x = x;
    a.bqkv = bqkvR;
    a.bproj = bproj;
    a.gamma = gamma;
    a.beta = beta;
    a.relpos = relposR;
    a.out = out;
    a.eps = eps;
    a.h = h;
    a.w = w;
    // torchvision drops the shift in a dimension the window covers (shifted_window_attention)
    a.shiftY = window >= h ? 0 : shift;
    a.shiftX = window >= w ? 0 : shift;
    a.nwx = w / window;
    a.nwy = h / window;
    a.windows = 0;
    plan->n = n;
    return plan;
}

void swinAttnDestroyPlan(SwinAttnPlan* plan) { delete plan; }

const char* swinAttnDescribe(const SwinAttnPlan* plan, char* buf, int cap) {
    std::snprintf(buf, cap, "swin-attn fused LN+QKV+window-attention%s (tcgen05) c=%d heads=%d window=%d shift=%d rows=%d windows/tile=%d weights=%s smem=%u",
                  plan->c == 96 ? "+proj+residual" : "", plan->c, kHeads, kWin, plan->args.shiftY > plan->args.shiftX ? plan->args.shiftY : plan->args.shiftX, kRows, kWinTile,
                  plan->c == 96 ? "resident" : "streamed", plan->c == 96 ? ACfg<96>::kSmem : ACfg<192>::kSmem);
    return buf;
}

void swinAttnLaunch(const SwinAttnPlan* plan, cudaStream_t s, int nImages) {
    static bool attrSet[64] = {};
    static int sms[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    if (!attrSet[dev]) {
        cudaFuncSetAttribute(swin_attn_kernel<96>, cudaFuncAttributeMaxDynamicSharedMemorySize, ACfg<96>::kSmem);
        cudaFuncSetAttribute(swin_attn_kernel<192>, cudaFuncAttributeMaxDynamicSharedMemorySize, ACfg<192>::kSmem);
        cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev);
        if (sms[dev] <= 0) sms[dev] = 148;
        attrSet[dev] = true;
    }
    if (nImages <= 0) return;
    AttnArgs a = plan->args;
    a.windows = (long long)nImages * a.nwx * a.nwy;
    const long long tiles = (a.windows + kWinTile - 1) / kWinTile;
    const dim3 grid((unsigned)(tiles < sms[dev] ? tiles : sms[dev]));
    const cudaError_t e = plan->c == 96 ? launchPdl(swin_attn_kernel<96>, grid, dim3(kAttnThreads), ACfg<96>::kSmem, s, a)
                                        : launchPdl(swin_attn_kernel<192>, grid, dim3(kAttnThreads), ACfg<192>::kSmem, s, a);
    if (e != cudaSuccess) throw Error(std::string("swin attention launch: ") + cudaGetErrorString(e));
}

}  // namespace w2x
