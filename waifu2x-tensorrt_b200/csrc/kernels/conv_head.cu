// Image head of the CUNet/UpCUNet family (the last conv of unet2: 3x3, 64 -> 3 channels, valid padding) fused with
// the cropped residual add and the [0, 1] clamp (EPI_FINAL).
//
// Why its own kernel: with 3 output channels the layer is bound by how often each input pixel is read, not by math.  A per-tap
// GEMM (the generic kernels) streams every pixel through the tensor-core operand path nine times (one shifted view per tap).
// Here the nine taps sit in the N dimension instead: one pass computes, for every pixel of an 18 x 18 input patch, the 27 partial
// products P[tap*3 + co][pixel] = sum_c w[co][tap][c] * x[pixel][c] (M = 128 pixels, N = 27 -> 32, K = 64: four tcgen05.mma per
// 128 pixels, fp32 accumulators in TMEM), parks them in shared memory and each output pixel then adds its nine shifted partials.
// Every input byte is read from HBM/L2 once and from shared memory once; algorithmic traffic = 128 B/pixel in + 8 B/pixel out (+ 8 B skip).
//
// One persistent CTA per SM (256 threads) walks 16 x 16 output tiles with a four-stage TMA ring of 18 x 18 input patches
// (128B-swizzled, three patch loads = 124 KB in flight per SM, out-of-image pixels zero-filled by TMA).  The patch is the A operand
// as TMA wrote it; the first version used mma.sync (336 HMMA per tile) and was bound by the legacy tensor path's issue rate
// (profiles/r01_hmma_issue_rate.txt: 5-9 SM cycles per HMMA at 8-16 warps), twelve N = 32 UMMAs per tile replace them.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "../hostutil.h"
#include "conv_params.h"
#include "launch.h"
#include "sm100_common.cuh"

namespace w2x {
using namespace sm100;
namespace {

constexpr int kTileY = 16, kTileX = 16;                  // output pixels per tile
constexpr int kPatchY = kTileY + 2, kPatchX = kTileX + 2;  // 18 x 18 input pixels
constexpr int kPatchPx = kPatchY * kPatchX;              // 324
constexpr int kBlocks = (kPatchPx + 15) / 16;            // 21 M-blocks of 16 pixels (336 rows; rows 324.. are never gathered)
constexpr int kPitch = 356;                              // floats per row of P^T; 356 % 32 == 4 makes the fragment stores conflict-free
constexpr int kCols = 27;                                // 9 taps x 3 channels
constexpr int kPatchBytes = kBlocks * 16 * 128;          // 43008
constexpr int kStages = 4;
constexpr int kHeadThreads = 256;
constexpr int kPatchTx = kPatchPx * 128;                 // bytes one TMA patch load delivers
constexpr int kPtBytes = kCols * kPitch * 4;             // 38448
constexpr int kWBytes = 32 * 128;                        // taps-in-N weight matrix [32 n][64 k] fp16, K-major, 128B-swizzled
constexpr int kMBlocks = (kPatchPx + 127) / 128;         // 3 UMMA row blocks of 128 patch pixels (rows 324.. are never gathered)
constexpr int kHeadSmem = 1024 + kStages * kPatchBytes + kWBytes + kPtBytes + 128;  // alignment slack + ring + W + P^T + mbarriers

struct HeadTile {
    int img, y0, x0;
};
__device__ __forceinline__ HeadTile headTile(int tile, int tilesX, int tilesY) {
    const int tx = tile % tilesX;
    tile /= tilesX;
    const int ty = tile % tilesY;
    return HeadTile{tile / tilesY, ty * kTileY, tx * kTileX};
}

__global__ void __launch_bounds__(kHeadThreads, 1) conv_head_kernel(const ConvParams p, const __grid_constant__ CUtensorMap tmIn, int tilesX,
                                                                    int tilesY, int numTiles) {
    extern __shared__ uint8_t smemRaw[];
    const uint32_t rawAddr = smemU32(smemRaw);
    const uint32_t patch0 = (rawAddr + 1023u) & ~1023u;  // 128B-swizzle atoms need a 1024-byte aligned ring
    uint8_t* sm = smemRaw + (patch0 - rawAddr);
    const uint32_t wsm = patch0 + kStages * kPatchBytes;  // 1024-aligned: kPatchBytes is a multiple of 1024
    float* pt = reinterpret_cast<float*>(sm + kStages * kPatchBytes + kWBytes);
    const uint32_t bar0 = wsm + kWBytes + kPtBytes;       // kStages patch barriers, then the MMA barrier, then the TMEM slot
    const uint32_t barMma = bar0 + 8u * kStages;
    volatile uint32_t* tmemSlot = reinterpret_cast<volatile uint32_t*>(sm + kStages * kPatchBytes + kWBytes + kPtBytes + 8 * kStages + 8);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile0 = blockIdx.x, stride = gridDim.x;

    auto loadPatch = [&](int tile, int slot) {  // thread 0 only
        if (tile >= numTiles) return;
        const HeadTile ht = headTile(tile, tilesX, tilesY);
        mbarExpectTx(bar0 + 8u * slot, kPatchTx);
        tmaLoad5d(patch0 + (uint32_t)slot * kPatchBytes, &tmIn, bar0 + 8u * slot, 0, ht.x0, 0, ht.y0, ht.img);
    };
    pdlLaunchDependents();
    if (tid == 0) {
        for (int st = 0; st < kStages; ++st) mbarInit(bar0 + 8u * st, 1);
        mbarInit(barMma, 1);
        mbarInitFence();
        tmaPrefetchDesc(&tmIn);
    }
    if (warp == 0) tmemAlloc(smemU32((const void*)tmemSlot), 128);  // 3 row blocks x 32 fp32 columns
    // B operand (constants): row n = tap*3 + co of the taps-in-N weight matrix, 64 input channels along K, zero rows for n >= 27;
    // K-major SWIZZLE_128B layout: row n is 128 bytes, 16-byte chunk c of the row sits at chunk c ^ (n & 7)
    for (int i = tid; i < 32 * 8; i += kHeadThreads) {
        const int n = i >> 3, c = i & 7;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (n < kCols) v = *reinterpret_cast<const uint4*>(p.w + (long long)(n % 3) * p.ktot + (n / 3) * 64 + c * 8);
        stsV4(wsm + (uint32_t)n * 128u + ((uint32_t)(c ^ (n & 7)) << 4), v);
    }
    fenceProxyAsync();  // generic-proxy stores -> visible to the tensor core's operand reads
    const float bias0 = __ldg(p.bias + 0), bias1 = __ldg(p.bias + 1), bias2 = __ldg(p.bias + 2);
    const int oy = tid >> 4, ox = tid & 15;
    int gatherOff[9];
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) gatherOff[tap] = 3 * tap * kPitch + (oy + p.tap[tap].dy) * kPatchX + ox + p.tap[tap].dx;
    tcFenceBefore();
    __syncthreads();  // barriers, TMEM slot and the weight tile are visible to every thread
    tcFenceAfter();
    const uint32_t tmemBase = *tmemSlot;
    const uint32_t idesc = instrDescF16(128, 32);
    const uint32_t hi = descHi(1024, 2);  // 8-row groups 1024 bytes apart, 128-byte swizzle
    // this thread's part of the accumulator: TMEM lanes of its warp's quarter, 16 of the 32 columns
    const int quarter = warp & 3, colHalf = warp >> 2;
    const uint32_t tRow = tmemBase + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(colHalf * 16);

    pdlWait();  // everything above touches constants only; the input and residual tensors come from preceding kernels
    if (tid == 0)
        for (int st = 0; st < kStages - 1; ++st) loadPatch(tile0 + st * stride, st);

    int stage = 0;
    uint32_t phase = 0, mmaPhase = 0;
    for (int tile = tile0; tile < numTiles; tile += stride) {
        // refill the slot freed by the previous tile (its MMAs completed before that iteration's accumulator read-out)
        if (tid == 0) loadPatch(tile + (kStages - 1) * stride, (stage + kStages - 1) % kStages);
        // this tile's residual pixel (z1 crop): issued now, consumed after the MMAs
        const HeadTile ht = headTile(tile, tilesX, tilesY);
        const int y = ht.y0 + oy, x = ht.x0 + ox;
        const bool valid = y < p.gy && x < p.gx;
        Half4 sv{};
        if (valid) sv = *reinterpret_cast<const Half4*>(p.skip + (((long long)ht.img * p.skip_h + y + p.skip_off) * p.skip_w + x + p.skip_off) * p.skip_c);
        if (warp == 0) {
            mbarWait(bar0 + 8u * stage, phase);  // patch complete
            tcFenceAfter();
            if (electOne()) {
                // P[pixel][tap*3 + co] for 3 x 128 patch pixels: four K = 16 steps over the 64 channels each
                const uint32_t patch = patch0 + (uint32_t)stage * kPatchBytes;
#pragma unroll
                for (int mb = 0; mb < kMBlocks; ++mb)
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)
                        umma(tmemBase + (uint32_t)(mb * 32), makeDesc(patch + (uint32_t)mb * 16384u + (uint32_t)ks * 32u, hi), makeDesc(wsm + (uint32_t)ks * 32u, hi), idesc,
                             ks != 0 ? 1u : 0u);
                tcCommit(barMma);
            }
            __syncwarp();
        }
        mbarWait(barMma, mmaPhase);
        tcFenceAfter();
        __syncthreads();  // every thread has finished gathering the previous tile from P^T
        // accumulators -> P^T[n][pixel] (fp32): lanes are consecutive pixels, so the stores are conflict-free
#pragma unroll
        for (int mb = 0; mb < kMBlocks; ++mb) {
            uint32_t r[32];
            tmemLd16(tRow + (uint32_t)(mb * 32), r);
            tmemLdWait();
            const int px = mb * 128 + quarter * 32 + lane;
            if (px < kBlocks * 16) {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int n = colHalf * 16 + j;
                    if (n < kCols) pt[n * kPitch + px] = __uint_as_float(r[j]);
                }
            }
        }
        tcFenceBefore();
        __syncthreads();  // P^T complete; the accumulator and this patch slot may be overwritten

        // one output pixel per thread: nine shifted partials per channel + bias + cropped residual, clamp, 8-byte store
        float acc0 = bias0, acc1 = bias1, acc2 = bias2;
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
            const float* src = pt + gatherOff[tap];
            acc0 += src[0];
            acc1 += src[kPitch];
            acc2 += src[2 * kPitch];
        }
        if (valid) {
            const float2 s0 = __half22float2(sv.a), s1 = __half22float2(sv.b);
            Half4 h{__floats2half2_rn(fminf(fmaxf(acc0 + s0.x, 0.f), 1.f), fminf(fmaxf(acc1 + s0.y, 0.f), 1.f)),
                    __floats2half2_rn(fminf(fmaxf(acc2 + s1.x, 0.f), 1.f), 0.f)};
            *reinterpret_cast<Half4*>(p.out + (((long long)ht.img * p.out_h + y) * p.out_w + x) * p.out_c) = h;
        }
        if (++stage == kStages) { stage = 0; phase ^= 1u; }
        mmaPhase ^= 1u;
    }
    tcFenceBefore();
    __syncthreads();
    if (warp == 0) {
        tcFenceAfter();
        tmemDealloc(tmemBase, 128);
    }
}

// ---- ConvTranspose 4x4 stride 2 pad 3, 64 -> 3 channels (the head of unet1 in the 2x models) ---------------------------------
// Same idea as the image head: output pixel (2y-1+py, 2x-1+px) is the sum over the 2x2 input neighbourhood (y+wy, x+wx) of
// w[wy][wx][py][px] . x, so one pass computes P[pixel][(wy,wx), (py,px), co] for every input pixel of an 18x18 patch (N = 4 taps
// x 4 phases x 4 channels = 64, K = 64: four N = 64 UMMAs per 128 pixels) and each 2x2 output block gathers its 4 x 4 partials.
constexpr int kUpStages = 3;
constexpr int kUpRows = 48;                               // P^T rows kept: (tap, phase, co < 3)
constexpr int kUpWBytes = 64 * 128;
constexpr int kUpPtBytes = kUpRows * kPitch * 4;          // 68352
constexpr int kUpSmem = 1024 + kUpStages * kPatchBytes + kUpWBytes + kUpPtBytes + 128;

__global__ void __launch_bounds__(kHeadThreads, 1) conv_up4_kernel(const ConvParams p, const __grid_constant__ CUtensorMap tmIn, int tilesX,
                                                                   int tilesY, int numTiles) {
    extern __shared__ uint8_t smemRaw[];
    const uint32_t rawAddr = smemU32(smemRaw);
    const uint32_t patch0 = (rawAddr + 1023u) & ~1023u;
    uint8_t* sm = smemRaw + (patch0 - rawAddr);
    const uint32_t wsm = patch0 + kUpStages * kPatchBytes;
    float* pt = reinterpret_cast<float*>(sm + kUpStages * kPatchBytes + kUpWBytes);
    const uint32_t bar0 = wsm + kUpWBytes + kUpPtBytes;
    const uint32_t barMma = bar0 + 8u * kUpStages;
    volatile uint32_t* tmemSlot = reinterpret_cast<volatile uint32_t*>(sm + kUpStages * kPatchBytes + kUpWBytes + kUpPtBytes + 8 * kUpStages + 8);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile0 = blockIdx.x, stride = gridDim.x;

    auto loadPatch = [&](int tile, int slot) {  // thread 0 only
        if (tile >= numTiles) return;
        const HeadTile ht = headTile(tile, tilesX, tilesY);
        mbarExpectTx(bar0 + 8u * slot, kPatchTx);
        tmaLoad5d(patch0 + (uint32_t)slot * kPatchBytes, &tmIn, bar0 + 8u * slot, 0, ht.x0, 0, ht.y0, ht.img);
    };
    pdlLaunchDependents();
    if (tid == 0) {
        for (int st = 0; st < kUpStages; ++st) mbarInit(bar0 + 8u * st, 1);
        mbarInit(barMma, 1);
        mbarInitFence();
        tmaPrefetchDesc(&tmIn);
    }
    if (warp == 0) tmemAlloc(smemU32((const void*)tmemSlot), 256);  // 3 row blocks x 64 fp32 columns
    // B operand: row n = tap*16 + phase*4 + co holds w[phase*4 + co][tap*64 .. tap*64 + 64) of the packed [16][256] matrix
    for (int i = tid; i < 64 * 8; i += kHeadThreads) {
        const int n = i >> 3, c = i & 7;
        const uint4 v = *reinterpret_cast<const uint4*>(p.w + (long long)(n & 15) * p.ktot + (n >> 4) * 64 + c * 8);
        stsV4(wsm + (uint32_t)n * 128u + ((uint32_t)(c ^ (n & 7)) << 4), v);
    }
    fenceProxyAsync();
    float biasQ[4][3];  // packed per (phase, channel) column
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int co = 0; co < 3; ++co) biasQ[q][co] = __ldg(p.bias + q * 4 + co);
    const int gy_ = tid >> 4, gx_ = tid & 15;  // this thread's GEMM pixel inside the 16 x 16 tile
    tcFenceBefore();
    __syncthreads();
    tcFenceAfter();
    const uint32_t tmemBase = *tmemSlot;
    const uint32_t idesc = instrDescF16(128, 64);
    const uint32_t hi = descHi(1024, 2);
    const int quarter = warp & 3, colHalf = warp >> 2;
    const uint32_t tRow = tmemBase + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(colHalf * 32);

    pdlWait();
    if (tid == 0)
        for (int st = 0; st < kUpStages - 1; ++st) loadPatch(tile0 + st * stride, st);

    int stage = 0;
    uint32_t phase = 0, mmaPhase = 0;
    for (int tile = tile0; tile < numTiles; tile += stride) {
        if (tid == 0) loadPatch(tile + (kUpStages - 1) * stride, (stage + kUpStages - 1) % kUpStages);
        const HeadTile ht = headTile(tile, tilesX, tilesY);
        if (warp == 0) {
            mbarWait(bar0 + 8u * stage, phase);
            tcFenceAfter();
            if (electOne()) {
                const uint32_t patch = patch0 + (uint32_t)stage * kPatchBytes;
#pragma unroll
                for (int mb = 0; mb < kMBlocks; ++mb)
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)
                        umma(tmemBase + (uint32_t)(mb * 64), makeDesc(patch + (uint32_t)mb * 16384u + (uint32_t)ks * 32u, hi), makeDesc(wsm + (uint32_t)ks * 32u, hi), idesc,
                             ks != 0 ? 1u : 0u);
                tcCommit(barMma);
            }
            __syncwarp();
        }
        mbarWait(barMma, mmaPhase);
        tcFenceAfter();
        __syncthreads();  // every thread has finished gathering the previous tile from P^T
#pragma unroll
        for (int mb = 0; mb < kMBlocks; ++mb) {
            uint32_t r[32];
            tmemLd32(tRow + (uint32_t)(mb * 64), r);
            tmemLdWait();
            const int px = mb * 128 + quarter * 32 + lane;
            if (px < kBlocks * 16) {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int n = colHalf * 32 + j;            // tap*16 + phase*4 + co
                    if ((n & 3) < 3) pt[((n >> 4) * 12 + ((n >> 2) & 3) * 3 + (n & 3)) * kPitch + px] = __uint_as_float(r[j]);
                }
            }
        }
        tcFenceBefore();
        __syncthreads();  // P^T complete; the accumulator and this patch slot may be overwritten

        // one GEMM pixel = one 2 x 2 output block per thread
        const int y = ht.y0 + gy_, x = ht.x0 + gx_;
        if (y < p.gy && x < p.gx) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int oy = 2 * y - 1 + (q >> 1), ox = 2 * x - 1 + (q & 1);
                float acc0 = biasQ[q][0], acc1 = biasQ[q][1], acc2 = biasQ[q][2];
#pragma unroll
                for (int tap = 0; tap < 4; ++tap) {
                    const float* src = pt + (tap * 12 + q * 3) * kPitch + (gy_ + (tap >> 1)) * kPatchX + gx_ + (tap & 1);
                    acc0 += src[0];
                    acc1 += src[kPitch];
                    acc2 += src[2 * kPitch];
                }
                if (oy >= 0 && ox >= 0 && oy < p.out_h && ox < p.out_w) {
                    Half4 h{__floats2half2_rn(acc0, acc1), __floats2half2_rn(acc2, 0.f)};
                    *reinterpret_cast<Half4*>(p.out + (((long long)ht.img * p.out_h + oy) * p.out_w + ox) * p.out_c) = h;
                }
            }
        }
        if (++stage == kUpStages) { stage = 0; phase ^= 1u; }
        mmaPhase ^= 1u;
    }
    tcFenceBefore();
    __syncthreads();
    if (warp == 0) {
        tcFenceAfter();
        tmemDealloc(tmemBase, 256);
    }
}

}  // namespace

struct HeadPlan {
    ConvParams p;
    CUtensorMap tmIn;
    bool up4;  // ConvTranspose 4x4 s2 p3 head (conv_up4_kernel) instead of the 3x3 image head
};

static bool convUp4Supported(const ConvParams& p) {
    if (!(p.mode == EPI_UP4 && p.ntaps == 4 && p.cin == 64 && p.ktot == 256 && p.npad == 16 && p.out_c == 4 && p.w_img_stride == 0 && p.sx == 64 &&
          p.dimc == 64 && p.dimz == 1 && p.slope == 1.f))
        return false;
    for (int i = 0; i < 4; ++i)
        if (p.tap[i].c0 != 0 || p.tap[i].dz != 0 || p.tap[i].dx != (i & 1) || p.tap[i].dy != (i >> 1)) return false;
    return true;
}

bool convHeadSupported(const ConvParams& p) {
    if (convUp4Supported(p)) return true;
    if (!(p.mode == EPI_FINAL && p.is3x3 && p.cin == 64 && p.ntaps == 9 && p.ktot == 576 && p.w_img_stride == 0 && p.skip && p.skip_c == 4 &&
          p.out_c == 4 && p.sx == 64 && p.dimc == 64 && p.dimz == 1))
        return false;
    for (int i = 0; i < 9; ++i)
        if (p.tap[i].c0 != 0 || p.tap[i].dz != 0 || p.tap[i].dx < 0 || p.tap[i].dx > 2 || p.tap[i].dy < 0 || p.tap[i].dy > 2) return false;
    return true;
}

HeadPlan* convHeadCreatePlan(const ConvParams& p) {
    if (!convHeadSupported(p)) throw Error("image-head kernel does not support this layer shape");
    HeadPlan* plan = new HeadPlan{p, {}, convUp4Supported(p)};
    try {
        encodeActivationMap5d(&plan->tmIn, p, kPatchX, kPatchY);
    } catch (...) {
        delete plan;
        throw;
    }
    return plan;
}

void convHeadDestroyPlan(HeadPlan* plan) { delete plan; }

void launchConvHead(const HeadPlan* plan, cudaStream_t s, __half* outOverride, int nImages) {
    static bool attrSet[64] = {};
    static int sms[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    if (!attrSet[dev]) {
        cudaFuncSetAttribute(conv_head_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kHeadSmem);
        cudaFuncSetAttribute(conv_up4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kUpSmem);
        cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev);
        if (sms[dev] <= 0) sms[dev] = 148;
        attrSet[dev] = true;
    }
    ConvParams p = plan->p;
    if (outOverride) p.out = outOverride;
    if (nImages > 0) p.gn = nImages;
    const int tilesX = (p.gx + kTileX - 1) / kTileX, tilesY = (p.gy + kTileY - 1) / kTileY;
    const int numTiles = tilesX * tilesY * p.gn;
    const dim3 grid(numTiles < sms[dev] ? numTiles : sms[dev]);
    if (plan->up4) launchPdl(conv_up4_kernel, grid, dim3(kHeadThreads), kUpSmem, s, p, plan->tmIn, tilesX, tilesY, numTiles);
    else launchPdl(conv_head_kernel, grid, dim3(kHeadThreads), kHeadSmem, s, p, plan->tmIn, tilesX, tilesY, numTiles);
}

}  // namespace w2x
