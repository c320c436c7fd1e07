// Image head of the CUNet/UpCUNet family (the last conv of unet2: 3x3, 64 -> 3 channels, valid padding) fused with
// the cropped residual add and the [0, 1] clamp (EPI_FINAL).
//
// Why its own kernel: with 3 output channels the layer is bound by how often each input pixel is read, not by math.  A per-tap
// GEMM (the generic kernels) streams every pixel through the tensor-core operand path nine times (one shifted view per tap).
// Here the nine taps sit in the N dimension instead: one pass computes, for every pixel of a 10 x 18 input patch, the 27 partial
// products P[tap*3 + co][pixel] = sum_c w[co][tap][c] * x[pixel][c] (M = pixels, N = 27 -> 32, K = 64: sixteen m16n8k16 MMAs per
// 16 pixels), parks them in shared memory and each output pixel then adds its nine shifted partials in fp32.  Every input byte
// is read from HBM/L2 once and from shared memory once; algorithmic traffic = 128 B/pixel in + 8 B/pixel out (+ 8 B skip).
//
// One persistent CTA per SM (256 threads) walks 16 x 16 output tiles with a four-stage TMA ring of 18 x 18 input patches
// (128B-swizzled, three patch loads = 124 KB in flight per SM, out-of-image pixels zero-filled by TMA); mma.sync is the right tool
// for the math: N = 32 would use a quarter of a tcgen05 instruction's width and the MMA time is < 10 % of the memory time.
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
constexpr int kHeadSmem = 1024 + kStages * kPatchBytes + kPtBytes + 64;  // alignment slack + ring + P^T + mbarriers

__device__ __forceinline__ void ldmatrixX4(uint32_t addr, uint32_t (&r)[4]) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

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
    float* pt = reinterpret_cast<float*>(sm + kStages * kPatchBytes);
    const uint32_t bar0 = patch0 + kStages * kPatchBytes + kPtBytes;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
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
        mbarInitFence();
        tmaPrefetchDesc(&tmIn);
    }

    // B fragments (constants, loaded before the dependency wait): column n = tap*3 + co of the taps-in-N weight matrix, zero for n >= 27
    uint32_t bf[4][4][2];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
        const int n = 8 * nt + g;
        const int tap = n / 3, co = n - 3 * tap;
        const __half* wrow = p.w + (long long)co * p.ktot + tap * 64 + 2 * t;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            bf[nt][ks][0] = n < kCols ? *reinterpret_cast<const uint32_t*>(wrow + 16 * ks) : 0u;
            bf[nt][ks][1] = n < kCols ? *reinterpret_cast<const uint32_t*>(wrow + 16 * ks + 8) : 0u;
        }
    }
    const float bias0 = __ldg(p.bias + 0), bias1 = __ldg(p.bias + 1), bias2 = __ldg(p.bias + 2);
    const int oy = tid >> 4, ox = tid & 15;
    int gatherOff[9];
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) gatherOff[tap] = 3 * tap * kPitch + (oy + p.tap[tap].dy) * kPatchX + ox + p.tap[tap].dx;

    pdlWait();  // everything above touches constants only; the input and residual tensors come from preceding kernels
    if (tid == 0)
        for (int st = 0; st < kStages - 1; ++st) loadPatch(tile0 + st * stride, st);
    __syncthreads();  // barrier init visible to all threads

    int stage = 0;
    uint32_t phase = 0;
    for (int tile = tile0; tile < numTiles; tile += stride) {
        // refill the slot freed by the previous tile (its MMAs finished before that iteration's second barrier)
        if (tid == 0) loadPatch(tile + (kStages - 1) * stride, (stage + kStages - 1) % kStages);
        // this tile's residual pixel (z1 crop): issued now, consumed after the MMAs
        const HeadTile ht = headTile(tile, tilesX, tilesY);
        const int y = ht.y0 + oy, x = ht.x0 + ox;
        const bool valid = y < p.gy && x < p.gx;
        Half4 sv{};
        if (valid) sv = *reinterpret_cast<const Half4*>(p.skip + (((long long)ht.img * p.skip_h + y + p.skip_off) * p.skip_w + x + p.skip_off) * p.skip_c);
        mbarWait(bar0 + 8u * stage, phase);  // patch complete
        __syncthreads();                     // every thread has finished gathering the previous tile from P^T

        // partial products: 16-pixel blocks round-robin over the warps; P^T[n][pixel] <- C fragments
        const uint32_t patch = patch0 + (uint32_t)stage * kPatchBytes;
        for (int b = warp; b < kBlocks; b += kHeadThreads / 32) {
            float d[4][4] = {};
            const int row = b * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
            const uint32_t rowAddr = patch + (uint32_t)row * 128u;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                uint32_t a[4];
                ldmatrixX4(rowAddr + ((uint32_t)((2 * ks + (lane >> 4)) ^ (row & 7)) << 4), a);
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) mma16816(d[nt], a, bf[nt][ks][0], bf[nt][ks][1]);
            }
            float* dst = pt + b * 16 + g;
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                const int n = 8 * nt + 2 * t;
                if (n < kCols) { dst[n * kPitch] = d[nt][0]; dst[n * kPitch + 8] = d[nt][2]; }
                if (n + 1 < kCols) { dst[(n + 1) * kPitch] = d[nt][1]; dst[(n + 1) * kPitch + 8] = d[nt][3]; }
            }
        }
        __syncthreads();  // P^T complete; this patch slot may be refilled by the load issued in the next iteration

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
    }
}

}  // namespace

struct HeadPlan {
    ConvParams p;
    CUtensorMap tmIn;
};

bool convHeadSupported(const ConvParams& p) {
    if (!(p.mode == EPI_FINAL && p.is3x3 && p.cin == 64 && p.ntaps == 9 && p.ktot == 576 && p.w_img_stride == 0 && p.skip && p.skip_c == 4 &&
          p.out_c == 4 && p.sx == 64 && p.dimc == 64 && p.dimz == 1))
        return false;
    for (int i = 0; i < 9; ++i)
        if (p.tap[i].c0 != 0 || p.tap[i].dz != 0 || p.tap[i].dx < 0 || p.tap[i].dx > 2 || p.tap[i].dy < 0 || p.tap[i].dy > 2) return false;
    return true;
}

HeadPlan* convHeadCreatePlan(const ConvParams& p) {
    if (!convHeadSupported(p)) throw Error("image-head kernel does not support this layer shape");
    HeadPlan* plan = new HeadPlan{p, {}};
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
        cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev);
        if (sms[dev] <= 0) sms[dev] = 148;
        attrSet[dev] = true;
    }
    ConvParams p = plan->p;
    if (outOverride) p.out = outOverride;
    if (nImages > 0) p.gn = nImages;
    const int tilesX = (p.gx + kTileX - 1) / kTileX, tilesY = (p.gy + kTileY - 1) / kTileY;
    const int numTiles = tilesX * tilesY * p.gn;
    launchPdl(conv_head_kernel, dim3(numTiles < sms[dev] ? numTiles : sms[dev]), dim3(kHeadThreads), kHeadSmem, s, p, plan->tmIn, tilesX, tilesY, numTiles);
}

}  // namespace w2x
