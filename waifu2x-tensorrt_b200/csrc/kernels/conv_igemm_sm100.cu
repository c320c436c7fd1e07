// Convolutions on the sm_100a tensor cores (tcgen05.mma, accumulators in TMEM, operands staged by TMA).
//
// These two kernels are the dense model hot path: together they replace the opaque TensorRT
// engine the reference enqueues at /root/reference/src/tensorrt/img2img_infer.cpp:80.
//
//   conv3x3_patch_kernel  3x3 valid convolutions with cin in {64,128}: the layer's weights for one 64-wide output slice
//                         stay RESIDENT in shared memory for the CTA's lifetime; each 16x8-pixel output tile loads ONE
//                         18x10-pixel input patch per 64-channel chunk by TMA (128B swizzle) and all nine taps are MMA'd
//                         from SHIFTED VIEWS of that patch (descriptor start moved by whole 128-byte pixel rows, 8-row groups
//                         SBO = 10 pixels apart).  L2->SM traffic per output pixel drops ~6x versus per-tap loads.
//   igemm_kernel          generic implicit GEMM: per-tap TMA box loads of A and a B tile per K block.  Used for the
//                         2x2/s2 convs, the ConvTranspose layers (depth-to-space epilogue), cin=32 and the 256-wide layers.
//
//   D[128 pixels, N] (fp32, TMEM)  +=  A[128 pixels, K] (fp16 smem, K-major, swizzled)  x  B[N, K] (fp16 smem, K-major)
//
// Shared structure: persistent CTAs (one per SM), warp-specialised: warp 0 = TMA producer, warp 1 = TMEM allocator +
// single-thread tcgen05.mma issuer, warps 2..5 = epilogue.  Two TMEM accumulator buffers (epilogue of tile i overlaps the
// MMAs of tile i+1), an mbarrier smem ring between TMA and MMA.  Epilogue: tcgen05.ld -> bias + LeakyReLU (+ skip tile that
// was TMA-loaded into the staging buffer) -> 128B-swizzled smem staging -> TMA store (coalesced, clips image edges); the
// depth-to-space of ConvTranspose 2x2 is expressed in the store's 5-D tensor map.  3-channel heads store directly.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "../hostutil.h"
#include "../model_pack.h"
#include "conv_params.h"
#include "launch.h"
#include "sm100_common.cuh"

namespace w2x {

using namespace sm100;

// timing experiments that switch parts of a kernel off (W2X_DBG) exist only in the development build
#ifdef W2X_DEV
#define W2X_DBG_ON(a, bit) (((a).dbg & (bit)) != 0)
// cycle counters of the patch kernel's roles: 0 MMA warp total, 1 wait accumulator-empty, 2 wait patch-full, 3 issue + commit,
// 4 producer total, 5 producer wait slot-empty, 6 epilogue group 0 total, 7 wait accumulator-full, 8 TMEM load + math + staging,
// 9 barriers + store issue, 10 tiles of this CTA, 11 %globaltimer ns of the MMA loop
#define W2X_PROF_DECL(name) long long name = 0
#define W2X_PROF_T(var) const long long var = clock64()
#define W2X_PROF_ADD(acc, t0) acc += clock64() - (t0)
// event trace of CTA 0 (first kTraceTiles tiles): prof[16 * grid + tile * 8 + slot] = clock64()
constexpr int kTraceTiles = 24;
#define W2X_TRACE(a, tile, slot) do { if ((a).prof && blockIdx.x == 0 && (tile) < kTraceTiles) (a).prof[16ll * gridDim.x + (tile) * 8 + (slot)] = clock64(); } while (0)
#else
#define W2X_TRACE(a, tile, slot)
#define W2X_DBG_ON(a, bit) false
#define W2X_PROF_DECL(name)
#define W2X_PROF_T(var)
#define W2X_PROF_ADD(acc, t0)
#endif

struct ConvArgs {
    CUtensorMap tmA;     // activation loads (igemm: per-tap box; patch kernel: 18x10 patch box)
    CUtensorMap tmB;     // weights [npad][ktot]
    CUtensorMap tmOut;   // TMA-store view of the output (useTma)
    CUtensorMap tmSkip;  // TMA-load view of the skip tensor (hasSkip)
    CUtensorMap tmRgb;   // fused first layer: [img][H][W*4] view of the NHWC4 input, 20 x 48-element boxes (no swizzle)
    ConvParams p;
    int kc, bn, bw, bh, bwShift;
    int tilesX, tilesY, tilesN, totalTiles;
    int stages, cchunks, kblocks;
    uint32_t idesc, tmemCols, bytesA, bytesB, descHiA, descHiB;
    int useTma, nsub, nbuf, hasSkip;
    int twoPerSm;  // staged token-wise layer planned for two CTAs per SM (EPI_K_STAGED2)
    int nAcc;  // TMEM accumulator buffers: 2, or 4 for the two-group epilogue (each group alternates between two of its own)
    int dbg;  // timing experiments only (W2X_DBG): 1 = epilogue skips math + staging, 2 = no output store, 4 = no MMAs, 16 = no activation loads
    uint32_t stageStride, wBytes, stagingBytes;
    int nSplit;
    int staged;                 // EPI_K_STAGED: generic-proxy staging + coalesced copy-out (any N, residual, GELU)
    uint32_t stagedPitch, stagedBuf;
    uint32_t headerBytes;       // barriers + TMEM slot (1 KB) | bias[npad] fp32 (1 KB multiple) | SE scratch (1 KB)
    // fused RGB first layer (conv3x3 4 -> 32 + LeakyReLU computed by four extra warps straight into the patch ring): see fusedFirstProducer
    int fused;
    const __half* fuseIn;       // NHWC4 fp16 input of the first layer, [gn][fuseH][fuseW][4]
    const __half* fuseW;        // first-layer weights [32][36], k = (ky*3+kx)*4 + c
    const float* fuseBias;      // [32]
    float fuseSlope;
    int fuseW_px, fuseH_px;
    long long fuseSn;           // elements between images of fuseIn
    uint32_t fuseOff;           // byte offset (from the aligned smem base) of the two RGB patch buffers
    int fuseImgBase;            // fused first layer reading a frame-wide tile buffer: index of this batch's first image in it
    long long* prof;            // W2X_DEV only: per-CTA cycle counters [grid][16] (W2X_PROF=1), see profSlot
};

struct IgemmPlan {
    ConvArgs args;
    int grid = 0;
    size_t smem = 0;
    bool patch = false;
    // fused first layer: a second RGB tensor map over the frame-wide unpacked-tile buffer (igemmSetFusedFrameInput)
    CUtensorMap tmRgbFrame;
    const __half* frameBase = nullptr;
    long long frameImages = 0;
};

namespace {

constexpr int kEpiWarps = 8;                       // two warps per TMEM lane quarter, splitting the columns
constexpr int kThreads = 64 + 32 * kEpiWarps;      // warp 0 = TMA producer, warp 1 = MMA issuer, warps 2.. = epilogue
constexpr int kEpiThreads = 32 * kEpiWarps;
constexpr int kPatchW = 10, kPatchH = 18;

// header layout (byte offsets from the 1024-aligned smem base)
constexpr uint32_t kOffFull = 0, kOffEmpty = 64, kOffTFull = 128, kOffTEmpty = 160, kOffSkip = 192, kOffW = 216, kOffRgbFull = 224, kOffRgbEmpty = 256,
                   kOffP1 = 288, kOffImFull = 312, kOffSlot = 336, kOffBias = 1024;
constexpr int kMaxAcc = 4;         // TMEM accumulator buffers (barriers for four; nAcc = 2 or 4 per plan)
constexpr int kRgbBufs = 4;        // fused first layer: RGB patches in flight
constexpr int kFuseWarps = 4;      // fused first layer: producer warps, one per TMEM lane quarter
constexpr int kFuseBufs = 3;       // fused first layer: im2col tiles / TMEM regions in flight.  The first-layer MMAs of tile k + 2 are queued
                                   // before the main MMAs of tile k: the tensor pipe's queue holds about one tile of main MMAs, so a lead of
                                   // one tile leaves the producer waiting for its accumulators (measured: 2080 instead of ~1330 cycles per tile)

enum EpiKind { EPI_K_DIRECT = 0, EPI_K_TMA = 1, EPI_K_TMA_SKIP = 2, EPI_K_STAGED = 3, EPI_K_TMA_GROUPS = 4,
               EPI_K_STAGED2 = 5 };  // STAGED2: the same code compiled for two CTAs per SM (register cap 102)

struct TileCoord {
    int img, y0, x0, n0;
};

// Walks this CTA's tile sequence t = first, first + step, ... without divisions in the loop.
struct TileWalker {
    int tn, tx, ty, img;          // current tile indices
    int sn, sx, sy, simg;         // step decomposition
    int tilesN, tilesX, tilesY;
    __device__ __forceinline__ void init(int first, int step, int tilesN_, int tilesX_, int tilesY_) {
        tilesN = tilesN_; tilesX = tilesX_; tilesY = tilesY_;
        int t = first;
        tn = t % tilesN; t /= tilesN;
        tx = t % tilesX; t /= tilesX;
        ty = t % tilesY; img = t / tilesY;
        t = step;
        sn = t % tilesN; t /= tilesN;
        sx = t % tilesX; t /= tilesX;
        sy = t % tilesY; simg = t / tilesY;
    }
    __device__ __forceinline__ void next() {
        tn += sn;
        int c = tn >= tilesN; tn -= c * tilesN;
        tx += sx + c;
        c = tx >= tilesX; tx -= c * tilesX;
        ty += sy + c;
        c = ty >= tilesY; ty -= c * tilesY;
        img += simg + c;
    }
    __device__ __forceinline__ TileCoord coord(int bh, int bw, int bn, int nBase) const {
        return TileCoord{img, ty * bh, tx * bw, nBase + tn * bn};
    }
};

// number of tiles CTA `first` (of `step`) processes (global round-robin)
__device__ __forceinline__ int tilesForCta(const ConvArgs& a, int first, int step) {
    return a.totalTiles > first ? (a.totalTiles - 1 - first) / step + 1 : 0;
}

// coordinates of 64-channel sub-tile `s` of an N tile in the output / skip tensor-map views
__device__ __forceinline__ void subTileCoords(const ConvArgs& a, const TileCoord& tc, int s, int& c0, int& cz) {
    const int j = tc.n0 + 64 * s;
    if (a.p.mode == EPI_D2S) {
        const int q = j / a.p.cout;
        c0 = (q & 1) * a.p.cout + (j - q * a.p.cout);  // (dx * C + channel) inside the merged [2][C] inner dimension
        cz = q >> 1;                                    // dy
    } else {
        c0 = j;
        cz = 0;
    }
}

// ---- epilogue shared by both kernels (warps 2..9 = 256 threads) --------------------------------------------------------
// Warp w may only touch TMEM lanes 32*(w%4)..+31; the two warps of a lane quarter split the accumulator columns.
template <int kEpi>
__device__ __forceinline__ void epilogueWarps(const ConvArgs& a, uint32_t base, uint32_t tmemBase, int nMine, int first, int step, int nBase) {
    constexpr bool kTma = kEpi == EPI_K_TMA || kEpi == EPI_K_TMA_SKIP;
    constexpr bool kSkip = kEpi == EPI_K_TMA_SKIP;
    constexpr bool kStaged = kEpi == EPI_K_STAGED || kEpi == EPI_K_STAGED2;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;
    const bool leader = threadIdx.x == 64;
    const int m = quarter * 32 + lane;
    const int yy = m >> a.bwShift, xx = m & (a.bw - 1);
    const uint32_t barTFull = base + kOffTFull, barTEmpty = base + kOffTEmpty, barSkip = base + kOffSkip;
    const uint32_t staging = base + a.headerBytes;
    const uint32_t bufBytes = (uint32_t)a.nsub * 16384u;
    const uint32_t biasS = base + kOffBias;
    const int skipHalf = a.p.skip_off >> 1;
    const bool split = a.bn >= 64 && (a.bn & 31) == 0;       // two warps per lane quarter split the columns
    const int colsPerWarp = split ? a.bn >> 1 : a.bn;        // small heads: only the first warp of each quarter works
    const int colBegin = split ? half * colsPerWarp : 0;
    const bool active = split || half == 0;

    TileWalker w;
    w.init(first, step, a.tilesN, a.tilesX, a.tilesY);
    TileWalker ws = w;  // lookahead walker for skip-tile prefetch (leader only)
    int ksNext = 0;
    auto issueSkip = [&](int k) {
        // ws points at tile k
        const TileCoord tc = ws.coord(a.bh, a.bw, a.bn, nBase);
        const int b = k % a.nbuf;
        mbarExpectTx(barSkip + 8u * b, bufBytes);
        for (int s = 0; s < a.nsub; ++s) {
            int c0, cz;
            subTileCoords(a, tc, s, c0, cz);
            tmaLoad5d(staging + b * bufBytes + s * 16384u, &a.tmSkip, barSkip + 8u * b, c0, tc.x0 + skipHalf, cz, tc.y0 + skipHalf, tc.img);
        }
        ws.next();
        ksNext = k + 1;
    };
    if (kSkip && leader) {
        for (int k = 0; k < a.nbuf - 1 && k < nMine; ++k) issueSkip(k);
    }

    // fused SE squeeze: while a warp converts its 32 rows x 32 columns of a tile, every column is summed over the 32 rows with one
    // redux.sync on the 2^10 fixed-point value (|v| <= 65504 cannot overflow the 32-row int32 sum); lane j keeps column j's total
    // in a 64-bit register and issues one integer atomic per (thread, column chunk, image).  Integer addition is exact, so the
    // result is independent of the tile -> CTA assignment and of the image's slot in the batch (byte-identical under any sharding).
    const bool doSe = kTma && !kSkip && a.p.se_sum != nullptr;
    const int et = threadIdx.x - 64;
    long long seAcc0 = 0, seAcc1 = 0;
    int seImg = -1;
    int skipScaleImg = -1;
    auto seFlush = [&]() {
        if (seImg < 0) return;
        unsigned long long* dst = reinterpret_cast<unsigned long long*>(a.p.se_sum + (size_t)seImg * a.p.npad + nBase + colBegin + lane);
        if (seAcc0 != 0) atomicAdd(dst, (unsigned long long)seAcc0);
        if (seAcc1 != 0) atomicAdd(dst + 32, (unsigned long long)seAcc1);
    };

    int acc = 0, b = 0;
    uint32_t accPhase = 0, bufPhase = 0;
    for (int k = 0; k < nMine; ++k, w.next()) {
        const TileCoord tc = w.coord(a.bh, a.bw, a.bn, nBase);
        const uint32_t stg = staging + b * bufBytes;
        if (kTma && !kSkip) {
            if (leader) bulkWaitRead(a.nbuf - 1);  // staging buffer b is no longer being read by the store of tile k - nbuf
            namedBarSync(1, kEpiThreads);
        }
        // image head: fetch this pixel's skip value (z1 crop) BEFORE blocking on the accumulator, so its latency is hidden
        Half4 preSkip{};
        const int py = tc.y0 + yy, px = tc.x0 + xx;
        const bool pvalid = py < a.p.gy && px < a.p.gx;
        if (!kTma && a.p.mode == EPI_FINAL && active && pvalid)
            preSkip = *reinterpret_cast<const Half4*>(a.p.skip + (((long long)tc.img * a.p.skip_h + py + a.p.skip_off) * a.p.skip_w + px + a.p.skip_off) * a.p.skip_c);
        if (kSkip && a.p.skip_scale && tc.img != skipScaleImg) {
            // per-image SE scale of the skip tensor -> smem scratch (image changes are uniform across the epilogue threads)
            namedBarSync(2, kEpiThreads);
            float* scratch = reinterpret_cast<float*>(__cvta_shared_to_generic((size_t)(base + a.headerBytes - 1024u)));
            if (et < a.p.skip_c) scratch[et] = a.p.skip_scale[(long long)tc.img * a.p.skip_c + et];
            namedBarSync(2, kEpiThreads);
            skipScaleImg = tc.img;
        }
        // token-wise Linear with a residual: the first round of residual chunks is fetched BEFORE blocking on the accumulator, so that
        // its global-memory latency hides behind the MMAs (the copy-out below was three dependent load round trips per tile)
        constexpr int kPre = 4;
        uint4 preRes[kPre];
        const bool prefetched = kStaged && a.p.skip && a.p.mode == EPI_STORE;  // (the pixel-shuffle layers address their residual differently)
        if (prefetched) {
            const int chunksPerRow = a.bn >> 3, total = 128 * chunksPerRow;
#pragma unroll
            for (int i = 0; i < kPre; ++i) {
                preRes[i] = make_uint4(0, 0, 0, 0);
                const int idx = et + i * kEpiThreads;
                if (idx < total) {
                    const int row = idx / chunksPerRow, cc = idx - row * chunksPerRow;
                    const int yy2 = tc.y0 + (row >> a.bwShift), xx2 = tc.x0 + (row & (a.bw - 1));
                    if (yy2 < a.p.gy && xx2 < a.p.gx)
                        preRes[i] = *reinterpret_cast<const uint4*>(a.p.skip + (((long long)tc.img * a.p.out_h + yy2) * a.p.out_w + xx2) * a.p.out_c + tc.n0 + cc * 8);
                }
            }
        }
        mbarWait(barTFull + 8u * acc, accPhase);
        tcFenceAfter();
        if (kSkip) mbarWait(barSkip + 8u * b, bufPhase);
        const uint32_t taddr = tmemBase + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * a.bn);
        uint32_t r[32];
        if (kTma) {
            if (doSe && tc.img != seImg) { seFlush(); seImg = tc.img; seAcc0 = seAcc1 = 0; }
            const float seScale = pvalid ? kSeFixedScale : 0.f;  // rows outside the layer's output do not count
            for (int c0 = colBegin; c0 < colBegin + colsPerWarp; c0 += 32) {
                tmemLd32(taddr + (uint32_t)c0, r);
                tmemLdWait();
                int seMine = 0;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int j0 = c0 + 8 * q;
                    const uint32_t addr = stg + (uint32_t)(j0 >> 6) * 16384u + (uint32_t)m * 128u + ((uint32_t)(((j0 & 63) >> 3) ^ (m & 7)) << 4);
                    const uint4 b0 = ldsV4(biasS + (uint32_t)(tc.n0 + j0) * 4u), b1 = ldsV4(biasS + (uint32_t)(tc.n0 + j0 + 4) * 4u);
                    const float bias[8] = {__uint_as_float(b0.x), __uint_as_float(b0.y), __uint_as_float(b0.z), __uint_as_float(b0.w),
                                           __uint_as_float(b1.x), __uint_as_float(b1.y), __uint_as_float(b1.z), __uint_as_float(b1.w)};
                    float v[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float t = __uint_as_float(r[8 * q + i]) + bias[i];
                        v[i] = fmaxf(t, t * a.p.slope);  // LeakyReLU for 0 < slope <= 1
                    }
                    if (kSkip) {
                        const uint4 sv = ldsV4(addr);
                        const __half2* sh = reinterpret_cast<const __half2*>(&sv);
                        if (a.p.skip_scale) {  // the skip tensor's SE scale, applied here instead of a separate in-place pass
                            const uint32_t co = (uint32_t)((tc.n0 + j0) % a.p.cout);
                            const uint4 s0 = ldsV4(base + a.headerBytes - 1024u + co * 4u), s1 = ldsV4(base + a.headerBytes - 1024u + co * 4u + 16u);
                            const float ss[8] = {__uint_as_float(s0.x), __uint_as_float(s0.y), __uint_as_float(s0.z), __uint_as_float(s0.w),
                                                 __uint_as_float(s1.x), __uint_as_float(s1.y), __uint_as_float(s1.z), __uint_as_float(s1.w)};
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const float2 f = __half22float2(sh[i]);
                                v[2 * i] = fmaf(f.x, ss[2 * i], v[2 * i]);
                                v[2 * i + 1] = fmaf(f.y, ss[2 * i + 1], v[2 * i + 1]);
                            }
                        } else {
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const float2 f = __half22float2(sh[i]);
                                v[2 * i] += f.x;
                                v[2 * i + 1] += f.y;
                            }
                        }
                    }
                    uint4 o;
                    __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
                    for (int i = 0; i < 4; ++i) oh[i] = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
                    stsV4(addr, o);
                    if (!kSkip && doSe) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int colSum = __reduce_add_sync(0xffffffffu, __float2int_rn(v[i] * seScale));
                            if (lane == 8 * q + i) seMine = colSum;
                        }
                    }
                }
                if (!kSkip && doSe) {
                    if (c0 == colBegin) seAcc0 += seMine;
                    else seAcc1 += seMine;
                }
            }
        } else if (kStaged) {
            // token-wise Linear layers (any N % 32 == 0): act(acc + bias) -> padded smem rows -> coalesced 16-byte copy-out
            // (+ residual read at the same address, so the residual stream can be updated in place)
            const uint32_t sbuf = staging + (uint32_t)(k & (a.nbuf - 1)) * a.stagedBuf;
            int c0 = colBegin;
            while (c0 < colBegin + colsPerWarp) {
                const bool wide = c0 + 32 <= colBegin + colsPerWarp;
                if (wide) tmemLd32(taddr + (uint32_t)c0, r);
                else tmemLd16(taddr + (uint32_t)c0, r);
                tmemLdWait();
                const int nq = wide ? 4 : 2;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (q < nq) {
                        const int j0 = c0 + 8 * q;
                        const uint4 b0 = ldsV4(biasS + (uint32_t)(tc.n0 + j0) * 4u), b1 = ldsV4(biasS + (uint32_t)(tc.n0 + j0 + 4) * 4u);
                        const float bias[8] = {__uint_as_float(b0.x), __uint_as_float(b0.y), __uint_as_float(b0.z), __uint_as_float(b0.w),
                                               __uint_as_float(b1.x), __uint_as_float(b1.y), __uint_as_float(b1.z), __uint_as_float(b1.w)};
                        float v[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const float tv = __uint_as_float(r[8 * q + i]) + bias[i];
                            v[i] = a.p.act == ACT_GELU ? geluErf(tv) : fmaxf(tv, tv * a.p.slope);
                        }
                        uint4 o;
                        __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
                        for (int i = 0; i < 4; ++i) oh[i] = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
                        stsV4(sbuf + (uint32_t)m * a.stagedPitch + (uint32_t)j0 * 2u, o);
                    }
                }
                c0 += wide ? 32 : 16;
            }
        } else if (active) {
            const int y = py, x = px;
            const bool valid = pvalid;
            const Half4* pre = a.p.mode == EPI_FINAL ? &preSkip : nullptr;
            if (a.bn >= 32) {
                int c0 = colBegin;
                for (; c0 + 32 <= colBegin + colsPerWarp; c0 += 32) {
                    tmemLd32(taddr + (uint32_t)c0, r);
                    tmemLdWait();
                    if (valid) {
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            conv_epilogue8(a.p, tc.img, y, x, tc.n0 + c0 + 8 * q, reinterpret_cast<const float*>(r) + 8 * q);
                    }
                }
                if (c0 < colBegin + colsPerWarp) {  // 16-column tail (bn = 96: 48 columns per warp)
                    tmemLd16(taddr + (uint32_t)c0, r);
                    tmemLdWait();
                    if (valid) {
#pragma unroll
                        for (int q = 0; q < 2; ++q)
                            conv_epilogue8(a.p, tc.img, y, x, tc.n0 + c0 + 8 * q, reinterpret_cast<const float*>(r) + 8 * q);
                    }
                }
            } else {
                tmemLd16(taddr, r);
                tmemLdWait();
                if (valid) {
#pragma unroll
                    for (int q = 0; q < 2; ++q) conv_epilogue8(a.p, tc.img, y, x, tc.n0 + 8 * q, reinterpret_cast<const float*>(r) + 8 * q, pre);
                }
            }
        }
        tcFenceBefore();
        __syncwarp();
        if (lane == 0) mbarArrive(barTEmpty + 8u * acc);  // accumulator buffer may be overwritten by the next-but-one tile
        if (++acc == 2) { acc = 0; accPhase ^= 1u; }
        if (++b == a.nbuf) { b = 0; bufPhase ^= 1u; }
        if (kStaged) {
            namedBarSync(1, kEpiThreads);  // the tile is complete in smem (and, two tiles later, this buffer is free again)
            const uint32_t sbuf = staging + (uint32_t)(k & (a.nbuf - 1)) * a.stagedBuf;
            const int chunksPerRow = a.bn >> 3;
            const int et2 = threadIdx.x - 64;
            const int total = 128 * chunksPerRow;
            auto chunkOffset = [&](int row, int cc, bool& ok) -> long long {
                const int yy2 = tc.y0 + (row >> a.bwShift), xx2 = tc.x0 + (row & (a.bw - 1));
                ok = yy2 < a.p.gy && xx2 < a.p.gx;
                if (a.p.mode == EPI_D2S) {  // pixel shuffle: column (q*cout + co) -> output pixel (2y + q/2, 2x + q%2), channel co
                    const int j = tc.n0 + cc * 8;
                    const int q = j / a.p.cout, co = j - q * a.p.cout;
                    return (((long long)tc.img * a.p.out_h + 2 * yy2 + (q >> 1)) * a.p.out_w + 2 * xx2 + (q & 1)) * a.p.out_c + co;
                }
                return (((long long)tc.img * a.p.out_h + yy2) * a.p.out_w + xx2) * a.p.out_c + tc.n0 + cc * 8;
            };
            if (a.p.skip) {
                // residual layers (same geometry as the output: the residual stream is updated in place): kPre chunks per thread per
                // round, all residual loads of a round in flight before any is consumed; round 0 was prefetched above
                for (int base0 = et2, round = 0; base0 < total; base0 += kPre * kEpiThreads, ++round) {
                    uint4 res[kPre];
                    long long offs[kPre];
                    bool oks[kPre];
#pragma unroll
                    for (int i = 0; i < kPre; ++i) {
                        const int idx = base0 + i * kEpiThreads;
                        oks[i] = false;
                        offs[i] = 0;
                        res[i] = make_uint4(0, 0, 0, 0);
                        if (idx < total) {
                            const int row = idx / chunksPerRow, cc = idx - row * chunksPerRow;
                            offs[i] = chunkOffset(row, cc, oks[i]);
                            if (round == 0 && prefetched) res[i] = preRes[i];
                            else if (oks[i]) res[i] = *reinterpret_cast<const uint4*>(a.p.skip + offs[i]);
                        }
                    }
#pragma unroll
                    for (int i = 0; i < kPre; ++i) {
                        const int idx = base0 + i * kEpiThreads;
                        if (idx >= total || !oks[i]) continue;
                        const int row = idx / chunksPerRow, cc = idx - row * chunksPerRow;
                        uint4 v = ldsV4(sbuf + (uint32_t)row * a.stagedPitch + (uint32_t)cc * 16u);
                        __half2* h = reinterpret_cast<__half2*>(&v);
                        const __half2* r = reinterpret_cast<const __half2*>(&res[i]);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float2 x0 = __half22float2(h[j]), x1 = __half22float2(r[j]);
                            h[j] = __floats2half2_rn(x0.x + x1.x, x0.y + x1.y);
                        }
                        *reinterpret_cast<uint4*>(a.p.out + offs[i]) = v;
                    }
                }
            } else {
                for (int idx = et2; idx < total; idx += kEpiThreads) {
                    const int row = idx / chunksPerRow, cc = idx - row * chunksPerRow;
                    bool ok;
                    const long long off = chunkOffset(row, cc, ok);
                    if (ok) *reinterpret_cast<uint4*>(a.p.out + off) = ldsV4(sbuf + (uint32_t)row * a.stagedPitch + (uint32_t)cc * 16u);
                }
            }
        }
        if (kTma) {
            fenceProxyAsync();
            namedBarSync(1, kEpiThreads);
            if (leader) {
                for (int s = 0; s < a.nsub; ++s) {
                    int c0, cz;
                    subTileCoords(a, tc, s, c0, cz);
                    tmaStore5d(&a.tmOut, stg + s * 16384u, c0, tc.x0, cz, tc.y0, tc.img);
                }
                bulkCommit();
            }
            if (leader) {
                if (kSkip) {
                    const int kn = k + a.nbuf - 1;
                    if (kn < nMine) {
                        bulkWaitRead(1);  // the store of tile k-1 (same buffer as tile kn) has finished reading smem
                        issueSkip(kn);
                    }
                }
            }
        }
    }
    if (doSe) seFlush();
    if (kTma && leader) bulkWaitAll();
}

// ---- TMA-store epilogue in two independent groups of four warps ------------------------------------------------------------
// Group g (one warp per TMEM lane quarter) owns accumulator buffer g, staging buffer g and every second tile of the CTA, so one
// group's barrier / TMEM / store latencies overlap the other group's math, and the per-tile bookkeeping (tile walker, barriers)
// is paid once per 64 accumulator columns of a thread instead of once per 32.  Requires nbuf == 2, no skip tensor.
__device__ __forceinline__ void epilogueTmaGroups(const ConvArgs& a, uint32_t base, uint32_t tmemBase, int nMine, int first, int step, int nBase) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int quarter = warp & 3;
    const int group = (warp - 2) >> 2;
    const bool leader = ((warp - 2) & 3) == 0 && lane == 0;
    const int m = quarter * 32 + lane;
    const int yy = m >> a.bwShift, xx = m & (a.bw - 1);
    const uint32_t stg = base + a.headerBytes + (uint32_t)group * (uint32_t)a.nsub * 16384u;
    const uint32_t rowAddr = stg + (uint32_t)m * 128u;
    const uint32_t sw = (uint32_t)(m & 7);
    const uint32_t biasS = base + kOffBias;
    const uint32_t taddrLane = tmemBase + ((uint32_t)(quarter * 32) << 16);
    const bool fourAcc = a.nAcc == 4;  // tile k of the CTA uses accumulator k % nAcc: this group's j-th tile (k = 2j + group) -> group + 2 * (j & 1)
    const int nPairs = a.bn >> 6;  // 64 accumulator columns per round
    const int barId = 1 + group;

    // fused SE squeeze (see epilogueWarps): redux.sync per column over the warp's 32 rows, 64-bit totals per lane, one integer
    // atomic per (thread, 32-column chunk, image)
    const bool doSe = a.p.se_sum != nullptr;
    long long seAcc[4] = {0, 0, 0, 0};
    int seImg = -1;
    auto seFlush = [&]() {
        if (seImg < 0) return;
        unsigned long long* dst = reinterpret_cast<unsigned long long*>(a.p.se_sum + (size_t)seImg * a.p.npad + nBase + lane);
#pragma unroll
        for (int c = 0; c < 4; ++c)
            if (seAcc[c] != 0) atomicAdd(dst + 32 * c, (unsigned long long)seAcc[c]);
    };

    const int nGroup = nMine > group ? (nMine - group + 1) >> 1 : 0;  // tiles group, group + 2, ... of this CTA
    TileWalker w;
    w.init(first + group * step, 2 * step, a.tilesN, a.tilesX, a.tilesY);
    W2X_PROF_DECL(pWaitFull); W2X_PROF_DECL(pMath); W2X_PROF_DECL(pSync);
    W2X_PROF_T(pT0);
    for (int j = 0; j < nGroup; ++j, w.next()) {
        const TileCoord tc = w.coord(a.bh, a.bw, a.bn, nBase);
        const int accIdx = fourAcc ? group + 2 * (j & 1) : group;
        const uint32_t phase = (uint32_t)(fourAcc ? (j >> 1) : j) & 1u;
        const uint32_t barTFull = base + kOffTFull + 8u * accIdx, barTEmpty = base + kOffTEmpty + 8u * accIdx;
        const uint32_t taddr = taddrLane + (uint32_t)(accIdx * a.bn);
        W2X_PROF_T(ps0);
        if (leader) bulkWaitRead(0);  // this group's previous store (two tiles ago) has finished reading the staging buffer
        namedBarSync(barId, 128);
        W2X_PROF_ADD(pSync, ps0);
        float seScale = 0.f;
        if (doSe) {
            if (tc.img != seImg) {
                seFlush();
                seImg = tc.img;
#pragma unroll
                for (int c = 0; c < 4; ++c) seAcc[c] = 0;
            }
            const bool pvalid = (tc.y0 + yy) < a.p.gy && (tc.x0 + xx) < a.p.gx;  // rows outside the layer's output do not count
            seScale = pvalid ? kSeFixedScale : 0.f;
        }
        W2X_PROF_T(pw0);
        mbarWait(barTFull, phase);
        tcFenceAfter();
        if (leader) W2X_TRACE(a, 2 * j + group, 6);   // accumulator of tile 2j + group complete
        W2X_PROF_ADD(pWaitFull, pw0);
        W2X_PROF_T(pm0);
#pragma unroll
        for (int pr = 0; pr < 2; ++pr) {
            if (pr < nPairs && !W2X_DBG_ON(a, 1)) {
                uint32_t rLo[32], rHi[32];
                tmemLd32(taddr + (uint32_t)(64 * pr), rLo);
                tmemLd32(taddr + (uint32_t)(64 * pr + 32), rHi);
                tmemLdWait();
                const uint32_t subAddr = rowAddr + (uint32_t)pr * 16384u;
                const uint32_t biasAddr = biasS + (uint32_t)(tc.n0 + 64 * pr) * 4u;
                int seMine0 = 0, seMine1 = 0;
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const uint4 b0 = ldsV4(biasAddr + 32u * q), b1 = ldsV4(biasAddr + 32u * q + 16u);
                    const float bias[8] = {__uint_as_float(b0.x), __uint_as_float(b0.y), __uint_as_float(b0.z), __uint_as_float(b0.w),
                                           __uint_as_float(b1.x), __uint_as_float(b1.y), __uint_as_float(b1.z), __uint_as_float(b1.w)};
                    float v[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float t = __uint_as_float(q < 4 ? rLo[8 * (q & 3) + i] : rHi[8 * (q & 3) + i]) + bias[i];
                        v[i] = fmaxf(t, t * a.p.slope);  // LeakyReLU for 0 < slope <= 1
                    }
                    uint4 o;
                    __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
                    for (int i = 0; i < 4; ++i) oh[i] = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
                    stsV4(subAddr + (((uint32_t)q ^ sw) << 4), o);
                    if (doSe) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int colSum = __reduce_add_sync(0xffffffffu, __float2int_rn(v[i] * seScale));
                            if (q < 4) { if (lane == 8 * q + i) seMine0 = colSum; }
                            else { if (lane == 8 * (q - 4) + i) seMine1 = colSum; }
                        }
                    }
                }
                if (doSe) {
                    seAcc[2 * pr] += seMine0;
                    seAcc[2 * pr + 1] += seMine1;
                }
            }
        }
        tcFenceBefore();
        __syncwarp();
        if (lane == 0) mbarArrive(barTEmpty);  // accumulator buffer `group` may be overwritten by this group's next tile
        W2X_PROF_ADD(pMath, pm0);
        W2X_PROF_T(ps1);
        fenceProxyAsync();
        namedBarSync(barId, 128);
        if (leader && !W2X_DBG_ON(a, 2)) {
            for (int sub = 0; sub < a.nsub; ++sub) {
                int c0, cz;
                subTileCoords(a, tc, sub, c0, cz);
                tmaStore5d(&a.tmOut, stg + sub * 16384u, c0, tc.x0, cz, tc.y0, tc.img);
            }
            bulkCommit();
        }
        W2X_PROF_ADD(pSync, ps1);
    }
#ifdef W2X_DEV
    if (a.prof && group == 0 && leader) {
        long long* pr = a.prof + 16ll * blockIdx.x;
        pr[6] = clock64() - pT0; pr[7] = pWaitFull; pr[8] = pMath; pr[9] = pSync;
    }
#endif
    if (doSe) seFlush();
    if (leader) bulkWaitAll();
}

// ---- fused RGB first layer ---------------------------------------------------------------------------------------------
// The 32-channel input of the second convolution of a UNet (conv1.conv.2) is itself conv3x3(RGB) + LeakyReLU: 27 MACs per value.
// Instead of a separate kernel writing that tensor to HBM and TMA reading it back, four extra warps compute each tile's 18x10x32
// patch on the tensor cores as well: they gather a [180 (-> 2 x 128) patch pixels] x [K = 27 -> 32] im2col matrix from the 20x12
// RGB patch (TMA-loaded, four deep) into shared memory in the SWIZZLE_64B operand layout, one elected thread issues four
// tcgen05.mma (M = 128, N = 32, K = 16) into a private TMEM region, and the same warps read the accumulators back (tcgen05.ld),
// add the bias, apply LeakyReLU and store fp16 into the patch ring slot in exactly the layout a SWIZZLE_64B TMA box would have
// produced, then publish the slot on the full barrier the main MMA warp waits on (generic-proxy writes + fence.proxy.async).
// Everything is double buffered (im2col tile, TMEM region) and software pipelined: the im2col + MMAs of tile k+1 are issued before
// tile k is converted.  (The first version of this producer used mma.sync: 1475 cycles per tile, latency-bound and sharing the
// tensor pipe with the main MMAs; this one costs ~4 x 43 pipe cycles per tile.)  Saves, per output pixel of the first layer, 64 B
// written + 64 B (x halo) read of HBM traffic and one kernel launch.
constexpr uint32_t kIm2colBytes = 2 * 128 * 64;   // two M = 128 blocks of 64-byte rows
constexpr uint32_t kFuseExtraBytes = kRgbBufs * 2048u + kFuseBufs * kIm2colBytes + 2048u + 1024u;  // RGB ring | im2col tiles | W1 | bias1 (+ pad)
constexpr int kFuseN = 32;         // first-layer output channels = N of its MMAs

// half `i` (0..26) of a patch pixel's im2col row lives in source word (ky, 2 * kx + (c >> 1)), 16-bit half (c & 1)
__device__ __forceinline__ constexpr int im2colWord(int i) { return (i / 9) * 6 + 2 * ((i % 9) / 3) + (((i % 9) % 3) >> 1); }
__device__ __forceinline__ constexpr int im2colHalf(int i) { return ((i % 9) % 3) & 1; }

__device__ __forceinline__ void fusedFirstProducer(const ConvArgs& a, uint32_t base, uint32_t stage0, uint32_t tmemBase, int nMine) {
    constexpr int kRgbW = kPatchW + 2;                       // 12 x 20 input pixels per tile
    constexpr int kRows = kPatchW * kPatchH;                 // 180 patch pixels
    constexpr int kProdThreads = 32 * kFuseWarps;
    const int tid = threadIdx.x - kThreads;                  // 0 .. 127
    const int lane = tid & 31;
    const uint32_t barFull = base + kOffFull, barEmpty = base + kOffEmpty;
    const uint32_t barRgbFull = base + kOffRgbFull, barRgbEmpty = base + kOffRgbEmpty, barP1 = base + kOffP1;
    const uint32_t rgb0 = base + a.fuseOff;
    const uint32_t im0 = rgb0 + kRgbBufs * 2048u;
    const uint32_t w1s = im0 + kFuseBufs * kIm2colBytes;
    const uint32_t bias1s = w1s + 2048u;
    const uint32_t tmemP = tmemBase + (uint32_t)(a.nAcc * a.bn);   // kFuseBufs regions of 2 x 32 columns behind the main accumulators

    // ---- constants of the loaded model: W1 as the B operand [32 rows n][32 halfs k'] (k' = tap * 3 + c, zero beyond 27), bias
    {
        const int n = tid >> 2, q = tid & 3;                 // one 16-byte chunk (8 halfs) per thread
        const __half* wrow = a.fuseW + n * 36;               // packed [32][tap * 4 + c]
        uint32_t wv[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int k0 = 8 * q + 2 * j, k1 = k0 + 1;
            const uint32_t lo = k0 < 27 ? (uint32_t)__half_as_ushort(wrow[(k0 / 3) * 4 + k0 % 3]) : 0u;
            const uint32_t hi = k1 < 27 ? (uint32_t)__half_as_ushort(wrow[(k1 / 3) * 4 + k1 % 3]) : 0u;
            wv[j] = lo | (hi << 16);
        }
        stsV4(w1s + (uint32_t)n * 64u + ((uint32_t)(q ^ ((n >> 1) & 3)) << 4), make_uint4(wv[0], wv[1], wv[2], wv[3]));
        if (tid < 32) {
            const float bv = a.fuseBias[tid];
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(bias1s + 4u * tid), "f"(bv) : "memory");
        }
        fenceProxyAsync();
    }
    pdlWait();  // weights / bias above are constants; the RGB tiles come from the preceding kernel

    const int quarter = (threadIdx.x >> 5) & 3;              // TMEM lane quarter this warp may read: its index in the CTA modulo 4

    // im2col of tile k into buffer k % kFuseBufs (the main MMA warp issues its MMAs: fusedFirstIssue)
    auto stageTile = [&](int k) {
        const int rb = k % kRgbBufs;
        mbarWait(barRgbFull + 8u * rb, (uint32_t)(k / kRgbBufs) & 1u);
        const uint32_t rgb = rgb0 + (uint32_t)rb * 2048u;
        const uint32_t im = im0 + (uint32_t)(k % kFuseBufs) * kIm2colBytes;
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
            const int prow = tid + rr * kProdThreads;
            if (prow < kRows) {
                const int py = prow / kPatchW, px = prow - py * kPatchW;
                uint32_t w[18];
#pragma unroll
                for (int ky = 0; ky < 3; ++ky) {
                    const uint32_t src = rgb + (uint32_t)(((py + ky) * kRgbW + px) * 8);
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx)
                        asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(w[ky * 6 + 2 * kx]), "=r"(w[ky * 6 + 2 * kx + 1]) : "r"(src + 8u * kx));
                }
                uint32_t o[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int i0 = 2 * j, i1 = 2 * j + 1;
                    if (i0 >= 27) { o[j] = 0u; continue; }
                    const uint32_t wa = w[im2colWord(i0)];
                    const uint32_t wb = i1 < 27 ? w[im2colWord(i1 < 27 ? i1 : 0)] : 0u;
                    const uint32_t sel = (im2colHalf(i0) ? 0x32u : 0x10u) | ((i1 < 27 ? (im2colHalf(i1 < 27 ? i1 : 0) ? 0x76u : 0x54u) : 0x54u) << 8);
                    o[j] = __byte_perm(wa, wb, sel);
                }
                const uint32_t rowAddr = im + (uint32_t)prow * 64u;
                const uint32_t sw = (uint32_t)(prow >> 1) & 3u;
#pragma unroll
                for (int q = 0; q < 4; ++q) stsV4(rowAddr + (((uint32_t)q ^ sw) << 4), make_uint4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]));
            }
        }
        fenceProxyAsync();   // generic-proxy stores -> visible to the tensor core's async-proxy operand reads
        tcFenceBefore();     // this thread's earlier tcgen05.ld of the TMEM buffer that the MMAs of this tile will overwrite
        __syncwarp();
        if (lane == 0) {
            mbarArrive(barRgbEmpty + 8u * rb);                 // this warp has read everything it needs from the RGB patch
            mbarArrive(base + kOffImFull + 8u * (k % kFuseBufs));  // the main MMA warp issues this tile's four first-layer MMAs
        }
    };

    int stage = 0;
    uint32_t phase = 0;
    W2X_PROF_DECL(pStage); W2X_PROF_DECL(pWaitP1); W2X_PROF_DECL(pWaitSlot); W2X_PROF_DECL(pConv);
    W2X_PROF_T(pT0);
    for (int k = 0; k < kFuseBufs - 1 && k < nMine; ++k) stageTile(k);
    for (int k = 0; k < nMine; ++k) {
        W2X_PROF_T(pq0);
        if (k + kFuseBufs - 1 < nMine) stageTile(k + kFuseBufs - 1);   // its buffer was tile k - 1's: those MMAs completed before convert(k - 1)
        W2X_PROF_ADD(pStage, pq0);
        const int fb = k % kFuseBufs;
        const uint32_t fphase = (uint32_t)(k / kFuseBufs) & 1u;
        // ---- convert tile k: TMEM -> bias + LeakyReLU -> fp16 -> ring slot (SWIZZLE_64B placement: chunk c of pixel row p at c ^ ((p >> 1) & 3))
        W2X_PROF_T(pq1);
        if (tid == 0) W2X_TRACE(a, k, 3);       // producer starts waiting for the first-layer accumulators of tile k
        mbarWait(barP1 + 8u * fb, fphase);
        tcFenceAfter();
        if (tid == 0) W2X_TRACE(a, k, 4);       // ... has them
        W2X_PROF_ADD(pWaitP1, pq1);
        W2X_PROF_T(pq2);
        mbarWait(barEmpty + 8u * stage, phase ^ 1u);  // the MMAs that read this ring slot have completed
        W2X_PROF_ADD(pWaitSlot, pq2);
        W2X_PROF_T(pq3);
        const uint32_t dst = stage0 + (uint32_t)stage * a.stageStride;
#pragma unroll
        for (int mb = 0; mb < 2; ++mb) {
            if (mb == 1 && quarter >= 2) break;       // rows 192.. do not exist (180 patch pixels): the whole warp skips
            const int prow = mb * 128 + quarter * 32 + lane;
            uint32_t r[32];
            tmemLd32(tmemP + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(fb * 2 * kFuseN + mb * kFuseN), r);
            tmemLdWait();
            if (prow < kRows) {
                const uint32_t rowAddr = dst + (uint32_t)prow * 64u;
                const uint32_t sw = (uint32_t)(prow >> 1) & 3u;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const uint4 b0 = ldsV4(bias1s + 32u * q), b1 = ldsV4(bias1s + 32u * q + 16u);
                    const float bias[8] = {__uint_as_float(b0.x), __uint_as_float(b0.y), __uint_as_float(b0.z), __uint_as_float(b0.w),
                                           __uint_as_float(b1.x), __uint_as_float(b1.y), __uint_as_float(b1.z), __uint_as_float(b1.w)};
                    uint4 ov;
                    __half2* oh = reinterpret_cast<__half2*>(&ov);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float t0 = __uint_as_float(r[8 * q + 2 * i]) + bias[2 * i], t1 = __uint_as_float(r[8 * q + 2 * i + 1]) + bias[2 * i + 1];
                        oh[i] = __floats2half2_rn(fmaxf(t0, t0 * a.fuseSlope), fmaxf(t1, t1 * a.fuseSlope));
                    }
                    stsV4(rowAddr + (((uint32_t)q ^ sw) << 4), ov);
                }
            }
        }
        tcFenceBefore();
        fenceProxyAsync();  // generic-proxy stores above -> visible to the tensor core's async-proxy reads
        __syncwarp();
        if (lane == 0) mbarArrive(barFull + 8u * stage);
        if (tid == 0) W2X_TRACE(a, k, 5);       // patch of tile k published
        W2X_PROF_ADD(pConv, pq3);
        if (++stage == a.stages) { stage = 0; phase ^= 1u; }
    }
#ifdef W2X_DEV
    if (a.prof && tid == 0) {
        long long* pr = a.prof + 16ll * blockIdx.x;
        pr[12] = clock64() - pT0; pr[13] = pStage; pr[14] = pWaitP1; pr[15] = pConv; pr[5] = pWaitSlot;
    }
#endif
}

__device__ __forceinline__ void setupCommon(const ConvArgs& a, uint32_t base, uint8_t* sm, int warp, int accReaders, int fullArrivals = 1) {
    if (threadIdx.x == 0) {
        for (int s = 0; s < 8; ++s) {
            mbarInit(base + kOffFull + 8u * s, fullArrivals);
            mbarInit(base + kOffEmpty + 8u * s, 1);
        }
        for (int i = 0; i < kMaxAcc; ++i) {
            mbarInit(base + kOffTFull + 8u * i, 1);
            mbarInit(base + kOffTEmpty + 8u * i, accReaders);  // warps that must release an accumulator buffer
        }
        for (int i = 0; i < 3; ++i) mbarInit(base + kOffSkip + 8u * i, 1);
        mbarInit(base + kOffW, 1);
        for (int i = 0; i < kRgbBufs; ++i) {
            mbarInit(base + kOffRgbFull + 8u * i, 1);
            mbarInit(base + kOffRgbEmpty + 8u * i, kFuseWarps);  // the producer warps of the fused first layer
        }
        for (int i = 0; i < kFuseBufs; ++i) {
            mbarInit(base + kOffP1 + 8u * i, 1);              // fused first layer: its MMAs into TMEM buffer i have completed
            mbarInit(base + kOffImFull + 8u * i, kFuseWarps);  // fused first layer: im2col tile i is written (one arrival per producer warp)
        }
        mbarInitFence();
        tmaPrefetchDesc(&a.tmA);
        tmaPrefetchDesc(&a.tmB);
        if (a.useTma) tmaPrefetchDesc(&a.tmOut);
        if (a.hasSkip) tmaPrefetchDesc(&a.tmSkip);
    }
    float* biasS = reinterpret_cast<float*>(sm + kOffBias);
    for (int i = threadIdx.x; i < a.p.npad && i < 1024; i += blockDim.x) biasS[i] = a.p.bias[i];
    if (warp == 1) tmemAlloc(smemU32(sm + kOffSlot), a.tmemCols);
    tcFenceBefore();
    __syncthreads();
    tcFenceAfter();
}

// ======================================================================================================================
// generic implicit GEMM (per-tap A loads)
// ======================================================================================================================
// Token-wise layers (staged epilogue: K <= 768, a few MMAs per tile) are latency-bound, not tensor-bound: the EPI_K_STAGED2 variant
// runs two CTAs per SM (half the shared memory and TMEM each), so one CTA's epilogue / copy-out overlaps the other's loads and MMAs.
template <int kEpi>
__global__ void __launch_bounds__(kThreads, kEpi == EPI_K_STAGED2 ? 2 : 1) igemm_kernel(const __grid_constant__ ConvArgs a) {
    extern __shared__ uint8_t smemRaw[];
    const uint32_t rawAddr = smemU32(smemRaw);
    const uint32_t base = (rawAddr + 1023u) & ~1023u;
    uint8_t* sm = smemRaw + (base - rawAddr);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    setupCommon(a, base, sm, warp, kEpi == EPI_K_TMA_GROUPS ? kEpiWarps / 2 : kEpiWarps);
    const uint32_t tmemBase = *reinterpret_cast<volatile uint32_t*>(sm + kOffSlot);
    const uint32_t stage0 = base + a.headerBytes + a.stagingBytes;
    const uint32_t stageBytes = a.bytesA + a.bytesB;
    const uint32_t barFull = base + kOffFull, barEmpty = base + kOffEmpty, barTFull = base + kOffTFull, barTEmpty = base + kOffTEmpty;
    const int first = blockIdx.x, step = gridDim.x;
    const int nMine = tilesForCta(a, first, step);
    pdlLaunchDependents();
    pdlWait();  // setup above touched constants only (bias, tensor maps); activations and SE-scaled weights come from earlier kernels

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            TileWalker w;
            w.init(first, step, a.tilesN, a.tilesX, a.tilesY);
            for (int k = 0; k < nMine; ++k, w.next()) {
                const TileCoord tc = w.coord(a.bh, a.bw, a.bn, 0);
                for (int tap = 0; tap < a.p.ntaps; ++tap) {
                    const ConvTap tp = a.p.tap[tap];
                    for (int cc = 0; cc < a.cchunks; ++cc) {
                        mbarWait(barEmpty + 8u * stage, phase ^ 1u);
                        const uint32_t full = barFull + 8u * stage;
                        const uint32_t dstA = stage0 + stage * stageBytes;
                        mbarExpectTx(full, stageBytes);
                        tmaLoad5d(dstA, &a.tmA, full, tp.c0 + cc * a.kc, tc.x0 + tp.dx, tp.dz, tc.y0 + tp.dy, tc.img);
                        tmaLoad3d(dstA + a.bytesA, &a.tmB, full, tap * a.p.cin + cc * a.kc, tc.n0, a.p.w_img_stride ? tc.img : 0);
                        if (++stage == a.stages) { stage = 0; phase ^= 1u; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // whole warp converged; one elected lane issues the MMAs and commits (keeps the issue loop on the uniform datapath)
        // The issue thread's time between two MMAs is a bubble in the tensor pipe (its queue is shallow: measured, W2X_PROF), so the
        // loop keeps everything but the MMAs themselves off that path: the NEXT K block's full barrier is probed (non-blocking)
        // before this block's MMAs are queued, and the accumulator-full commit rides in the same elected section as the last MMAs.
        int stage = 0, acc = 0;
        uint32_t phase = 0, accPhase = 0;
        const int kSteps = a.kc / 16;
        uint32_t fullOk = 0;
        for (int k = 0; k < nMine; ++k) {
            mbarWait(barTEmpty + 8u * acc, accPhase ^ 1u);
            const uint32_t tmemD = tmemBase + (uint32_t)(acc * a.bn);
            for (int kb = 0; kb < a.kblocks; ++kb) {
                if (!fullOk) mbarWait(barFull + 8u * stage, phase);
                tcFenceAfter();
                int nstage = stage + 1;
                uint32_t nphase = phase;
                if (nstage == a.stages) { nstage = 0; nphase ^= 1u; }
                fullOk = mbarTest(barFull + 8u * nstage, nphase);
                if (electOne()) {
                    const uint32_t aLo = descLo(stage0 + stage * stageBytes), bLo = descLo(stage0 + stage * stageBytes + a.bytesA);
                    for (int ks = 0; ks < kSteps; ++ks)
                        ummaLoHi(tmemD, aLo + 2u * ks, a.descHiA, bLo + 2u * ks, a.descHiB, a.idesc, (kb | ks) != 0 ? 1u : 0u);
                    tcCommit(barEmpty + 8u * stage);
                    if (kb == a.kblocks - 1) tcCommit(barTFull + 8u * acc);
                }
                __syncwarp();
                stage = nstage; phase = nphase;
            }
            if (++acc == 2) { acc = 0; accPhase ^= 1u; }
        }
    } else {
        if constexpr (kEpi == EPI_K_TMA_GROUPS) epilogueTmaGroups(a, base, tmemBase, nMine, first, step, 0);
        else epilogueWarps<kEpi>(a, base, tmemBase, nMine, first, step, 0);
    }

    tcFenceBefore();
    __syncthreads();
    if (warp == 1) {
        tcFenceAfter();
        tmemDealloc(tmemBase, a.tmemCols);
    }
}

// ======================================================================================================================
// 3x3 convolution from one input patch per tile, weights resident in shared memory
// ======================================================================================================================
template <int kEpi, int kKC, bool kFused = false>
__global__ void __launch_bounds__(kFused ? kThreads + 32 * kFuseWarps : kThreads, 1) conv3x3_patch_kernel(const __grid_constant__ ConvArgs a) {
    extern __shared__ uint8_t smemRaw[];
    const uint32_t rawAddr = smemU32(smemRaw);
    const uint32_t base = (rawAddr + 1023u) & ~1023u;
    uint8_t* sm = smemRaw + (base - rawAddr);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    setupCommon(a, base, sm, warp, kEpi == EPI_K_TMA_GROUPS ? kEpiWarps / 2 : kEpiWarps, kFused ? kFuseWarps : 1);
    const uint32_t tmemBase = *reinterpret_cast<volatile uint32_t*>(sm + kOffSlot);
    const uint32_t wBase = base + a.headerBytes + a.stagingBytes;
    const uint32_t stage0 = wBase + a.wBytes;
    const uint32_t barFull = base + kOffFull, barEmpty = base + kOffEmpty, barTFull = base + kOffTFull, barTEmpty = base + kOffTEmpty;
    const uint32_t barW = base + kOffW;
    // this CTA owns output-channel slice `slice` and every (gridDim/nSplit)-th pixel tile
    const int slice = blockIdx.x % a.nSplit, first = blockIdx.x / a.nSplit, step = gridDim.x / a.nSplit;
    const int n0 = slice * a.bn;
    const int nMine = tilesForCta(a, first, step);
    const uint32_t rowBytes = (uint32_t)a.kc * 2u;     // one pixel / one weight row of a K chunk: 128 B (kc=64) or 64 B (kc=32)
    const uint32_t tapBytes = (uint32_t)a.bn * rowBytes;  // one (tap, K chunk) block of B: [bn rows][kc]

    pdlLaunchDependents();
    if (warp == 0 && lane == 0) {
        // resident weights (constants of the loaded model): 9 * cchunks TMA boxes on one barrier, issued before the dependency wait
        mbarExpectTx(barW, a.wBytes);
        for (int tap = 0; tap < 9; ++tap)
            for (int cc = 0; cc < a.cchunks; ++cc)
                tmaLoad3d(wBase + (uint32_t)(tap * a.cchunks + cc) * tapBytes, &a.tmB, barW, tap * a.p.cin + cc * a.kc, n0, 0);
    }
    pdlWait();  // the input activation (and the output buffer's previous readers) belong to earlier kernels

    if (kFused && warp >= kThreads / 32) {
        fusedFirstProducer(a, base, stage0, tmemBase, nMine);
    } else if (warp == 0 && kFused) {
        if (lane == 0) {
            // RGB patch loader: 20 rows x 12 pixels x 4 channels per tile, kRgbBufs tiles ahead of the producer warps
            const uint32_t rgb0 = base + a.fuseOff;
            TileWalker w;
            w.init(first, step, 1, a.tilesX, a.tilesY);
            for (int k = 0; k < nMine; ++k, w.next()) {
                const TileCoord tc = w.coord(a.bh, a.bw, a.bn, 0);
                const int rb = k % kRgbBufs;
                mbarWait(base + kOffRgbEmpty + 8u * rb, ((uint32_t)(k / kRgbBufs) & 1u) ^ 1u);
                mbarExpectTx(base + kOffRgbFull + 8u * rb, (uint32_t)((kPatchW + 2) * (kPatchH + 2) * 8));
                tmaLoad3d(rgb0 + (uint32_t)rb * 2048u, &a.tmRgb, base + kOffRgbFull + 8u * rb, tc.x0 * 4, tc.y0, tc.img + a.fuseImgBase);
            }
        }
    } else if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            TileWalker w;
            w.init(first, step, 1, a.tilesX, a.tilesY);
            W2X_PROF_DECL(pWaitEmpty);
            W2X_PROF_T(pT0);
            for (int k = 0; k < nMine; ++k, w.next()) {
                const TileCoord tc = w.coord(a.bh, a.bw, a.bn, n0);
                for (int cc = 0; cc < a.cchunks; ++cc) {
                    W2X_PROF_T(pa);
                    mbarWait(barEmpty + 8u * stage, phase ^ 1u);
                    W2X_PROF_ADD(pWaitEmpty, pa);
                    const uint32_t full = barFull + 8u * stage;
                    if (W2X_DBG_ON(a, 16)) {  // timing experiment: no activation traffic at all, the MMAs run on whatever the slot holds
                        mbarArrive(full);
                    } else {
                        mbarExpectTx(full, a.bytesA);
                        tmaLoad5d(stage0 + stage * a.stageStride, &a.tmA, full, cc * a.kc, tc.x0, 0, tc.y0, tc.img);
                    }
                    if (++stage == a.stages) { stage = 0; phase ^= 1u; }
                }
            }
#ifdef W2X_DEV
            if (a.prof) { a.prof[16ll * blockIdx.x + 4] = clock64() - pT0; a.prof[16ll * blockIdx.x + 5] = pWaitEmpty; }
#endif
        }
    } else if (warp == 1) {
        // MMA issuer.  The tensor pipe's queue is shallow, so every cycle this thread spends between two MMAs is a bubble (measured
        // with W2X_PROF: ~400 cycles per tile in the first version of this loop = 15 % of the layer).  Hence: the barriers of the NEXT
        // patch / tile are probed with non-blocking test_wait while the current MMAs are being queued (the blocking waits then fall
        // through), four accumulators make the accumulator-empty wait trivially true, and both commits ride in the elected section.
        int stage = 0, acc = 0;
        uint32_t phase = 0, accPhase = 0;
        mbarWait(barW, 0);
        const uint32_t bTapStep = ((uint32_t)a.cchunks * tapBytes) >> 4;  // descriptor-lo distance between consecutive taps of B
        const int nAcc = a.nAcc;
        uint32_t fullOk = 0, accOk = 0;
        // fused first layer: this warp also issues the producer's four M = 128, N = 32 MMAs per tile, one tile ahead of the main MMAs
        // (a second issuing thread would only queue behind this one: the pipe is shared)
        auto fusedFirstIssue = [&](int k) {
            const int fb = k % kFuseBufs;
            const uint32_t im = base + a.fuseOff + kRgbBufs * 2048u + (uint32_t)fb * kIm2colBytes;
            mbarWait(base + kOffImFull + 8u * fb, (uint32_t)(k / kFuseBufs) & 1u);
            tcFenceAfter();
            if (lane == 0) W2X_TRACE(a, k, 0);   // first-layer MMAs of tile k issued
            if (electOne()) {
                const uint32_t tD = tmemBase + (uint32_t)(a.nAcc * a.bn) + (uint32_t)(fb * 2 * kFuseN);
                const uint32_t bLo = descLo(base + a.fuseOff + kRgbBufs * 2048u + kFuseBufs * kIm2colBytes);
                const uint32_t hi64 = descHi(512u, 4u);       // SWIZZLE_64B, 8-row groups 512 bytes apart
                const uint32_t idesc1 = instrDescF16(128, kFuseN);
#pragma unroll
                for (int mb = 0; mb < 2; ++mb) {
                    const uint32_t aLo = descLo(im + (uint32_t)mb * 8192u);
#pragma unroll
                    for (int ks = 0; ks < 2; ++ks)
                        if (!W2X_DBG_ON(a, 32)) ummaLoHi(tD + (uint32_t)kFuseN * mb, aLo + 2u * ks, hi64, bLo + 2u * ks, hi64, idesc1, ks ? 1u : 0u);
                }
                tcCommit(base + kOffP1 + 8u * fb);
            }
            __syncwarp();
            if (lane == 0) W2X_TRACE(a, k, 7);   // first-layer MMAs of tile k: issue done
        };
        if (kFused)
            for (int k = 0; k < kFuseBufs - 1 && k < nMine; ++k) fusedFirstIssue(k);
        W2X_PROF_DECL(pWaitAcc); W2X_PROF_DECL(pWaitFull); W2X_PROF_DECL(pIssue);
        W2X_PROF_T(pT0);
#ifdef W2X_DEV
        unsigned long long gt0 = 0;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt0));
#endif
        for (int k = 0; k < nMine; ++k) {
            if (kFused && k + kFuseBufs - 1 < nMine) fusedFirstIssue(k + kFuseBufs - 1);
            W2X_PROF_T(pa);
            if (!accOk) mbarWait(barTEmpty + 8u * acc, accPhase ^ 1u);
            W2X_PROF_ADD(pWaitAcc, pa);
            const uint32_t tmemD = tmemBase + (uint32_t)(acc * a.bn);
            int nacc = acc + 1;
            uint32_t naccPhase = accPhase;
            if (nacc == nAcc) { nacc = 0; naccPhase ^= 1u; }
            for (int cc = 0; cc < a.cchunks; ++cc) {
                W2X_PROF_T(pb);
                if (!fullOk) mbarWait(barFull + 8u * stage, phase);
                tcFenceAfter();
                W2X_PROF_ADD(pWaitFull, pb);
                W2X_PROF_T(pc);
                if (lane == 0) W2X_TRACE(a, k, 1);   // main MMAs of tile k: issue starts
                int nstage = stage + 1;
                uint32_t nphase = phase;
                if (nstage == a.stages) { nstage = 0; nphase ^= 1u; }
                fullOk = mbarTest(barFull + 8u * nstage, nphase);
                const bool last = cc == a.cchunks - 1;
                if (last) accOk = mbarTest(barTEmpty + 8u * nacc, naccPhase ^ 1u);
                if (electOne()) {
                    // Nine taps = shifted views of the patch: whole pixel rows (kRowBytes each) into the swizzled tile.  The ky loop stays
                    // rolled on purpose: fully unrolled, the 36 precomputed descriptor pairs do not fit the uniform register file and the
                    // issue thread spends ~25 spill/fill instructions per MMA (measured with w2x_probe_mma_tiles: 58.6 vs 48 cycles per MMA).
                    constexpr uint32_t kRowBytes = kKC * 2;
                    uint32_t aRow = descLo(stage0 + stage * a.stageStride);
                    uint32_t b0 = descLo(wBase + (uint32_t)cc * tapBytes);
                    uint32_t accum = cc != 0 ? 1u : 0u;
                    const uint32_t descHiA = a.descHiA, descHiB = a.descHiB, idesc = a.idesc;
                    if (!W2X_DBG_ON(a, 4)) {
#pragma unroll 1
                        for (int ky = 0; ky < 3; ++ky) {
                            const uint32_t b1 = b0 + bTapStep, b2 = b1 + bTapStep;
#pragma unroll
                            for (int ks = 0; ks < kKC / 16; ++ks) { ummaLoHi(tmemD, aRow + 2u * ks, descHiA, b0 + 2u * ks, descHiB, idesc, accum); accum = 1u; }
#pragma unroll
                            for (int ks = 0; ks < kKC / 16; ++ks) ummaLoHi(tmemD, aRow + (kRowBytes >> 4) + 2u * ks, descHiA, b1 + 2u * ks, descHiB, idesc, 1u);
#pragma unroll
                            for (int ks = 0; ks < kKC / 16; ++ks) ummaLoHi(tmemD, aRow + 2u * (kRowBytes >> 4) + 2u * ks, descHiA, b2 + 2u * ks, descHiB, idesc, 1u);
                            aRow += (uint32_t)(kPatchW * kRowBytes) >> 4;
                            b0 = b2 + bTapStep;
                        }
                    }
                    tcCommit(barEmpty + 8u * stage);
                    if (last) tcCommit(barTFull + 8u * acc);
                }
                __syncwarp();
                W2X_PROF_ADD(pIssue, pc);
                if (lane == 0) W2X_TRACE(a, k, 2);   // main MMAs of tile k: issue done
                stage = nstage; phase = nphase;
            }
            acc = nacc; accPhase = naccPhase;
        }
#ifdef W2X_DEV
        if (a.prof && lane == 0) {
            unsigned long long gt1 = 0;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt1));
            long long* pr = a.prof + 16ll * blockIdx.x;
            pr[0] = clock64() - pT0; pr[1] = pWaitAcc; pr[2] = pWaitFull; pr[3] = pIssue; pr[10] = nMine; pr[11] = (long long)(gt1 - gt0);
        }
#endif
    } else {
        // tilesN == 1 for this kernel: the walker's N index stays 0 and the slice offset comes in as nBase
        if constexpr (kEpi == EPI_K_TMA_GROUPS) epilogueTmaGroups(a, base, tmemBase, nMine, first, step, n0);
        else epilogueWarps<kEpi>(a, base, tmemBase, nMine, first, step, n0);
    }

    tcFenceBefore();
    __syncthreads();
    if (warp == 1) {
        tcFenceAfter();
        tmemDealloc(tmemBase, a.tmemCols);
    }
}

#ifdef W2X_DEV
// ------------------------------------------------------------------------------------------------------------
// UMMA descriptor probe (development aid, reachable through w2x_probe_umma): one TMA-loaded, 128B-swizzled patch of
// 18 x 16 pixels x 64 channels; for every 3x3 tap the A operand is a SHIFTED VIEW of that patch (start address moved by
// (ky*16+kx) pixels = 128-byte rows, 8-row groups SBO = 16 pixels apart).  Checks which base_offset convention makes
// shifted views of a swizzled tile read the right data.  mode 0: base_offset = (start >> 7) & 7, mode 1: 0.
// ------------------------------------------------------------------------------------------------------------
struct ProbeArgs {
    CUtensorMap tmP, tmB;
    float* out;  // [9][128][64]
    int mode, pitch, sbo;
};

__global__ void __launch_bounds__(128, 1) umma_probe_kernel(const __grid_constant__ ProbeArgs a) {
    extern __shared__ uint8_t smemRaw[];
    const uint32_t rawAddr = smemU32(smemRaw);
    const uint32_t base = (rawAddr + 1023u) & ~1023u;
    uint8_t* sm = smemRaw + (base - rawAddr);
    const uint32_t bar = base, barMma = base + 8;
    volatile uint32_t* tmemSlot = reinterpret_cast<volatile uint32_t*>(sm + 32);
    const uint32_t sPatch = base + 1024, sB = sPatch + 40960;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mbarInit(bar, 1);
        mbarInit(barMma, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smemU32((const void*)tmemSlot)), "r"(64u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcFenceBefore();
    __syncthreads();
    tcFenceAfter();
    const uint32_t tmemBase = *tmemSlot;
    if (threadIdx.x == 0) {
        mbarExpectTx(bar, 18u * a.pitch * 128u + 64u * 128u);
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(sPatch),
                     "l"(&a.tmP), "r"(bar), "r"(0), "r"(0), "r"(0)
                     : "memory");
        tmaLoad2d(sB, &a.tmB, bar, 0, 0);
    }
    mbarWait(bar, 0);
    const uint32_t idesc = (1u << 4) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
    for (int tap = 0; tap < 9; ++tap) {
        if (threadIdx.x == 0) {
            tcFenceAfter();
            const int ky = tap / 3, kx = tap % 3;
            const uint32_t start = sPatch + (uint32_t)(ky * a.pitch + kx) * 128u;
            const uint32_t bo = a.mode == 0 ? ((start >> 7) & 7u) : 0u;
            const uint32_t descHiA = ((uint32_t)a.sbo >> 4) | (1u << 14) | (bo << 17) | (2u << 29);
            const uint32_t descHiB = (1024u >> 4) | (1u << 14) | (2u << 29);
            for (int k = 0; k < 4; ++k) umma(tmemBase, makeDesc(start + 32u * k, descHiA), makeDesc(sB + 32u * k, descHiB), idesc, k != 0);
            tcCommit(barMma);
        }
        mbarWait(barMma, tap & 1);
        tcFenceAfter();
        uint32_t r[32];
        for (int c0 = 0; c0 < 64; c0 += 32) {
            tmemLd32(tmemBase + ((uint32_t)(warp * 32) << 16) + c0, r);
            tmemLdWait();
            for (int i = 0; i < 32; ++i) a.out[((size_t)tap * 128 + warp * 32 + lane) * 64 + c0 + i] = __uint_as_float(r[i]);
        }
        tcFenceBefore();
        __syncthreads();
    }
    if (warp == 0) {
        tcFenceAfter();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmemBase), "r"(64u) : "memory");
    }
}


// ------------------------------------------------------------------------------------------------------------
// UMMA issue-rate probe (development aid, w2x_probe_mma_rate): every SM issues `iters` x 4 back-to-back
// tcgen05.mma (M = 128, N = n, K = 16) on the same smem operands; cycles per MMA = time * clock / (4 * iters).
// ------------------------------------------------------------------------------------------------------------
// With streamBytes > 0 a third warp concurrently streams an L2-resident buffer into a separate shared-memory ring with
// cp.async.bulk (four copies of streamBytes in flight) for as long as the MMAs run: measures how much the tensor core's operand
// reads and asynchronous shared-memory writes slow each other down.  out[2*cta] = MMA cycles, out[2*cta+1] = bytes streamed.
__global__ void __launch_bounds__(128, 1) umma_rate_kernel(int n, int iters, int sboA, const uint8_t* streamSrc, int streamBytes, long long* out) {
    extern __shared__ uint8_t smemRaw[];
    const uint32_t rawAddr = smemU32(smemRaw);
    const uint32_t base = (rawAddr + 1023u) & ~1023u;
    uint8_t* sm = smemRaw + (base - rawAddr);
    const uint32_t bar = base;
    volatile uint32_t* tmemSlot = reinterpret_cast<volatile uint32_t*>(sm + 32);
    volatile uint32_t* doneFlag = reinterpret_cast<volatile uint32_t*>(sm + 40);
    const uint32_t sA = base + 1024, sB = sA + 32768, sRing = sB + 32768;
    for (uint32_t i = threadIdx.x; i < (32768u + 32768u) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sm + 1024)[i] = 0u;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        mbarInit(bar, 1);
        for (int i = 0; i < 4; ++i) mbarInit(base + 64 + 8u * i, 1);
        mbarInitFence();
        *doneFlag = 0;
    }
    if (warp == 0) tmemAlloc(smemU32((const void*)tmemSlot), 256);
    fenceProxyAsync();
    tcFenceBefore();
    __syncthreads();
    tcFenceAfter();
    const uint32_t tmemBase = *tmemSlot;
    if (warp == 1) {
        const long long t0 = clock64();
        if (electOne()) {
            const uint32_t idesc = instrDescF16(128, n);
            const uint32_t hiA = descHi((uint32_t)sboA, 2), hiB = descHi(1024, 2);
            // streamBytes < 0: no stream, A operand starts -streamBytes - 1 rows (128 B each) into the tile, like a shifted 3x3 tap view
            const uint32_t aLo = descLo(sA + (streamBytes < 0 ? (uint32_t)(-streamBytes - 1) * 128u : 0u)), bLo = descLo(sB);
            for (int it = 0; it < iters; ++it) {
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) ummaLoHi(tmemBase, aLo + 2u * ks, hiA, bLo + 2u * ks, hiB, idesc, 1u);
            }
            tcCommit(bar);
        }
        __syncwarp();
        mbarWait(bar, 0);
        if (threadIdx.x == 32) {
            if (out) out[2 * blockIdx.x] = clock64() - t0;
            *doneFlag = 1;
        }
    } else if (warp == 2 && threadIdx.x == 64 && streamBytes > 0) {
        long long copied = 0;
        int k = 0;
        for (;; ++k) {
            const int sl = k & 3;
            const uint32_t b = base + 64 + 8u * sl;
            if (k >= 4) mbarWait(b, (uint32_t)((k - 4) >> 2) & 1u);
            if (*doneFlag) break;
            mbarExpectTx(b, (uint32_t)streamBytes);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sRing + (uint32_t)sl * (uint32_t)streamBytes),
                         "l"(streamSrc), "r"(streamBytes), "r"(b) : "memory");
            copied += streamBytes;
        }
        // drain the copies still in flight before the CTA (and its shared memory) goes away
        for (int j = 1; j <= 3 && k - j >= 0; ++j) {
            const int kk = k - j;
            if (kk + 4 > k - 0 && kk >= 0) mbarWait(base + 64 + 8u * (kk & 3), (uint32_t)(kk >> 2) & 1u);
        }
        if (out) out[2 * blockIdx.x + 1] = copied;
    }
    mbarWait(bar, 0);
    tcFenceBefore();
    __syncthreads();
    if (warp == 0) { tcFenceAfter(); tmemDealloc(tmemBase, 256); }
}

// Tile-loop probe (w2x_probe_mma_tiles): the patch kernel's MMA schedule without any data movement or epilogue -- per "tile" 36
// UMMAs (nine tap views of one 18x10 patch x four K steps, N = 64) into alternating TMEM accumulators, one commit per tile.
// mode bit 0: wait for the commit of tile t-2 before issuing tile t (the accumulator hand-back of the real kernel, with a
// zero-latency epilogue); bit 1: every tap uses the same B tile; bit 2: every tap uses the unshifted A view;
// bit 3: commit only once at the end; bit 4: always the same accumulator; bit 5: never overwrite (accumulate = 1 throughout).  out[cta] = cycles from first issue to last completion.
__global__ void __launch_bounds__(128, 1) umma_tile_probe_kernel(int tiles, int mode, long long* out) {
    extern __shared__ uint8_t smemRaw[];
    const uint32_t rawAddr = smemU32(smemRaw);
    const uint32_t base = (rawAddr + 1023u) & ~1023u;
    uint8_t* sm = smemRaw + (base - rawAddr);
    volatile uint32_t* tmemSlot = reinterpret_cast<volatile uint32_t*>(sm + 48);
    const uint32_t sA = base + 1024, sB = sA + 24576;
    for (uint32_t i = threadIdx.x; i < (24576u + 73728u) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sm + 1024)[i] = 0u;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        mbarInit(base, 1);
        mbarInit(base + 8, 1);
        mbarInit(base + 16, 1);
        mbarInit(base + 24, 1);
        mbarInit(base + 32, 1);
        mbarInitFence();
    }
    if (warp == 0) tmemAlloc(smemU32((const void*)tmemSlot), 128);
    fenceProxyAsync();
    tcFenceBefore();
    __syncthreads();
    tcFenceAfter();
    const uint32_t tmemBase = *tmemSlot;
    // mode bit 8 (256): two issuing warps, warp 1 takes the even tiles (accumulator 0) and warp 2 the odd ones (accumulator 1)
    const bool dual = (mode & 256) != 0;
    if (warp == 1 || (dual && warp == 2)) {
        const int which = warp - 1;
        const long long t0 = clock64();
        const uint32_t idesc = instrDescF16(128, 64);
        const uint32_t hiA = descHi(kPatchW * 128u, 2), hiB = descHi(1024, 2);
        const uint32_t aLo = descLo(sA), bLo0 = descLo(sB);
        uint32_t ph[2] = {0, 0};
        for (int t = dual ? which : 0; t < tiles; t += dual ? 2 : 1) {
            const int acc = (mode & 16) ? 0 : (t & 1);
            // (dual: each warp owns one accumulator, so the hand-back wait is for its own previous tile)
            if ((mode & 1) && t >= 2) { mbarWait(base + 8u * acc, ph[acc]); ph[acc] ^= 1u; tcFenceAfter(); }
            if (electOne()) {
                const uint32_t tmemD = tmemBase + (uint32_t)(acc * 64);
                uint32_t bLo = bLo0;
                if (mode & 128) {
                    // the patch kernel's issue loop: ky rolled, kx and ks unrolled (12 MMAs per iteration)
                    uint32_t aRow = aLo, b0 = bLo0, accum = 0u;
                    const uint32_t step = 8192u >> 4;
#pragma unroll 1
                    for (int ky = 0; ky < 3; ++ky) {
                        const uint32_t b1 = b0 + step, b2 = b1 + step;
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) { ummaLoHi(tmemD, aRow + 2u * ks, hiA, b0 + 2u * ks, hiB, idesc, accum); accum = 1u; }
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) ummaLoHi(tmemD, aRow + 8u + 2u * ks, hiA, b1 + 2u * ks, hiB, idesc, 1u);
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) ummaLoHi(tmemD, aRow + 16u + 2u * ks, hiA, b2 + 2u * ks, hiB, idesc, 1u);
                        aRow += (uint32_t)(kPatchW * 128) >> 4;
                        b0 = b2 + step;
                    }
                } else if (mode & 64) {
                    // rolled tap loops: a handful of live uniform registers instead of 36 precomputed descriptor pairs
                    uint32_t aRow = aLo, accum = (mode & 32) ? 1u : 0u;
#pragma unroll 1
                    for (int ky = 0; ky < 3; ++ky) {
#pragma unroll 1
                        for (int kx = 0; kx < 3; ++kx) {
                            const uint32_t aTap = (mode & 4) ? aLo : aRow + (uint32_t)kx * (128u >> 4);
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks) {
                                ummaLoHi(tmemD, aTap + 2u * ks, hiA, bLo + 2u * ks, hiB, idesc, accum);
                                accum = 1u;
                            }
                            if (!(mode & 2)) bLo += 8192u >> 4;
                        }
                        aRow += (uint32_t)(kPatchW * 128) >> 4;
                    }
                } else {
#pragma unroll
                    for (int tap = 0; tap < 9; ++tap) {
                        const uint32_t aTap = aLo + ((mode & 4) ? 0u : (uint32_t)(((tap / 3) * kPatchW + (tap % 3)) * 128 >> 4));
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) ummaLoHi(tmemD, aTap + 2u * ks, hiA, bLo + 2u * ks, hiB, idesc, ((tap | ks) != 0 || (mode & 32)) ? 1u : 0u);
                        if (!(mode & 2)) bLo += 8192u >> 4;
                    }
                }
                if (mode & 512) tcCommit(base + 32);            // a second commit per tile, like the kernel's slot-empty + accumulator-full pair
                if (mode & 1) tcCommit(base + 8u * acc);       // accumulator hand-back
                else if (!(mode & 8)) tcCommit(base + 24);      // a commit per tile that nobody waits for (its cost only)
            }
            __syncwarp();
        }
        // one more commit on a third barrier: it completes when every MMA issued above has completed
        if (electOne()) tcCommit(base + 16 + 16u * which);
        __syncwarp();
        mbarWait(base + 16 + 16u * which, 0);
        if ((threadIdx.x & 31) == 0 && out) out[2 * blockIdx.x + which] = clock64() - t0;
    }
    tcFenceBefore();
    __syncthreads();
    if (warp == 0) { tcFenceAfter(); tmemDealloc(tmemBase, 128); }
}

#endif  // W2X_DEV

// ---- host side -------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encodeTiled() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p)
            throw Error("cuTensorMapEncodeTiled is not available from the CUDA driver");
        fn = (EncodeTiledFn)p;
    }
    return fn;
}

int numSMs() {
    static int n = 0;
    if (!n) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

constexpr size_t kSmemLimit = 227 * 1024;

void checkCuda(cudaError_t e) {
    if (e != cudaSuccess) throw Error(std::string("cuda: ") + cudaGetErrorString(e));
}

// 5-D fp16 tensor map over (c, x, z, y, img) with element strides for x, z, y, img
void encode5d(CUtensorMap* tm, const void* ptr, const long long dims[5], const long long stridesElems[4], const int box[5], bool sw128,
              const char* what) {
    cuuint64_t d[5], s[4];
    cuuint32_t b[5], es[5] = {1, 1, 1, 1, 1};
    for (int i = 0; i < 5; ++i) { d[i] = (cuuint64_t)dims[i]; b[i] = (cuuint32_t)box[i]; }
    for (int i = 0; i < 4; ++i) s[i] = (cuuint64_t)stridesElems[i] * 2;
    CUresult r = encodeTiled()(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, (void*)ptr, d, s, b, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               sw128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw Error(std::string("cuTensorMapEncodeTiled(") + what + ") failed with code " + std::to_string((int)r));
}

void encodeWeights(CUtensorMap* tm, const ConvParams& p, int kc, int rows, bool sw128) {
    // (k, n, img): img > 0 only when the layer carries per-image weights (SE scale folded in)
    const bool perImage = p.w_img_stride != 0;
    cuuint64_t dims[3] = {(cuuint64_t)p.ktot, (cuuint64_t)p.npad, (cuuint64_t)(perImage ? p.gn : 1)};
    cuuint64_t strides[2] = {(cuuint64_t)p.ktot * 2, (cuuint64_t)(perImage ? p.w_img_stride : (long long)p.npad * p.ktot) * 2};
    cuuint32_t box[3] = {(cuuint32_t)kc, (cuuint32_t)rows, 1};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = encodeTiled()(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, (void*)p.w, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               sw128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw Error("cuTensorMapEncodeTiled(weights) failed with code " + std::to_string((int)r));
}

// can the epilogue go through swizzled smem staging + TMA store?
bool tmaEpilogueOk(const ConvParams& p) {
    if (p.mode == EPI_STORE) return p.cout % 64 == 0 && p.npad == p.cout && p.out_c == p.cout && p.act == ACT_LRELU && !p.skip;
    if (p.mode == EPI_D2S) {
        if (p.cout % 64 != 0 || p.out_c != p.cout || (p.out_w & 1) || (p.out_h & 1)) return false;
        if (p.skip && ((p.skip_off & 1) || (p.skip_w & 1) || (p.skip_h & 1) || p.skip_c != p.cout)) return false;
        return true;
    }
    return false;
}

void encodeOutMaps(ConvArgs& a) {
    const ConvParams& p = a.p;
    const int box[5] = {64, a.bw, 1, a.bh, 1};
    if (p.mode == EPI_STORE) {
        const long long dims[5] = {p.out_c, p.out_w, 1, p.out_h, p.gn};
        const long long st[4] = {p.out_c, (long long)p.out_w * p.out_c, (long long)p.out_w * p.out_c, (long long)p.out_h * p.out_w * p.out_c};
        encode5d(&a.tmOut, p.out, dims, st, box, true, "out");
    } else {  // EPI_D2S: (2C | W/2 | dy | H/2 | img) view so that the depth-to-space scatter is a plain box store
        const long long C = p.cout;
        const long long dims[5] = {2 * C, p.out_w / 2, 2, p.out_h / 2, p.gn};
        const long long st[4] = {2 * C, (long long)p.out_w * C, 2ll * p.out_w * C, (long long)p.out_h * p.out_w * C};
        encode5d(&a.tmOut, p.out, dims, st, box, true, "out(d2s)");
        if (p.skip) {
            const long long sd[5] = {2 * C, p.skip_w / 2, 2, p.skip_h / 2, p.gn};
            const long long ss[4] = {2 * C, (long long)p.skip_w * C, 2ll * p.skip_w * C, (long long)p.skip_h * p.skip_w * C};
            encode5d(&a.tmSkip, p.skip, sd, ss, box, true, "skip(d2s)");
        }
    }
}

// N tile: the largest divisor of npad that is a multiple of 32 and <= 256 (16-wide heads stay 16)
int pickBn(int npad) {
    if (npad <= 32) return npad;
    for (int bn = 256; bn >= 32; bn -= 32)
        if (npad % bn == 0) return bn;
    return 0;
}

bool wantsPatchKernel(const ConvParams& p) {
    if (!p.is3x3 || !(p.cin == 32 || p.cin == 64 || p.cin == 128) || p.w_img_stride) return false;
    if (!(p.npad % 64 == 0 || p.npad == 16)) return false;
    if (p.npad % 64 == 0 && !tmaEpilogueOk(p)) return false;
    return true;
}

uint32_t headerBytesFor(int npad) { return 1024u + (((uint32_t)npad * 4u + 1023u) & ~1023u) + 1024u; }

void planPatch(IgemmPlan* plan) {
    ConvArgs& a = plan->args;
    const ConvParams& p = a.p;
    a.headerBytes = headerBytesFor(p.npad);
    plan->patch = true;
    a.kc = p.cin % 64 == 0 ? 64 : 32;
    const bool sw128 = a.kc == 64;
    a.bn = p.npad % 64 == 0 ? 64 : 16;
    a.nSplit = p.npad / a.bn;
    a.cchunks = p.cin / a.kc;
    a.kblocks = a.cchunks;
    a.bw = 8; a.bh = 16; a.bwShift = 3;
    a.tilesX = (p.gx + 7) / 8;
    a.tilesY = (p.gy + 15) / 16;
    a.tilesN = 1;
    a.totalTiles = a.tilesX * a.tilesY * p.gn;  // per output-channel slice
    a.useTma = a.bn == 64 ? 1 : 0;
    a.hasSkip = 0;
    a.nsub = 1;
    a.nbuf = a.useTma ? 2 : 1;
    a.stagingBytes = a.useTma ? a.nbuf * 16384u : 0u;
    a.wBytes = 9u * a.cchunks * a.bn * a.kc * 2u;
    a.bytesA = (uint32_t)(kPatchW * kPatchH) * a.kc * 2u;     // 23040 (kc=64) / 11520 (kc=32)
    a.stageStride = (a.bytesA + 1023u) & ~1023u;
    const size_t fixed = 1024 + a.headerBytes + a.stagingBytes + a.wBytes;
    if (fixed + 2 * (size_t)a.stageStride > kSmemLimit) throw Error("conv3x3 patch kernel: weights do not fit in shared memory");
    a.stages = (int)std::min<size_t>(8, (kSmemLimit - fixed) / a.stageStride);
    // the two-group epilogue (TMA store, two staging buffers) alternates between two accumulators per group: four in all
    a.nAcc = (a.useTma && a.nbuf == 2 && 4 * a.bn <= 512) ? 4 : 2;
    uint32_t cols = 32;
    while (cols < (uint32_t)(a.nAcc * a.bn)) cols *= 2;
    a.tmemCols = cols;
    a.idesc = instrDescF16(128, a.bn);
    const uint32_t rowBytes = a.kc * 2u, layout = sw128 ? 2u : 4u;
    a.descHiA = ((kPatchW * rowBytes) >> 4) | (1u << 14) | (layout << 29);  // 8-pixel row groups are one patch row (10 pixels) apart
    a.descHiB = ((8u * rowBytes) >> 4) | (1u << 14) | (layout << 29);
    a.bytesB = 0;
    {
        const long long dims[5] = {p.dimc, p.dimx, p.dimz, p.dimy, p.gn};
        const long long st[4] = {p.sx, p.sz, p.sy, p.sn};
        const int box[5] = {a.kc, kPatchW, 1, kPatchH, 1};
        encode5d(&a.tmA, p.in, dims, st, box, sw128, "patch");
    }
    encodeWeights(&a.tmB, p, a.kc, a.bn, sw128);
    if (a.useTma) encodeOutMaps(a);
    const int sms = numSMs();
    int grid = std::min(sms, a.totalTiles * a.nSplit);
    grid = std::max(a.nSplit, grid / a.nSplit * a.nSplit);
    plan->grid = grid;
    plan->smem = fixed + (size_t)a.stages * a.stageStride;
}

void planIgemm(IgemmPlan* plan) {
    ConvArgs& a = plan->args;
    const ConvParams& p = a.p;
    a.headerBytes = headerBytesFor(p.npad);
    plan->patch = false;
    a.kc = (p.cin % 64 == 0) ? 64 : 32;
    a.useTma = tmaEpilogueOk(p) ? 1 : 0;
    a.hasSkip = (a.useTma && p.mode == EPI_D2S && p.skip) ? 1 : 0;
    a.bn = pickBn(p.npad);
    if (p.mode == EPI_D2S && a.useTma) a.bn = (p.cout % 128 == 0 || 128 % p.cout == 0) ? 128 : 64;  // whole 64-channel sub-tiles of one phase
    // token-wise layers whose shape rules out the TMA epilogue: coalescing staged epilogue
    const bool skipSameGeom = !p.skip || (p.skip_off == 0 && p.skip_c == p.out_c && p.skip_h == p.out_h && p.skip_w == p.out_w && !p.skip_scale);
    a.staged = (!a.useTma && a.bn >= 32 && skipSameGeom &&
                ((p.mode == EPI_STORE && p.cout == p.npad && p.out_c == p.npad) ||
                 (p.mode == EPI_D2S && p.cout % 8 == 0 && p.out_c == p.cout && p.npad == 4 * p.cout && p.act == ACT_LRELU))) ? 1 : 0;
    a.twoPerSm = 0;
    if (a.staged && p.ntaps == 1) {
        // two CTAs per SM: N tiles of at most 128 columns (2 x 2 x 128 accumulator columns fit the SM's 512), and a K chunk whose
        // three-deep ring + staging fits half the shared memory
        const size_t halfAvail = kSmemLimit / 2 - 2048 - 1024 - a.headerBytes;
        for (int bn2 = 128; bn2 >= 32 && !a.twoPerSm; bn2 -= 32) {
            if (p.npad % bn2) continue;
            for (int kc : {a.kc, 32}) {
                if (p.cin % kc) continue;
                const size_t stageB = (size_t)(128 + bn2) * kc * 2, staging = 2 * (((size_t)128 * (bn2 * 2 + 16) + 1023) & ~(size_t)1023);
                if (halfAvail >= staging + 3 * stageB) { a.twoPerSm = 1; a.bn = bn2; a.kc = kc; break; }
            }
        }
    }
    a.nSplit = 1;
    a.cchunks = p.cin / a.kc;
    a.kblocks = p.ntaps * a.cchunks;
    long long best = -1;
    for (int bw = 8; bw <= 128; bw *= 2) {
        const int bh = 128 / bw;
        const long long cover = (long long)((p.gx + bw - 1) / bw) * bw * ((p.gy + bh - 1) / bh) * bh;
        if (best < 0 || cover < best) { best = cover; a.bw = bw; a.bh = bh; }
    }
    a.bwShift = 0;
    while ((1 << a.bwShift) < a.bw) ++a.bwShift;
    a.tilesX = (p.gx + a.bw - 1) / a.bw;
    a.tilesY = (p.gy + a.bh - 1) / a.bh;
    a.tilesN = p.npad / a.bn;
    a.totalTiles = a.tilesX * a.tilesY * a.tilesN * p.gn;
    a.bytesA = 128u * a.kc * 2u;
    a.bytesB = (uint32_t)a.bn * a.kc * 2u;
    a.nsub = a.useTma ? a.bn / 64 : 0;
    const size_t stageBytes = a.bytesA + a.bytesB;
    const size_t avail = kSmemLimit - 1024 - a.headerBytes;
    a.nbuf = 1;
    a.stagingBytes = 0;
    if (a.staged) {
        a.stagedPitch = (uint32_t)a.bn * 2u + 16u;
        a.stagedBuf = (128u * a.stagedPitch + 1023u) & ~1023u;
        a.nbuf = 2;
        a.stagingBytes = 2 * a.stagedBuf;
    }
    if (a.useTma) {
        const size_t buf = (size_t)a.nsub * 16384;
        int nbuf = a.hasSkip ? 3 : 2;
        const int minBuf = a.hasSkip ? 2 : 1;
        while (nbuf > minBuf && (avail - nbuf * buf) / stageBytes < 3) --nbuf;
        a.nbuf = nbuf;
        a.stagingBytes = (uint32_t)(nbuf * buf);
    }
    if (avail < a.stagingBytes + 2 * stageBytes) throw Error("igemm: tile does not fit in shared memory");
    a.stages = (int)std::min<size_t>(8, (avail - a.stagingBytes) / stageBytes);
    const bool twoPerSm = a.twoPerSm != 0;
    if (twoPerSm) a.stages = (int)std::min<size_t>(6, (kSmemLimit / 2 - 2048 - 1024 - a.headerBytes - a.stagingBytes) / stageBytes);
    a.nAcc = 2;
    uint32_t cols = 32;
    while (cols < (uint32_t)(2 * a.bn)) cols *= 2;
    a.tmemCols = cols;
    a.idesc = instrDescF16(128, a.bn);
    const bool sw128 = a.kc == 64;
    const uint32_t sbo = sw128 ? 1024u : 512u;
    a.descHiA = a.descHiB = (sbo >> 4) | (1u << 14) | ((sw128 ? 2u : 4u) << 29);
    {
        const long long dims[5] = {p.dimc, p.dimx, p.dimz, p.dimy, p.gn};
        const long long st[4] = {p.sx, p.sz, p.sy, p.sn};
        const int box[5] = {a.kc, a.bw, 1, a.bh, 1};
        encode5d(&a.tmA, p.in, dims, st, box, sw128, "A");
    }
    encodeWeights(&a.tmB, p, a.kc, a.bn, sw128);
    if (a.useTma) encodeOutMaps(a);
    plan->grid = std::min(a.totalTiles, (twoPerSm ? 2 : 1) * numSMs());
    plan->smem = 1024 + a.headerBytes + a.stagingBytes + (size_t)a.stages * stageBytes;
}

}  // namespace

bool igemmSupported(const ConvParams& p) {
    if (p.cin % 32 != 0 || p.npad % 16 != 0 || p.npad > 1024) return false;
    if (p.npad > 32 && p.npad % 32 != 0) return false;
    if (p.ntaps < 1 || p.ntaps > 9) return false;
    return true;
}

IgemmPlan* igemmCreatePlan(const ConvParams& p) {
    if (!igemmSupported(p)) throw Error("igemm: unsupported layer shape");
    IgemmPlan* plan = new IgemmPlan();
    plan->args = ConvArgs{};
    plan->args.p = p;
    try {
        static const bool noPatch = [] { const char* e = devEnv("W2X_NO_PATCH"); return e && *e == '1'; }();
        if (!noPatch && wantsPatchKernel(p)) planPatch(plan);
        else planIgemm(plan);
        if (p.se_sum && !igemmSeFusable(plan)) throw Error("igemm: fused SE squeeze is not available for this layer shape");
        static bool attrSet[64] = {};  // cudaFuncSetAttribute is per device (one engine per GPU in one process: row-band mode)
        int dev = 0;
        cudaGetDevice(&dev);
        if (dev < 0 || dev >= 64) dev = 0;
        if (!attrSet[dev]) {
            checkCuda(cudaFuncSetAttribute(igemm_kernel<EPI_K_DIRECT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemLimit));
            checkCuda(cudaFuncSetAttribute(igemm_kernel<EPI_K_TMA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemLimit));
            checkCuda(cudaFuncSetAttribute(igemm_kernel<EPI_K_TMA_GROUPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemLimit));
            checkCuda(cudaFuncSetAttribute(conv3x3_patch_kernel<EPI_K_TMA_GROUPS, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemLimit));
            checkCuda(cudaFuncSetAttribute(conv3x3_patch_kernel<EPI_K_TMA_GROUPS, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemLimit));
            checkCuda(cudaFuncSetAttribute(conv3x3_patch_kernel<EPI_K_TMA_GROUPS, 32, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemLimit));
            checkCuda(cudaFuncSetAttribute(igemm_kernel<EPI_K_TMA_SKIP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemLimit));
            checkCuda(cudaFuncSetAttribute(igemm_kernel<EPI_K_STAGED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemLimit));
            checkCuda(cudaFuncSetAttribute(igemm_kernel<EPI_K_STAGED2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemLimit));
            checkCuda(cudaFuncSetAttribute(conv3x3_patch_kernel<EPI_K_DIRECT, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemLimit));
            checkCuda(cudaFuncSetAttribute(conv3x3_patch_kernel<EPI_K_TMA, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemLimit));
            checkCuda(cudaFuncSetAttribute(conv3x3_patch_kernel<EPI_K_DIRECT, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemLimit));
            checkCuda(cudaFuncSetAttribute(conv3x3_patch_kernel<EPI_K_TMA, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemLimit));
            attrSet[dev] = true;
        }
    } catch (...) {
        delete plan;
        throw;
    }
    return plan;
}

// conv1.conv.2 of a UNet with its RGB first layer computed inside the kernel (fusedFirstProducer).  `first` is the 4 -> 32 layer
// whose output tensor `second` would have read; that tensor is never written.
bool igemmFusedFirstSupported(const ConvParams& second, const ConvParams& first) {
    // the RGB patches are TMA boxes over [img][H][W*4]: rows must be a multiple of 16 bytes and the base 16-byte aligned
    if (first.sy != (long long)first.dimx * 4 || (first.dimx & 1) || (reinterpret_cast<uintptr_t>(first.in) & 15)) return false;
    return first.is3x3 && first.cin == 4 && first.ktot == 36 && first.npad == 32 && first.mode == EPI_STORE && first.act == ACT_LRELU && !first.skip &&
           first.w_img_stride == 0 && first.sx == 4 && second.is3x3 && second.cin == 32 && second.in == first.out && second.dimx == first.gx &&
           second.dimy == first.gy && second.npad == 64 && second.mode == EPI_STORE && tmaEpilogueOk(second) && !second.se_sum && wantsPatchKernel(second);
}

IgemmPlan* igemmCreatePlanFusedFirst(const ConvParams& second, const ConvParams& first) {
    if (!igemmFusedFirstSupported(second, first)) throw Error("igemm: layer pair cannot be fused");
    IgemmPlan* plan = igemmCreatePlan(second);
    ConvArgs& a = plan->args;
    while (a.stages > 3 && (a.stages > 6 || plan->smem + kFuseExtraBytes > kSmemLimit)) {  // the producer stays two tiles ahead: six ring slots are plenty
        --a.stages;
        plan->smem -= a.stageStride;
    }
    if (!plan->patch || a.kc != 32 || !a.useTma || a.nbuf != 2 || plan->smem + kFuseExtraBytes > kSmemLimit) {
        igemmDestroyPlan(plan);
        throw Error("igemm: fused first layer does not fit this plan");
    }
    a.fused = 1;
    a.nAcc = 2;                                        // TMEM: 2 x 64 accumulator columns + kFuseBufs x 64 for the producer's first-layer tiles
    a.tmemCols = 512;
    a.fuseIn = first.in;
    a.fuseW = first.w;
    a.fuseBias = first.bias;
    a.fuseSlope = first.slope;
    a.fuseW_px = first.dimx;
    a.fuseH_px = first.dimy;
    a.fuseSn = first.sn;
    a.fuseOff = a.headerBytes + a.stagingBytes + a.wBytes + (uint32_t)a.stages * a.stageStride;
    {
        cuuint64_t dims[3] = {(cuuint64_t)first.dimx * 4, (cuuint64_t)first.dimy, (cuuint64_t)first.gn};
        cuuint64_t strides[2] = {(cuuint64_t)first.dimx * 8, (cuuint64_t)first.sn * 2};
        cuuint32_t box[3] = {(cuuint32_t)(kPatchW + 2) * 4, (cuuint32_t)(kPatchH + 2), 1};
        cuuint32_t es[3] = {1, 1, 1};
        if (first.sy != (long long)first.dimx * 4 || (first.dimx & 1) || (reinterpret_cast<uintptr_t>(first.in) & 15)) {
            igemmDestroyPlan(plan);
            throw Error("igemm: fused first layer needs a dense, 16-byte aligned NHWC4 input");
        }
        const CUresult r = encodeTiled()(&a.tmRgb, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, (void*)first.in, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            igemmDestroyPlan(plan);
            throw Error("cuTensorMapEncodeTiled(rgb) failed with code " + std::to_string((int)r));
        }
    }
    plan->smem += kFuseExtraBytes;  // RGB patch ring, two im2col tiles, first-layer weights + bias
    return plan;
}

void igemmDestroyPlan(IgemmPlan* plan) { delete plan; }

// Fused first layer: let the plan read its RGB tiles from a frame-wide buffer [images][H][W][4] (all tiles of a frame unpacked by
// one launch) instead of the per-batch input tensor; igemmLaunch(..., inOverride) then selects the batch inside it.
void igemmSetFusedFrameInput(IgemmPlan* plan, const __half* base, long long images) {
    ConvArgs& a = plan->args;
    if (!a.fused) throw Error("igemm: not a fused first-layer plan");
    if ((reinterpret_cast<uintptr_t>(base) & 15) || images < 1) throw Error("igemm: frame input must be 16-byte aligned");
    cuuint64_t dims[3] = {(cuuint64_t)a.fuseW_px * 4, (cuuint64_t)a.fuseH_px, (cuuint64_t)images};
    cuuint64_t strides[2] = {(cuuint64_t)a.fuseW_px * 8, (cuuint64_t)a.fuseSn * 2};
    cuuint32_t box[3] = {(cuuint32_t)(kPatchW + 2) * 4, (cuuint32_t)(kPatchH + 2), 1};
    cuuint32_t es[3] = {1, 1, 1};
    const CUresult r = encodeTiled()(&plan->tmRgbFrame, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, (void*)base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw Error("cuTensorMapEncodeTiled(rgb frame) failed with code " + std::to_string((int)r));
    plan->frameBase = base;
    plan->frameImages = images;
}

// 128B-swizzled tensor map over the layer's input view (c, x, z, y, img) with a (64, boxX, 1, boxY, 1) box, for kernels outside
// this file that stage activation patches with TMA (tm points at a CUtensorMap)
void encodeActivationMap5d(void* tm, const ConvParams& p, int boxX, int boxY) {
    const long long dims[5] = {p.dimc, p.dimx, p.dimz, p.dimy, p.gn};
    const long long st[4] = {p.sx, p.sz, p.sy, p.sn};
    const int box[5] = {64, boxX, 1, boxY, 1};
    encode5d(static_cast<CUtensorMap*>(tm), p.in, dims, st, box, true, "activation patch");
}

// swizzled 2-D tensor map over a K-major fp16 matrix [rows][k] with a (boxK, boxRows) box (64-byte swizzle for boxK = 32, 128-byte for
// boxK = 64), for kernels outside this file that keep weight matrices resident in shared memory as UMMA operands
void encodeMatrixMap2d(void* tm, const void* ptr, long long k, long long rows, int boxK, int boxRows, bool sw128) {
    cuuint64_t dims[2] = {(cuuint64_t)k, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)k * 2};
    cuuint32_t box[2] = {(cuuint32_t)boxK, (cuuint32_t)boxRows};
    cuuint32_t es[2] = {1, 1};
    const CUresult r = encodeTiled()(static_cast<CUtensorMap*>(tm), CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, es,
                                     CU_TENSOR_MAP_INTERLEAVE_NONE, sw128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw Error("cuTensorMapEncodeTiled(matrix) failed with code " + std::to_string((int)r));
}

bool igemmSeFusable(const IgemmPlan* plan) {
    const ConvArgs& a = plan->args;
    return a.useTma && !a.hasSkip && a.nbuf >= 2 && a.tilesN == 1 && a.bn >= 64;
}

const char* igemmDescribe(const IgemmPlan* plan, char* buf, int cap) {
    const ConvArgs& a = plan->args;
    std::snprintf(buf, cap, "%s bn=%d kc=%d tile=%dx%d stages=%d nbuf=%d tma=%d skip=%d staged=%d split=%d grid=%d smem=%zu", a.fused ? "patch3x3+rgb-first-layer" : plan->patch ? "patch3x3" : "igemm",
                  a.bn, a.kc, a.bh, a.bw, a.stages, a.nbuf, a.useTma, a.hasSkip, a.staged, a.nSplit, plan->grid, plan->smem);
    return buf;
}

void igemmLaunch(const IgemmPlan* plan, cudaStream_t s, __half* outOverride, int nImages, const __half* inOverride) {
    if (plan->grid <= 0) return;
    const bool fewer = nImages > 0 && nImages < plan->args.p.gn;  // last, partially filled batch: skip the padding slots
    const bool redirect = outOverride && outOverride != plan->args.p.out;
    if (redirect && plan->args.staged) throw Error("igemm: output redirection is not available for staged-epilogue layers");
    if (redirect && plan->args.useTma) throw Error("igemm: output redirection is not available for TMA-store layers");
    ConvArgs local;
    const ConvArgs* a = &plan->args;
    const bool frameIn = inOverride && plan->args.fused && inOverride != plan->args.fuseIn;
    if (inOverride && !plan->args.fused && inOverride != plan->args.p.in) throw Error("igemm: input redirection is only available for the fused first layer");
    if (frameIn) {
        const long long off = inOverride - plan->frameBase;
        if (!plan->frameBase || off < 0 || off % plan->args.fuseSn || off / plan->args.fuseSn + (nImages > 0 ? nImages : plan->args.p.gn) > plan->frameImages)
            throw Error("igemm: input override lies outside the registered frame buffer");
    }
    if (redirect || fewer || frameIn) {
        local = plan->args;  // only the epilogue's destination / the image count change; the tensor maps stay valid
        if (redirect) local.p.out = outOverride;
        if (frameIn) {
            local.tmRgb = plan->tmRgbFrame;
            local.fuseImgBase = (int)((inOverride - plan->frameBase) / plan->args.fuseSn);
        }
        if (fewer) {
            local.totalTiles = local.totalTiles / local.p.gn * nImages;
            local.p.gn = nImages;
        }
        a = &local;
    }
    // TMA-store layers without a skip tensor and with two staging buffers run the two-group epilogue
    static const bool noGroups = devEnv("W2X_NO_EPI_GROUPS") != nullptr;
    static const int dbg = devEnv("W2X_DBG") ? std::atoi(devEnv("W2X_DBG")) : 0;
    ConvArgs dbgLocal;
    if (dbg) { dbgLocal = *a; dbgLocal.dbg = dbg; a = &dbgLocal; }
#ifdef W2X_DEV
    // W2X_PROF=1: per-role cycle counters of the patch kernel, averaged over the CTAs and printed after a synchronising launch
    static const bool profOn = devEnv("W2X_PROF") != nullptr;
    long long* profBuf = nullptr;
    ConvArgs profLocal;
    if (profOn && plan->patch) {
        checkCuda(cudaMalloc(&profBuf, sizeof(long long) * (16 * plan->grid + 8 * kTraceTiles)));
        checkCuda(cudaMemsetAsync(profBuf, 0, sizeof(long long) * (16 * plan->grid + 8 * kTraceTiles), s));
        profLocal = *a; profLocal.prof = profBuf; a = &profLocal;
    }
#endif
    const bool grouped = a->useTma && !a->hasSkip && a->nbuf == 2 && a->bn <= 128 && !noGroups;
    ConvArgs accLocal;
    if (!grouped && a->nAcc != 2) { accLocal = *a; accLocal.nAcc = 2; a = &accLocal; }  // only the two-group epilogue handles four accumulators
    if (plan->patch) {
        if (a->kc == 64) {
            if (grouped) launchPdl(conv3x3_patch_kernel<EPI_K_TMA_GROUPS, 64>, dim3(plan->grid), dim3(kThreads), plan->smem, s, *a);
            else if (a->useTma) launchPdl(conv3x3_patch_kernel<EPI_K_TMA, 64>, dim3(plan->grid), dim3(kThreads), plan->smem, s, *a);
            else launchPdl(conv3x3_patch_kernel<EPI_K_DIRECT, 64>, dim3(plan->grid), dim3(kThreads), plan->smem, s, *a);
        } else {
            if (grouped && a->fused) launchPdl(conv3x3_patch_kernel<EPI_K_TMA_GROUPS, 32, true>, dim3(plan->grid), dim3(kThreads + 32 * kFuseWarps), plan->smem, s, *a);
            else if (a->fused) throw Error("fused first layer needs the two-group TMA epilogue");
            else if (grouped) launchPdl(conv3x3_patch_kernel<EPI_K_TMA_GROUPS, 32>, dim3(plan->grid), dim3(kThreads), plan->smem, s, *a);
            else if (a->useTma) launchPdl(conv3x3_patch_kernel<EPI_K_TMA, 32>, dim3(plan->grid), dim3(kThreads), plan->smem, s, *a);
            else launchPdl(conv3x3_patch_kernel<EPI_K_DIRECT, 32>, dim3(plan->grid), dim3(kThreads), plan->smem, s, *a);
        }
    } else {
        if (a->hasSkip) launchPdl(igemm_kernel<EPI_K_TMA_SKIP>, dim3(plan->grid), dim3(kThreads), plan->smem, s, *a);
        else if (grouped) launchPdl(igemm_kernel<EPI_K_TMA_GROUPS>, dim3(plan->grid), dim3(kThreads), plan->smem, s, *a);
        else if (a->useTma) launchPdl(igemm_kernel<EPI_K_TMA>, dim3(plan->grid), dim3(kThreads), plan->smem, s, *a);
        else if (a->staged && a->twoPerSm) launchPdl(igemm_kernel<EPI_K_STAGED2>, dim3(plan->grid), dim3(kThreads), plan->smem, s, *a);
        else if (a->staged) launchPdl(igemm_kernel<EPI_K_STAGED>, dim3(plan->grid), dim3(kThreads), plan->smem, s, *a);
        else launchPdl(igemm_kernel<EPI_K_DIRECT>, dim3(plan->grid), dim3(kThreads), plan->smem, s, *a);
    }
#ifdef W2X_DEV
    if (profBuf) {
        checkCuda(cudaStreamSynchronize(s));
        std::vector<long long> h(16 * (size_t)plan->grid + 8 * kTraceTiles);
        checkCuda(cudaMemcpy(h.data(), profBuf, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
        cudaFree(profBuf);
        double sum[16] = {0};
        long long maxTotal = 0;
        for (int c = 0; c < plan->grid; ++c) {
            for (int i = 0; i < 16; ++i) sum[i] += (double)h[16 * c + i];
            maxTotal = std::max(maxTotal, h[16 * c]);
        }
        const double g = plan->grid, tiles = sum[10] / g, mmas = tiles * a->cchunks * 9 * (a->kc / 16);
        std::fprintf(stderr,
                     "[w2x prof] cin=%d cout=%d %dx%d n=%d | tiles/CTA %.1f | MMA warp: total %.0f (max %lld) = waitAcc %.0f + waitFull %.0f + issue %.0f ; %.1f cyc/MMA, clock %.3f GHz | "
                     "producer: total %.0f waitEmpty %.0f | epi g0: total %.0f waitTFull %.0f math %.0f sync+store %.0f\n",
                     a->p.cin, a->p.npad, a->p.gx, a->p.gy, a->p.gn, tiles, sum[0] / g, maxTotal, sum[1] / g, sum[2] / g, sum[3] / g, sum[0] / g / std::max(mmas, 1.0),
                     sum[11] > 0 ? sum[0] / sum[11] : 0.0, sum[4] / g, sum[5] / g, sum[6] / g, sum[7] / g, sum[8] / g, sum[9] / g);
        if (a->fused && devEnv("W2X_TRACE")) {
            const long long* tr = h.data() + 16 * (size_t)plan->grid;
            long long t0 = tr[0];
            std::fprintf(stderr, "[w2x trace] CTA 0, cycles since the first event: tile | L1 issued | main issue start | main issue end | producer waits P1 | has P1 | patch published | acc complete\n");
            for (int t = 0; t < kTraceTiles; ++t)
                std::fprintf(stderr, "[w2x trace] %2d | %7lld (done %7lld) | %7lld | %7lld | %7lld | %7lld | %7lld | %7lld\n", t, tr[t * 8 + 0] - t0, tr[t * 8 + 7] - t0, tr[t * 8 + 1] - t0, tr[t * 8 + 2] - t0,
                             tr[t * 8 + 3] - t0, tr[t * 8 + 4] - t0, tr[t * 8 + 5] - t0, tr[t * 8 + 6] - t0);
        }
        if (a->fused)
            std::fprintf(stderr, "[w2x prof]   first-layer producer: total %.0f = im2col %.0f + wait P1 %.0f + wait slot %.0f + convert %.0f  (per tile %.0f / %.0f / %.0f / %.0f)\n", sum[12] / g,
                         sum[13] / g, sum[14] / g, sum[5] / g, sum[15] / g, sum[13] / g / std::max(tiles, 1.0), sum[14] / g / std::max(tiles, 1.0), sum[5] / g / std::max(tiles, 1.0), sum[15] / g / std::max(tiles, 1.0));
    }
#endif
}


#ifdef W2X_DEV
// returns 0 on success; err9[tap] = max |device - host| for the given base_offset mode
int probeUmma(int mode, int pitch, float* err9) {
    if (pitch % 8 != 0 && mode < 2) { /* allowed: explores non-1024 SBO */ }
    const int rows = 18, pix = rows * pitch;
    std::vector<uint16_t> hp((size_t)pix * 64), hb(64 * 64);
    uint32_t st = 12345u;
    auto rnd = [&]() { st = st * 1664525u + 1013904223u; return ((st >> 9) & 0xFFFF) / 65536.0f - 0.5f; };
    for (auto& v : hp) v = floatToHalfBits(rnd());
    for (auto& v : hb) v = floatToHalfBits(rnd());
    __half *dP = nullptr, *dB = nullptr;
    float* dOut = nullptr;
    if (cudaMalloc(&dP, hp.size() * 2) || cudaMalloc(&dB, hb.size() * 2) || cudaMalloc(&dOut, 9 * 128 * 64 * 4)) return 1;
    cudaMemcpy(dP, hp.data(), hp.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice);
    cudaMemset(dOut, 0, 9 * 128 * 64 * 4);
    ProbeArgs a{};
    a.out = dOut; a.mode = mode; a.pitch = pitch; a.sbo = pitch * 128;
    {
        cuuint64_t dims[3] = {64, (cuuint64_t)pitch, (cuuint64_t)rows};
        cuuint64_t strides[2] = {128, (cuuint64_t)pitch * 128};
        cuuint32_t box[3] = {64, (cuuint32_t)pitch, (cuuint32_t)rows};
        cuuint32_t es[3] = {1, 1, 1};
        if (encodeTiled()(&a.tmP, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, dP, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return 2;
        cuuint64_t dimsB[2] = {64, 64};
        cuuint64_t stridesB[1] = {128};
        cuuint32_t boxB[2] = {64, 64};
        if (encodeTiled()(&a.tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, dB, dimsB, stridesB, boxB, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return 2;
    }
    cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    umma_probe_kernel<<<1, 128, 2048 + 40960 + 8192, nullptr>>>(a);
    if (cudaDeviceSynchronize() != cudaSuccess) return 3;
    std::vector<float> ho(9 * 128 * 64);
    cudaMemcpy(ho.data(), dOut, ho.size() * 4, cudaMemcpyDeviceToHost);
    for (int tap = 0; tap < 9; ++tap) {
        const int ky = tap / 3, kx = tap % 3;
        double md = 0;
        for (int m = 0; m < 128; ++m)
            for (int n = 0; n < 64; ++n) {
                double acc = 0;
                const int px = (ky + m / 8) * pitch + kx + (m % 8);
                for (int c = 0; c < 64; ++c) acc += (double)halfBitsToFloat(hp[(size_t)px * 64 + c]) * halfBitsToFloat(hb[(size_t)n * 64 + c]);
                md = std::max(md, std::fabs(acc - ho[((size_t)tap * 128 + m) * 64 + n]));
            }
        err9[tap] = (float)md;
    }
    cudaFree(dP); cudaFree(dB); cudaFree(dOut);
    return 0;
}



// ms for `iters` x 4 MMAs per SM (all SMs busy); returns < 0 on error
float probeMmaRate(int n, int iters, int sboA) {
    if (n < 16 || n > 256 || n % 16) return -1.f;
    cudaFuncSetAttribute(umma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    umma_rate_kernel<<<numSMs(), 128, 70 * 1024>>>(n, 16, sboA, nullptr, 0, nullptr);  // warm-up
    cudaEventRecord(e0);
    umma_rate_kernel<<<numSMs(), 128, 70 * 1024>>>(n, iters, sboA, nullptr, 0, nullptr);
    cudaEventRecord(e1);
    if (cudaEventSynchronize(e1) != cudaSuccess) return -2.f;
    float ms = -1.f;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return ms;
}

// cycles per MMA of the tile-loop probe (see umma_tile_probe_kernel); negative on error
float probeMmaTiles(int tiles, int mode) {
    if (tiles < 2 || tiles > 100000 || (tiles & 1)) return -1.f;
    const int sms = numSMs();
    cudaFuncSetAttribute(umma_tile_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 104 * 1024);
    long long* out = nullptr;
    if (cudaMalloc(&out, sizeof(long long) * 2 * sms) != cudaSuccess) return -2.f;
    cudaMemset(out, 0, sizeof(long long) * 2 * sms);
    umma_tile_probe_kernel<<<sms, 128, 102 * 1024>>>(4, mode, nullptr);
    umma_tile_probe_kernel<<<sms, 128, 102 * 1024>>>(tiles, mode, out);
    float res = -3.f;
    if (cudaDeviceSynchronize() == cudaSuccess) {
        std::vector<long long> h(2 * (size_t)sms);
        cudaMemcpy(h.data(), out, sizeof(long long) * 2 * sms, cudaMemcpyDeviceToHost);
        double cyc = 0;
        for (int i = 0; i < sms; ++i) cyc += (double)std::max(h[2 * i], h[2 * i + 1]);
        res = (float)(cyc / sms / (36.0 * tiles));
    }
    cudaFree(out);
    return res;
}

// UMMA issue rate while a second warp streams `streamBytes`-sized L2-resident copies into shared memory (0 = off).
// res[0] = average SM cycles per MMA, res[1] = average streamed bytes per SM cycle.  Returns 0 on success.
int probeMmaRateStream(int n, int iters, int streamBytes, float* res) {
    if (n < 16 || n > 256 || n % 16 || streamBytes < -64 || streamBytes > 16384 || (streamBytes > 0 && (streamBytes & 15))) return -1;
    const int sms = numSMs();
    const size_t smem = 70 * 1024 + 4 * (size_t)(streamBytes > 0 ? streamBytes : 0);
    cudaFuncSetAttribute(umma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 140 * 1024);
    uint8_t* src = nullptr;
    long long* out = nullptr;
    if (cudaMalloc(&src, 16384) != cudaSuccess || cudaMalloc(&out, sizeof(long long) * 2 * sms) != cudaSuccess) return -2;
    cudaMemset(src, 0, 16384);
    cudaMemset(out, 0, sizeof(long long) * 2 * sms);
    const int sbo = streamBytes < 0 ? 1280 : 1024;  // shifted-view experiment uses the patch kernel's 10-pixel row-group pitch
    umma_rate_kernel<<<sms, 128, smem>>>(n, 16, sbo, src, streamBytes, nullptr);  // warm-up
    umma_rate_kernel<<<sms, 128, smem>>>(n, iters, sbo, src, streamBytes, out);
    int rc = cudaDeviceSynchronize() == cudaSuccess ? 0 : -3;
    std::vector<long long> h(2 * (size_t)sms);
    if (rc == 0) cudaMemcpy(h.data(), out, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
    double cyc = 0, by = 0;
    for (int i = 0; i < sms; ++i) { cyc += (double)h[2 * i]; by += (double)h[2 * i + 1]; }
    if (rc == 0 && cyc > 0) { res[0] = (float)(cyc / sms / (4.0 * iters)); res[1] = (float)(by / cyc); }
    cudaFree(src); cudaFree(out);
    return rc;
}
#endif  // W2X_DEV
}  // namespace w2x
