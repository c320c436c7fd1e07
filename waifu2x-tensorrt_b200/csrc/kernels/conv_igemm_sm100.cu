// Implicit-GEMM convolution on the sm_100a tensor cores (tcgen05.mma, accumulators in TMEM, operands staged by TMA).
//
// This is the kernel behind every dense CUNet layer with cin >= 32, replacing the
// opaque TensorRT engine the reference enqueues at /root/reference/src/tensorrt/img2img_infer.cpp:80.
//
//   D[128 pixels, BN] (fp32, TMEM)  +=  A[128 pixels, KC] (fp16, smem, K-major, 128B/64B swizzle)
//                                     x B[BN, KC]         (fp16, smem, K-major, same swizzle)
//
// * A is never materialised (no im2col buffer): for each filter tap the TMA engine loads a BH x BW pixel box of the
//   NHWC activation, shifted by the tap offset, straight into the swizzled K-major operand layout (box rows are
//   pixels, the 64/128-byte inner box is the channel chunk).  Out-of-range pixels are zero-filled by TMA and their
//   rows are dropped in the epilogue, so valid convolutions of any (odd) size need no padding pass.
// * Persistent CTAs (one per SM), static round-robin tile schedule, warp-specialised: warp 0 = TMA producer,
//   warp 1 = TMEM allocator + single-thread tcgen05.mma issuer, warps 2..5 = epilogue (tcgen05.ld -> bias,
//   LeakyReLU, depth-to-space / skip-add / clamp -> global).  Two TMEM accumulator buffers so the epilogue of tile i
//   overlaps the MMAs of tile i+1; a multi-stage smem ring between TMA and MMA, all synchronised with mbarriers.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <string>

#include "../hostutil.h"
#include "conv_params.h"

namespace w2x {

struct IgemmArgs {
    CUtensorMap tmA;
    CUtensorMap tmB;
    ConvParams p;
    int kc, bn, bw, bh, bwShift;
    int tilesX, tilesY, tilesN, totalTiles;
    int stages, cchunks, kblocks;
    uint32_t idesc, tmemCols, bytesA, bytesB, descHi;
};

struct IgemmPlan {
    IgemmArgs args;
    int grid = 0;
    size_t smem = 0;
};

namespace {

constexpr int kThreads = 192;
constexpr int kHeaderBytes = 1024;

// ---- PTX wrappers ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smemU32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbarInit(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbarExpectTx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbarArrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbarWait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void tmaLoad5d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2, int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(dst),
        "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}
__device__ __forceinline__ void tmaLoad2d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
                 "l"(tm), "r"(bar), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tcFenceBefore() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcFenceAfter() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcCommit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma(uint32_t tmemD, uint64_t descA, uint64_t descB, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmemD),
        "l"(descA), "l"(descB), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmemLd32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmemLd16(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmemLdWait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major swizzled operand descriptor (cute::UMMA::SmemDescriptor layout): start>>4 [0,14), LBO>>4 [16,30),
// SBO>>4 [32,46), version=1 [46,48), layout type [61,64).
__device__ __forceinline__ uint64_t makeDesc(uint32_t smemAddr, uint32_t descHi) {
    return (uint64_t)((smemAddr >> 4) & 0x3FFFu) | (1ull << 16) | ((uint64_t)descHi << 32);
}

struct TileCoord {
    int img, y0, x0, n0;
};
__device__ __forceinline__ TileCoord decodeTile(const IgemmArgs& a, int t) {
    TileCoord c;
    const int nt = t % a.tilesN;
    t /= a.tilesN;
    const int tx = t % a.tilesX;
    t /= a.tilesX;
    const int ty = t % a.tilesY;
    c.img = t / a.tilesY;
    c.y0 = ty * a.bh;
    c.x0 = tx * a.bw;
    c.n0 = nt * a.bn;
    return c;
}

__global__ void __launch_bounds__(kThreads, 1) igemm_kernel(const __grid_constant__ IgemmArgs a) {
    extern __shared__ uint8_t smemRaw[];
    const uint32_t rawAddr = smemU32(smemRaw);
    const uint32_t base = (rawAddr + 1023u) & ~1023u;  // 1024-byte alignment for the 128B swizzle atoms
    uint8_t* sm = smemRaw + (base - rawAddr);
    // header: full[stages] | empty[stages] | tmemFull[2] | tmemEmpty[2] | tmem base slot
    const uint32_t barFull = base, barEmpty = base + 8u * a.stages;
    const uint32_t barTFull = base + 16u * a.stages, barTEmpty = barTFull + 16u;
    volatile uint32_t* tmemSlot = reinterpret_cast<volatile uint32_t*>(sm + 16 * a.stages + 32);
    const uint32_t stage0 = base + kHeaderBytes;
    const uint32_t stageBytes = a.bytesA + a.bytesB;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < a.stages; ++s) {
            mbarInit(barFull + 8u * s, 1);
            mbarInit(barEmpty + 8u * s, 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbarInit(barTFull + 8u * i, 1);
            mbarInit(barTEmpty + 8u * i, 4);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&a.tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&a.tmB) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smemU32((const void*)tmemSlot)), "r"(a.tmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcFenceBefore();
    __syncthreads();
    tcFenceAfter();
    const uint32_t tmemBase = *tmemSlot;

    if (warp == 0) {
        // ================= TMA producer (one thread) =================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int t = blockIdx.x; t < a.totalTiles; t += gridDim.x) {
                const TileCoord tc = decodeTile(a, t);
                for (int tap = 0; tap < a.p.ntaps; ++tap) {
                    const ConvTap tp = a.p.tap[tap];
                    for (int cc = 0; cc < a.cchunks; ++cc) {
                        mbarWait(barEmpty + 8u * stage, phase ^ 1u);
                        const uint32_t full = barFull + 8u * stage;
                        const uint32_t dstA = stage0 + stage * stageBytes;
                        mbarExpectTx(full, stageBytes);
                        tmaLoad5d(dstA, &a.tmA, full, tp.c0 + cc * a.kc, tc.x0 + tp.dx, tp.dz, tc.y0 + tp.dy, tc.img);
                        tmaLoad2d(dstA + a.bytesA, &a.tmB, full, tap * a.p.cin + cc * a.kc, tc.n0);
                        if (++stage == a.stages) { stage = 0; phase ^= 1u; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer (one thread) =================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t accPhase = 0;
            const int kSteps = a.kc / 16;
            for (int t = blockIdx.x; t < a.totalTiles; t += gridDim.x) {
                mbarWait(barTEmpty + 8u * acc, accPhase ^ 1u);
                tcFenceAfter();
                const uint32_t tmemD = tmemBase + (uint32_t)(acc * a.bn);
                for (int kb = 0; kb < a.kblocks; ++kb) {
                    mbarWait(barFull + 8u * stage, phase);
                    tcFenceAfter();
                    const uint32_t sA = stage0 + stage * stageBytes;
                    const uint32_t sB = sA + a.bytesA;
                    for (int k = 0; k < kSteps; ++k) {
                        umma(tmemD, makeDesc(sA + 32u * k, a.descHi), makeDesc(sB + 32u * k, a.descHi), a.idesc,
                             (kb | k) != 0 ? 1u : 0u);
                    }
                    tcCommit(barEmpty + 8u * stage);  // frees the smem slot when these MMAs retire
                    if (++stage == a.stages) { stage = 0; phase ^= 1u; }
                }
                tcCommit(barTFull + 8u * acc);  // accumulator complete
                if (++acc == 2) { acc = 0; accPhase ^= 1u; }
            }
        }
    } else {
        // ================= epilogue warps (TMEM lanes 32*(warp%4) .. +31) =================
        const int quarter = warp & 3;
        const int m = quarter * 32 + lane;
        const int yy = m >> a.bwShift, xx = m & (a.bw - 1);
        int acc = 0;
        uint32_t accPhase = 0;
        for (int t = blockIdx.x; t < a.totalTiles; t += gridDim.x) {
            const TileCoord tc = decodeTile(a, t);
            const int y = tc.y0 + yy, x = tc.x0 + xx;
            const bool valid = y < a.p.gy && x < a.p.gx;
            mbarWait(barTFull + 8u * acc, accPhase);
            tcFenceAfter();
            const uint32_t taddr = tmemBase + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * a.bn);
            uint32_t r[32];
            if (a.bn >= 32) {
                for (int c0 = 0; c0 < a.bn; c0 += 32) {
                    tmemLd32(taddr + (uint32_t)c0, r);
                    tmemLdWait();
                    if (valid) {
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            conv_epilogue8(a.p, tc.img, y, x, tc.n0 + c0 + 8 * q, reinterpret_cast<const float*>(r) + 8 * q);
                    }
                }
            } else {
                tmemLd16(taddr, r);
                tmemLdWait();
                if (valid) {
#pragma unroll
                    for (int q = 0; q < 2; ++q)
                        conv_epilogue8(a.p, tc.img, y, x, tc.n0 + 8 * q, reinterpret_cast<const float*>(r) + 8 * q);
                }
            }
            tcFenceBefore();
            __syncwarp();
            if (lane == 0) mbarArrive(barTEmpty + 8u * acc);
            if (++acc == 2) { acc = 0; accPhase ^= 1u; }
        }
    }

    tcFenceBefore();
    __syncthreads();
    if (warp == 1) {
        tcFenceAfter();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmemBase), "r"(a.tmemCols) : "memory");
    }
}

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

}  // namespace

bool igemmSupported(const ConvParams& p) {
    if (p.cin % 32 != 0 || p.npad % 16 != 0) return false;
    if (p.npad > 256 && p.npad % 256 != 0) return false;
    if (p.npad < 256 && (p.npad & (p.npad - 1)) != 0) return false;  // 16, 32, 64, 128
    if (p.ntaps < 1 || p.ntaps > 9) return false;
    return true;
}

IgemmPlan* igemmCreatePlan(const ConvParams& p) {
    if (!igemmSupported(p)) throw Error("igemm: unsupported layer shape");
    IgemmPlan* plan = new IgemmPlan();
    IgemmArgs& a = plan->args;
    a.p = p;
    a.kc = (p.cin % 64 == 0) ? 64 : 32;
    a.bn = std::min(p.npad, 256);
    a.cchunks = p.cin / a.kc;
    a.kblocks = p.ntaps * a.cchunks;
    // pixel tile: pick the BH x BW (= 128) split that wastes the fewest rows
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
    const size_t budget = 200 * 1024;
    a.stages = (int)std::min<size_t>(8, budget / (a.bytesA + a.bytesB));
    if (a.stages < 2) throw Error("igemm: stage does not fit in shared memory");
    uint32_t cols = 32;
    while (cols < (uint32_t)(2 * a.bn)) cols *= 2;
    a.tmemCols = cols;
    // instruction descriptor (cute::UMMA::InstrDescriptor): D=F32 [4,6), A=B=F16, both K-major, N>>3 [17,23), M>>4 [24,29)
    a.idesc = (1u << 4) | ((uint32_t)(a.bn >> 3) << 17) | ((128u >> 4) << 24);
    const bool sw128 = a.kc == 64;
    const uint32_t sbo = sw128 ? 1024u : 512u;  // 8 rows x (128 | 64) bytes
    a.descHi = (sbo >> 4) | (1u << 14) | ((sw128 ? 2u : 4u) << 29);

    // A: 5-D view (c, x, z, y, img)
    {
        cuuint64_t dims[5] = {(cuuint64_t)p.dimc, (cuuint64_t)p.dimx, (cuuint64_t)p.dimz, (cuuint64_t)p.dimy, (cuuint64_t)p.gn};
        cuuint64_t strides[4] = {(cuuint64_t)p.sx * 2, (cuuint64_t)p.sz * 2, (cuuint64_t)p.sy * 2, (cuuint64_t)p.sn * 2};
        cuuint32_t box[5] = {(cuuint32_t)a.kc, (cuuint32_t)a.bw, 1, (cuuint32_t)a.bh, 1};
        cuuint32_t es[5] = {1, 1, 1, 1, 1};
        CUresult r = encodeTiled()(&a.tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, (void*)p.in, dims, strides, box, es,
                                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { delete plan; throw Error("igemm: cuTensorMapEncodeTiled(A) failed with code " + std::to_string((int)r)); }
    }
    // B: [npad][ktot]
    {
        cuuint64_t dims[2] = {(cuuint64_t)p.ktot, (cuuint64_t)p.npad};
        cuuint64_t strides[1] = {(cuuint64_t)p.ktot * 2};
        cuuint32_t box[2] = {(cuuint32_t)a.kc, (cuuint32_t)a.bn};
        cuuint32_t es[2] = {1, 1};
        CUresult r = encodeTiled()(&a.tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void*)p.w, dims, strides, box, es,
                                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { delete plan; throw Error("igemm: cuTensorMapEncodeTiled(B) failed with code " + std::to_string((int)r)); }
    }
    plan->grid = std::min(a.totalTiles, numSMs());
    plan->smem = 1024 + kHeaderBytes + (size_t)a.stages * (a.bytesA + a.bytesB);
    static bool attrSet = false;
    if (!attrSet) {
        cudaError_t e = cudaFuncSetAttribute(igemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) { delete plan; throw Error(std::string("igemm: cudaFuncSetAttribute: ") + cudaGetErrorString(e)); }
        attrSet = true;
    }
    return plan;
}

void igemmDestroyPlan(IgemmPlan* plan) { delete plan; }

void igemmLaunch(const IgemmPlan* plan, cudaStream_t s, __half* outOverride) {
    if (plan->grid <= 0) return;
    if (outOverride && outOverride != plan->args.p.out) {
        IgemmArgs a = plan->args;  // only the epilogue's destination changes; the tensor maps stay valid
        a.p.out = outOverride;
        igemm_kernel<<<plan->grid, kThreads, plan->smem, s>>>(a);
    } else {
        igemm_kernel<<<plan->grid, kThreads, plan->smem, s>>>(plan->args);
    }
}

}  // namespace w2x
