// CUDA-core convolution kernels.
//
//  * conv_direct_kernel: scalar reference of the implicit-GEMM contract in conv_params.h.  NOT on the product path:
//    it exists so the tcgen05 kernels can be checked on the device at full layer sizes (w2x_selftest_conv, tests).
//  * conv_first_kernel: the RGB first layers (cin = 3 stored as 4, K = 27): far below a UMMA K-block and bound by the
//    32-channel fp16 store, so it uses warp-level mma.sync with register-resident weight fragments (see below).
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "conv_params.h"
#include "launch.h"

namespace w2x {

__global__ void __launch_bounds__(256) conv_direct_kernel(ConvParams p) {
    const int chunks = p.npad / 8;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long pix = idx / chunks;
    const int jc = (int)(idx - pix * chunks);
    if (pix >= (long long)p.gn * p.gy * p.gx) return;
    const int x = (int)(pix % p.gx);
    const int y = (int)((pix / p.gx) % p.gy);
    const int img = (int)(pix / ((long long)p.gx * p.gy));
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int t = 0; t < p.ntaps; ++t) {
        const ConvTap tp = p.tap[t];
        const __half* a = p.in + (long long)img * p.sn + (long long)(y + tp.dy) * p.sy + (long long)tp.dz * p.sz +
                          (long long)(x + tp.dx) * p.sx + tp.c0;
        const __half* w = p.w + (long long)img * p.w_img_stride + (long long)(jc * 8) * p.ktot + (long long)t * p.cin;
        for (int c = 0; c < p.cin; ++c) {
            const float av = __half2float(a[c]);
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = fmaf(av, __half2float(w[(long long)i * p.ktot + c]), acc[i]);
        }
    }
    conv_epilogue8(p, img, y, x, jc * 8, acc);
}

void launchConvDirect(const ConvParams& p, cudaStream_t s) {
    const long long total = (long long)p.gn * p.gy * p.gx * (p.npad / 8);
    const int block = 256;
    conv_direct_kernel<<<(unsigned)((total + block - 1) / block), block, 0, s>>>(p);
}

// ---- first layer: NHWC4 fp16 -> 32 channels, 3x3 valid, bias + LeakyReLU -----------------------------------
// K = 27 is far below a UMMA K block and the layer is bound by its 64-byte-per-pixel store, so it runs on warp-level
// mma.sync (m16n8k16, fp16 in / fp32 accumulate): one warp = 16 consecutive output pixels x 32 channels, three K steps
// (one per filter row: 3 taps x 4 stored channels + one zero-weight pad tap).  A fragments are loaded straight from the
// NHWC4 tile (4-byte loads, L1-resident), the weight fragments live in registers for the warp's lifetime, and a quad
// transpose (3 shuffles) turns the accumulator layout into 16-byte, fully coalesced channel-contiguous stores.
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

constexpr int kFirstRows = 32;  // output rows per work unit (a warp walks down a 16-pixel-wide column strip)

template <int NT>  // n-tiles of 8 output channels: 4 (CUNet, 32 ch) or 8 (SwinUNet patch embed, 64 stored ch)
__global__ void __launch_bounds__(256) conv_first_kernel(ConvParams p, int segsX, int chunksY, int totalUnits) {
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int warpId = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int warpCount = gridDim.x * (blockDim.x >> 5);
    pdlLaunchDependents();
    // B fragments: b[ky][j][0..1]; k index kk -> (kx = kk >> 2, ci = kk & 3); kx == 3 is the zero pad tap
    uint32_t bf[3][NT][2];
    float bias[NT][2];
#pragma unroll
    for (int j = 0; j < NT; ++j) {
        const int n = 8 * j + g;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            const __half* wrow = p.w + (long long)n * p.ktot + ky * 12;
            bf[ky][j][0] = *reinterpret_cast<const uint32_t*>(wrow + (t >> 1) * 4 + (t & 1) * 2);
            bf[ky][j][1] = (t >> 1) == 0 ? *reinterpret_cast<const uint32_t*>(wrow + 8 + (t & 1) * 2) : 0u;
        }
        bias[j][0] = p.bias[8 * j + 2 * t];
        bias[j][1] = p.bias[8 * j + 2 * t + 1];
    }
    // A fragment of one input row for this lane: pixels (x0+g, x0+g+8) and the same shifted by two taps.  Reads past the
    // row end only meet zero weights or discarded pixels (activation buffers carry slack).
    auto loadRow = [&](const __half* row, uint32_t (&a)[4]) {
        a[0] = *reinterpret_cast<const uint32_t*>(row);
        a[1] = *reinterpret_cast<const uint32_t*>(row + 8 * 4);
        a[2] = *reinterpret_cast<const uint32_t*>(row + 2 * 4);
        a[3] = *reinterpret_cast<const uint32_t*>(row + 10 * 4);
    };
    pdlWait();  // weights and bias above are constants; the input tiles come from the preceding kernel
    for (int unit = warpId; unit < totalUnits; unit += warpCount) {
        const int sx = unit % segsX;
        const int rest = unit / segsX;
        const int cy = rest % chunksY;
        const int img = rest / chunksY;
        const int x0 = sx * 16, y0 = cy * kFirstRows;
        const int y1 = min(p.gy, y0 + kFirstRows);
        const __half* in = p.in + (long long)img * p.sn + (long long)y0 * p.sy + (long long)(x0 + g + (t >> 1)) * 4 + (t & 1) * 2;
        __half* out = p.out + (((long long)img * p.out_h + y0) * p.out_w + x0 + g) * p.out_c + 2 * t;
        const bool okLo = x0 + g < p.gx, okHi = x0 + g + 8 < p.gx;
        uint32_t a0[4], a1[4], a2[4], a3[4];
        loadRow(in, a0);
        loadRow(in + p.sy, a1);
        loadRow(in + 2 * p.sy, a2);
        for (int y = y0; y < y1; ++y) {
            // the row after next is in flight while this row's MMAs run (the last prefetch of a strip stays inside the tile:
            // input rows y+3 <= gy+1 = H-1 ... except the very last, which is clamped)
            const int yn = min(y + 3, p.gy + 1) - y0;
            loadRow(in + (long long)yn * p.sy, a3);
            float d[NT][4];
#pragma unroll
            for (int j = 0; j < NT; ++j) { d[j][0] = bias[j][0]; d[j][1] = bias[j][1]; d[j][2] = bias[j][0]; d[j][3] = bias[j][1]; }
#pragma unroll
            for (int j = 0; j < NT; ++j) mma16816(d[j], a0, bf[0][j][0], bf[0][j][1]);
#pragma unroll
            for (int j = 0; j < NT; ++j) mma16816(d[j], a1, bf[1][j][0], bf[1][j][1]);
#pragma unroll
            for (int j = 0; j < NT; ++j) mma16816(d[j], a2, bf[2][j][0], bf[2][j][1]);
            // LeakyReLU + pack; a quad writes 16 contiguous bytes per (pixel, n-tile), four n-tiles complete the 64-byte pixel
#pragma unroll
            for (int j = 0; j < NT; ++j) {
                const __half2 l = __floats2half2_rn(fmaxf(d[j][0], d[j][0] * p.slope), fmaxf(d[j][1], d[j][1] * p.slope));
                const __half2 h = __floats2half2_rn(fmaxf(d[j][2], d[j][2] * p.slope), fmaxf(d[j][3], d[j][3] * p.slope));
                if (okLo) *reinterpret_cast<__half2*>(out + 8 * j) = l;
                if (okHi) *reinterpret_cast<__half2*>(out + 8 * p.out_c + 8 * j) = h;
            }
            out += (long long)p.out_w * p.out_c;
#pragma unroll
            for (int q = 0; q < 4; ++q) { a0[q] = a1[q]; a1[q] = a2[q]; a2[q] = a3[q]; }
        }
    }
}

void launchConvFirst(const ConvParams& p, cudaStream_t s) {
    // contract: 3x3 taps in (ky,kx) order on a plain NHWC4 view, ktot == 36, npad == out_c in {32, 64}, EPI_STORE
    const int segsX = (p.gx + 15) / 16, chunksY = (p.gy + kFirstRows - 1) / kFirstRows;
    const int total = segsX * chunksY * p.gn;
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0) sms = 148;
    }
    const int blocksNeeded = (total + 7) / 8;
    const int grid = blocksNeeded < sms * 3 ? blocksNeeded : sms * 3;
    if (p.npad == 64) launchPdl(conv_first_kernel<8>, dim3(grid), dim3(256), 0, s, p, segsX, chunksY, total);
    else launchPdl(conv_first_kernel<4>, dim3(grid), dim3(256), 0, s, p, segsX, chunksY, total);
}

#ifdef W2X_DEV
// Development probe (w2x_probe_hmma_rate): `iters` rounds of `chains` independent mma.sync.m16n8k16 (fp16 in, fp32 accumulate)
// per warp, `warps` warps per SM on every SM, operands in registers: the issue rate of the legacy tensor path that the
// first-layer, image-head and window-attention kernels use.  Returns milliseconds.
__global__ void hmma_rate_kernel(int iters, int chains, float* sink) {
    float d[8][4];
#pragma unroll
    for (int c = 0; c < 8; ++c) d[c][0] = d[c][1] = d[c][2] = d[c][3] = 0.f;
    const uint32_t a0 = 0x3c003c00u + threadIdx.x, a1 = 0x38003800u, a2 = 0x34003400u, a3 = 0x3c003800u, b0 = 0x38003c00u, b1 = 0x34003800u;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < 8; ++c)
            if (c < chains)
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(d[c][0]), "+f"(d[c][1]), "+f"(d[c][2]), "+f"(d[c][3])
                             : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) s += d[c][0] + d[c][1] + d[c][2] + d[c][3];
    if (s == 12345.678f) sink[0] = s;  // keeps the accumulators live
}

float probeHmmaRate(int warps, int chains, int iters) {
    if (warps < 1 || warps > 32 || chains < 1 || chains > 8 || iters < 1) return -1.f;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    float* sink = nullptr;
    if (cudaMalloc(&sink, 4) != cudaSuccess) return -2.f;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    hmma_rate_kernel<<<sms, warps * 32>>>(16, chains, sink);  // warm-up
    cudaEventRecord(e0);
    hmma_rate_kernel<<<sms, warps * 32>>>(iters, chains, sink);
    cudaEventRecord(e1);
    float ms = -3.f;
    if (cudaEventSynchronize(e1) == cudaSuccess) cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(sink);
    return ms;
}

// Development probe (w2x_probe_l2_stream): every SM streams the SAME `bytes`-sized global buffer (L2-resident after the first
// touch) into a four-slot shared-memory ring with cp.async.bulk, `iters` copies per SM, nothing else running: the L2 -> SM
// ingest rate available to a kernel that streams its weights instead of keeping them resident.  Returns milliseconds.
__global__ void l2_stream_kernel(const uint8_t* src, int bytes, int iters) {
    extern __shared__ __align__(128) uint8_t ring[];
    __shared__ __align__(8) unsigned long long bars[4];
    if (threadIdx.x != 0) return;
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(ring);
    uint32_t bar[4];
    for (int i = 0; i < 4; ++i) {
        bar[i] = (uint32_t)__cvta_generic_to_shared(&bars[i]);
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar[i]));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    for (int k = 0; k < iters + 4; ++k) {
        const int s = k & 3;
        if (k >= 4) {  // wait for the copy issued four iterations ago into this slot
            const uint32_t parity = (uint32_t)((k - 4) >> 2) & 1u;
            uint32_t done = 0;
            while (!done)
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar[s]), "r"(parity) : "memory");
        }
        if (k < iters) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar[s]), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(base + (uint32_t)s * (uint32_t)bytes),
                         "l"(src), "r"(bytes), "r"(bar[s]) : "memory");
        }
    }
}

float probeL2Stream(int bytes, int iters) {
    if (bytes < 1024 || bytes > 49152 || (bytes & 15) || iters < 1) return -1.f;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    uint8_t* src = nullptr;
    if (cudaMalloc(&src, bytes) != cudaSuccess) return -2.f;
    cudaMemset(src, 1, bytes);
    cudaFuncSetAttribute(l2_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 49152);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    l2_stream_kernel<<<sms, 32, 4 * bytes>>>(src, bytes, 8);  // warm-up
    cudaEventRecord(e0);
    l2_stream_kernel<<<sms, 32, 4 * bytes>>>(src, bytes, iters);
    cudaEventRecord(e1);
    float ms = -3.f;
    if (cudaEventSynchronize(e1) == cudaSuccess) cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(src);
    return ms;
}

#endif  // W2X_DEV
}  // namespace w2x
