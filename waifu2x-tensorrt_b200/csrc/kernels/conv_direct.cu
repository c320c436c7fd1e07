// CUDA-core convolution kernels.
//
//  * conv_direct_kernel: scalar reference of the implicit-GEMM contract in conv_params.h.  NOT on the product path:
//    it exists so the tcgen05 kernels can be checked on the device at full layer sizes (w2x_selftest_conv, tests).
//  * conv_first_kernel: the RGB first layers (cin = 3 stored as 4, K = 27): far below a UMMA K-block and bound by the
//    32-channel fp16 store, so it runs on CUDA cores with the 3x3x4 neighbourhood in registers and the weights
//    broadcast from shared memory.  Two pixels per thread so each weight read feeds two FMAs.
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "conv_params.h"

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
        const __half* w = p.w + (long long)(jc * 8) * p.ktot + (long long)t * p.cin;
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
// Persistent blocks (grid = a few per SM): the 36x32 weight table is converted to fp32 in shared memory ONCE per block,
// then the block strides over 64x4-pixel work items.  Each thread: 2 horizontally adjacent pixels x 32 channels.
__global__ void __launch_bounds__(128) conv_first_kernel(ConvParams p, int itemsX, int itemsY, int totalItems) {
    __shared__ float sw[36 * 32];  // [k][co], k = (ky*3+kx)*4 + ci
    __shared__ float sb[32];
    for (int i = threadIdx.x; i < 36 * 32; i += blockDim.x) {
        const int k = i >> 5, co = i & 31;
        sw[i] = __half2float(p.w[(long long)co * p.ktot + k]);
    }
    if (threadIdx.x < 32) sb[threadIdx.x] = p.bias[threadIdx.x];
    __syncthreads();
    for (int item = blockIdx.x; item < totalItems; item += gridDim.x) {
        const int ix = item % itemsX;
        const int iy = (item / itemsX) % itemsY;
        const int img = item / (itemsX * itemsY);
        const int x = (ix * 32 + (threadIdx.x & 31)) * 2;
        const int y = iy * 4 + (threadIdx.x >> 5);
        if (x >= p.gx || y >= p.gy) continue;
        const bool two = x + 1 < p.gx;
        // 3 rows x 4 columns x 3 channels neighbourhood (the 4th column only when the second pixel exists)
        float in[3][4][3];
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            const __half* row = p.in + (long long)img * p.sn + (long long)(y + ky) * p.sy + (long long)x * p.sx;
#pragma unroll
            for (int kx = 0; kx < 4; ++kx) {
                if (kx < 3 || two) {
                    const Half4 v = *reinterpret_cast<const Half4*>(row + (long long)kx * p.sx);
                    const float2 a = __half22float2(v.a), b = __half22float2(v.b);
                    in[ky][kx][0] = a.x; in[ky][kx][1] = a.y; in[ky][kx][2] = b.x;
                } else {
                    in[ky][kx][0] = in[ky][kx][1] = in[ky][kx][2] = 0.f;
                }
            }
        }
        float acc0[32], acc1[32];
#pragma unroll
        for (int co = 0; co < 32; ++co) { acc0[co] = sb[co]; acc1[co] = sb[co]; }
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx)
#pragma unroll
                for (int ci = 0; ci < 3; ++ci) {
                    const float a0 = in[ky][kx][ci], a1 = in[ky][kx + 1][ci];
                    const float4* wrow = reinterpret_cast<const float4*>(sw + ((ky * 3 + kx) * 4 + ci) * 32);
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const float4 wv = wrow[q];
                        acc0[4 * q + 0] = fmaf(a0, wv.x, acc0[4 * q + 0]); acc1[4 * q + 0] = fmaf(a1, wv.x, acc1[4 * q + 0]);
                        acc0[4 * q + 1] = fmaf(a0, wv.y, acc0[4 * q + 1]); acc1[4 * q + 1] = fmaf(a1, wv.y, acc1[4 * q + 1]);
                        acc0[4 * q + 2] = fmaf(a0, wv.z, acc0[4 * q + 2]); acc1[4 * q + 2] = fmaf(a1, wv.z, acc1[4 * q + 2]);
                        acc0[4 * q + 3] = fmaf(a0, wv.w, acc0[4 * q + 3]); acc1[4 * q + 3] = fmaf(a1, wv.w, acc1[4 * q + 3]);
                    }
                }
        __half* o = p.out + (((long long)img * p.out_h + y) * p.out_w + x) * p.out_c;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            Half8 h{__floats2half2_rn(lrelu(acc0[8 * q + 0], p.slope), lrelu(acc0[8 * q + 1], p.slope)),
                    __floats2half2_rn(lrelu(acc0[8 * q + 2], p.slope), lrelu(acc0[8 * q + 3], p.slope)),
                    __floats2half2_rn(lrelu(acc0[8 * q + 4], p.slope), lrelu(acc0[8 * q + 5], p.slope)),
                    __floats2half2_rn(lrelu(acc0[8 * q + 6], p.slope), lrelu(acc0[8 * q + 7], p.slope))};
            reinterpret_cast<Half8*>(o)[q] = h;
        }
        if (two) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                Half8 h{__floats2half2_rn(lrelu(acc1[8 * q + 0], p.slope), lrelu(acc1[8 * q + 1], p.slope)),
                        __floats2half2_rn(lrelu(acc1[8 * q + 2], p.slope), lrelu(acc1[8 * q + 3], p.slope)),
                        __floats2half2_rn(lrelu(acc1[8 * q + 4], p.slope), lrelu(acc1[8 * q + 5], p.slope)),
                        __floats2half2_rn(lrelu(acc1[8 * q + 6], p.slope), lrelu(acc1[8 * q + 7], p.slope))};
                reinterpret_cast<Half8*>(o + p.out_c)[q] = h;
            }
        }
    }
}

void launchConvFirst(const ConvParams& p, cudaStream_t s) {
    // contract: 3x3 taps in (ky,kx) order on a plain NHWC4 view, npad == 32, EPI_STORE with cout == 32
    const int itemsX = (p.gx + 63) / 64, itemsY = (p.gy + 3) / 4;
    const int total = itemsX * itemsY * p.gn;
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0) sms = 148;
    }
    const int grid = total < sms * 4 ? total : sms * 4;
    conv_first_kernel<<<grid, 128, 0, s>>>(p, itemsX, itemsY, total);
}

}  // namespace w2x
