// Squeeze/excite (SEBlock of the CUNet family): per-tile global mean -> FC -> ReLU -> FC -> sigmoid ->
// per-(tile, channel) scale.  Deterministic two-stage reduction (no float atomics): the result must not depend on
// scheduling, because tile outputs are compared bit-for-bit between 1-GPU and N-GPU runs.
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "conv_params.h"

namespace w2x {

// grid (nblk, n), 256 threads.  Thread t owns channel pair (t % (c/2)) of pixels p = t / (c/2) + k * (256 / (c/2)).
__global__ void __launch_bounds__(256) se_squeeze_kernel(const __half* __restrict__ x, int hw, int c,
                                                         float* __restrict__ partial, int nblk) {
    extern __shared__ float red[];  // [256][2]
    const int img = blockIdx.y, blk = blockIdx.x;
    const int cp = c >> 1;
    const int lane_c = threadIdx.x % cp, lane_p = threadIdx.x / cp, pstride = 256 / cp;
    const int per = (hw + nblk - 1) / nblk;
    const int p0 = blk * per, p1 = min(hw, p0 + per);
    const __half2* base = reinterpret_cast<const __half2*>(x + (size_t)img * hw * c);
    float s0 = 0.f, s1 = 0.f;
    for (int p = p0 + lane_p; p < p1; p += pstride) {
        const float2 v = __half22float2(base[(size_t)p * cp + lane_c]);
        s0 += v.x;
        s1 += v.y;
    }
    red[threadIdx.x * 2] = s0;
    red[threadIdx.x * 2 + 1] = s1;
    __syncthreads();
    if (threadIdx.x < cp) {
        float a = 0.f, b = 0.f;
        for (int k = 0; k < pstride; ++k) {
            a += red[(k * cp + threadIdx.x) * 2];
            b += red[(k * cp + threadIdx.x) * 2 + 1];
        }
        float* o = partial + ((size_t)img * nblk + blk) * c + 2 * threadIdx.x;
        o[0] = a;
        o[1] = b;
    }
}

void launchSeSqueeze(const __half* x, int n, int h, int w, int c, float* partial, int nblk, cudaStream_t s) {
    dim3 grid(nblk, n);
    se_squeeze_kernel<<<grid, 256, 256 * 2 * sizeof(float), s>>>(x, h * w, c, partial, nblk);
}

// grid n, block 256.  mean -> relu(W1 mean + b1) -> sigmoid(W2 h + b2)
__global__ void __launch_bounds__(256) se_excite_kernel(const float* __restrict__ partial, int nblk, int c, int r, float inv_hw,
                                                        const float* __restrict__ w1, const float* __restrict__ b1,
                                                        const float* __restrict__ w2, const float* __restrict__ b2,
                                                        float* __restrict__ scale) {
    __shared__ float mean[256];
    __shared__ float hid[64];
    __shared__ float sw1[2048], sw2[2048];  // FC weights (r * c <= 2048 for c <= 128, r = c / 8); larger blocks read global
    const int img = blockIdx.x;
    const int fcElems = r * c;
    const bool fcInSmem = fcElems <= 2048;
    if (fcInSmem)
        for (int i = threadIdx.x; i < fcElems; i += blockDim.x) { sw1[i] = w1[i]; sw2[i] = w2[i]; }
    {
        // all 256 threads reduce the partial sums: (256 / c) slot groups per channel, fixed combination order
        __shared__ float part[256];
        const int groups = 256 / c, ch = threadIdx.x % c, g = threadIdx.x / c;
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
        const float* pb = partial + (size_t)img * nblk * c + ch;
        int b = g;
        for (; b + 3 * groups < nblk; b += 4 * groups) {
            s0 += pb[(size_t)b * c];
            s1 += pb[(size_t)(b + groups) * c];
            s2 += pb[(size_t)(b + 2 * groups) * c];
            s3 += pb[(size_t)(b + 3 * groups) * c];
        }
        for (; b < nblk; b += groups) s0 += pb[(size_t)b * c];
        part[threadIdx.x] = (s0 + s1) + (s2 + s3);
        __syncthreads();
        if (threadIdx.x < c) {
            float s = 0.f;
            for (int q = 0; q < groups; ++q) s += part[q * c + threadIdx.x];
            mean[threadIdx.x] = s * inv_hw;
        }
    }
    __syncthreads();
    const float* W1 = fcInSmem ? sw1 : w1;
    const float* W2 = fcInSmem ? sw2 : w2;
    {
        // hid[j] = relu(b1[j] + <W1[j], mean>): one warp per j, lanes stride the channels
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        for (int j = warp; j < r; j += (blockDim.x >> 5)) {
            float s = 0.f;
            for (int ch = lane; ch < c; ch += 32) s = fmaf(W1[j * c + ch], mean[ch], s);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (lane == 0) hid[j] = fmaxf(s + b1[j], 0.f);
        }
    }
    __syncthreads();
    for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
        float s = b2[ch];
        for (int j = 0; j < r; ++j) s = fmaf(W2[ch * r + j], hid[j], s);
        scale[(size_t)img * c + ch] = 1.f / (1.f + expf(-s));
    }
}

void launchSeExcite(const float* partial, int nblk, int n, int c, int r, int hw, const float* w1, const float* b1,
                    const float* w2, const float* b2, float* scale, cudaStream_t s) {
    se_excite_kernel<<<n, 256, 0, s>>>(partial, nblk, c, r, 1.0f / (float)hw, w1, b1, w2, b2, scale);
}

// in-place x *= scale[img][ch], 8 channels (16 B) per thread
__global__ void __launch_bounds__(256) se_scale_kernel(__half* __restrict__ x, long long vec_per_img, int c,
                                                       const float* __restrict__ scale) {
    const int img = blockIdx.y;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= vec_per_img) return;
    const int c0 = (int)((i * 8) % c);
    const float* sc = scale + (size_t)img * c + c0;
    Half8* px = reinterpret_cast<Half8*>(x) + (size_t)img * vec_per_img + i;
    Half8 v = *px;
    float2 t;
    t = __half22float2(v.a); v.a = __floats2half2_rn(t.x * sc[0], t.y * sc[1]);
    t = __half22float2(v.b); v.b = __floats2half2_rn(t.x * sc[2], t.y * sc[3]);
    t = __half22float2(v.c); v.c = __floats2half2_rn(t.x * sc[4], t.y * sc[5]);
    t = __half22float2(v.d); v.d = __floats2half2_rn(t.x * sc[6], t.y * sc[7]);
    *px = v;
}

__global__ void __launch_bounds__(256) scale_weights_kernel(const __half* __restrict__ w, __half* __restrict__ wOut, const float* __restrict__ scale,
                                                            int total, int ktot, int cin) {
    const int img = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;  // pair index
    if (2 * i >= total) return;
    const int k = (2 * i) % ktot;
    const float2 v = __half22float2(reinterpret_cast<const __half2*>(w)[i]);
    const float* sc = scale + (size_t)img * cin + (k % cin);
    reinterpret_cast<__half2*>(wOut + (size_t)img * total)[i] = __floats2half2_rn(v.x * sc[0], v.y * sc[1]);
}

void launchScaleWeights(const __half* w, __half* wOut, const float* scale, int nimg, int npad, int ktot, int cin, cudaStream_t s) {
    const int total = npad * ktot;  // even; k and k+1 share a tap because cin is even
    dim3 grid((total / 2 + 255) / 256, nimg);
    scale_weights_kernel<<<grid, 256, 0, s>>>(w, wOut, scale, total, ktot, cin);
}

void launchSeScale(__half* x, int n, int h, int w, int c, const float* scale, cudaStream_t s) {
    const long long vec = (long long)h * w * c / 8;
    dim3 grid((unsigned)((vec + 255) / 256), n);
    se_scale_kernel<<<grid, 256, 0, s>>>(x, vec, c, scale);
}

}  // namespace w2x
