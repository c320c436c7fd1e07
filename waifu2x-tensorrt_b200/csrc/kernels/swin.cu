// SwinUNet token kernels (arithmetic of torchvision's SwinTransformerBlock v1 / shifted_window_attention):
//   layernorm_kernel        y = (x - mean) / sqrt(var + eps) * gamma + beta over the channel dimension, one warp per token
//   window_attention_kernel softmax(q k^T / sqrt(d) + relative_position_bias [+ shift mask]) v for one (window, head) per
//                           warp.  The cyclic shift (torch.roll) and the window partition are pure index arithmetic on the
//                           NHWC token tensor: nothing is rolled or re-laid-out in memory.
// The four Linear layers of a block run on the tcgen05 implicit-GEMM kernel (1x1 "convolutions" over the token grid) with
// bias / GELU / residual fused in its epilogue.
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "conv_params.h"

namespace w2x {

__global__ void __launch_bounds__(256) layernorm_kernel(const __half* __restrict__ x, __half* __restrict__ y, long long tokens, int c,
                                                        const float* __restrict__ gamma, const float* __restrict__ beta, float eps) {
    const long long tok = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (tok >= tokens) return;
    const int lane = threadIdx.x & 31;
    const int pairs = c >> 1;  // c <= 256: at most 4 half2 per lane
    const __half2* xr = reinterpret_cast<const __half2*>(x + tok * c);
    float2 v[4];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int p = lane + 32 * i;
        v[i] = p < pairs ? __half22float2(xr[p]) : make_float2(0.f, 0.f);
        sum += v[i].x + v[i].y;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum / (float)c;
    float var = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int p = lane + 32 * i;
        if (p < pairs) {
            const float a = v[i].x - mean, b = v[i].y - mean;
            var += a * a + b * b;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) var += __shfl_xor_sync(0xffffffffu, var, o);
    const float rstd = rsqrtf(var / (float)c + eps);
    __half2* yr = reinterpret_cast<__half2*>(y + tok * c);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int p = lane + 32 * i;
        if (p < pairs) {
            const float a = (v[i].x - mean) * rstd * gamma[2 * p] + beta[2 * p];
            const float b = (v[i].y - mean) * rstd * gamma[2 * p + 1] + beta[2 * p + 1];
            yr[p] = __floats2half2_rn(a, b);
        }
    }
}

void launchLayerNorm(const __half* x, __half* y, long long tokens, int c, const float* gamma, const float* beta, float eps, cudaStream_t s) {
    const int warpsPerBlock = 8;
    const long long blocks = (tokens + warpsPerBlock - 1) / warpsPerBlock;
    layernorm_kernel<<<(unsigned)blocks, 32 * warpsPerBlock, 0, s>>>(x, y, tokens, c, gamma, beta, eps);
}

// One warp per (window, head).  HD = head dim (16 or 32), window = 6 (n = 36 tokens).
template <int HD>
__global__ void __launch_bounds__(128) window_attention_kernel(const __half* __restrict__ qkv, __half* __restrict__ out, int n, int h, int w, int c,
                                                               int heads, int shift, const float* __restrict__ relpos, long long totalUnits) {
    constexpr int WIN = 6, NT = 36;
    __shared__ float sk[4][NT][HD + 1];
    __shared__ float sv[4][NT][HD + 1];
    __shared__ int stok[4][NT];   // token index (img*h*w + y*w + x) of each window position
    __shared__ int sreg[4][NT];   // shift-mask region id
    const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long unit = (long long)blockIdx.x * 4 + wib;
    if (unit >= totalUnits) return;  // whole warp exits together
    const int head = (int)(unit % heads);
    long long win = unit / heads;
    const int nwx = w / WIN, nwy = h / WIN;
    const int wx = (int)(win % nwx);
    win /= nwx;
    const int wy = (int)(win % nwy);
    const int img = (int)(win / nwy);
    // window positions -> source tokens (torch.roll by -shift then partition) and mask regions (computed on shifted coords)
    for (int t = lane; t < NT; t += 32) {
        const int ys = wy * WIN + t / WIN, xs = wx * WIN + t % WIN;
        const int y = (ys + shift) % h, x = (xs + shift) % w;
        stok[wib][t] = (img * h + y) * w + x;
        const int hid = ys < h - WIN ? 0 : (ys < h - shift ? 1 : 2);
        const int wid = xs < w - WIN ? 0 : (xs < w - shift ? 1 : 2);
        sreg[wib][t] = hid * 3 + wid;
    }
    __syncwarp();
    // stage K and V of this head (fp32) -- 36 x HD each
    for (int i = lane; i < NT * HD; i += 32) {
        const int t = i / HD, d = i - t * HD;
        const __half* base = qkv + (long long)stok[wib][t] * (3 * c) + head * HD + d;
        sk[wib][t][d] = __half2float(base[c]);
        sv[wib][t][d] = __half2float(base[2 * c]);
    }
    __syncwarp();
    const float scale = rsqrtf((float)HD);
    const float* bias = relpos + (long long)head * NT * NT;
    for (int r = lane; r < NT; r += 32) {
        float q[HD];
        const __half* qp = qkv + (long long)stok[wib][r] * (3 * c) + head * HD;
#pragma unroll
        for (int d = 0; d < HD; ++d) q[d] = __half2float(qp[d]) * scale;
        float sc[NT];
        float mx = -1e30f;
        const int myReg = sreg[wib][r];
#pragma unroll
        for (int t = 0; t < NT; ++t) {
            float a = 0.f;
#pragma unroll
            for (int d = 0; d < HD; ++d) a = fmaf(q[d], sk[wib][t][d], a);
            a += bias[r * NT + t];
            if (shift > 0 && sreg[wib][t] != myReg) a += -100.f;
            sc[t] = a;
            mx = fmaxf(mx, a);
        }
        float den = 0.f;
#pragma unroll
        for (int t = 0; t < NT; ++t) { sc[t] = __expf(sc[t] - mx); den += sc[t]; }
        const float inv = 1.f / den;
        float o[HD];
#pragma unroll
        for (int d = 0; d < HD; ++d) o[d] = 0.f;
#pragma unroll
        for (int t = 0; t < NT; ++t) {
            const float pw = sc[t] * inv;
#pragma unroll
            for (int d = 0; d < HD; ++d) o[d] = fmaf(pw, sv[wib][t][d], o[d]);
        }
        __half* op = out + (long long)stok[wib][r] * c + head * HD;
#pragma unroll
        for (int d = 0; d < HD; d += 2) *reinterpret_cast<__half2*>(op + d) = __floats2half2_rn(o[d], o[d + 1]);
    }
}

void launchWindowAttention(const __half* qkv, __half* out, int n, int h, int w, int c, int heads, int window, int shift,
                           const float* relpos, cudaStream_t s) {
    // contract: window == 6, h % 6 == 0, w % 6 == 0, head dim in {16, 32}
    const long long units = (long long)n * (h / window) * (w / window) * heads;
    const unsigned blocks = (unsigned)((units + 3) / 4);
    const int hd = c / heads;
    // torchvision drops the shift when the window covers the whole extent (shifted_window_attention :159-163)
    const int sh = (window >= h || window >= w) ? 0 : shift;
    if (hd == 16) window_attention_kernel<16><<<blocks, 128, 0, s>>>(qkv, out, n, h, w, c, heads, sh, relpos, units);
    else window_attention_kernel<32><<<blocks, 128, 0, s>>>(qkv, out, n, h, w, c, heads, sh, relpos, units);
}

}  // namespace w2x
