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
#include "launch.h"

namespace w2x {

// Each lane owns one 16-byte chunk (8 channels) of a token: a warp covers 32 / (c/8) tokens per iteration (c = 96 -> 2 full
// tokens + idle lanes, c = 192 -> 1 token); the reduction runs over the c/8 lanes of a token with segmented shuffles.
template <int LANES>  // lanes per token = c / 8: 12 (c = 96) or 24 (c = 192)
__global__ void __launch_bounds__(256) layernorm_kernel(const __half* __restrict__ x, __half* __restrict__ y, long long tokens, int c,
                                                        const float* __restrict__ gamma, const float* __restrict__ beta, float eps) {
    constexpr int TPW = 32 / LANES;  // tokens per warp iteration (2 or 1)
    pdlLaunchDependents();
    const int lane = threadIdx.x & 31;
    const int sub = lane / LANES, li = lane - sub * LANES;
    const bool act = sub < TPW;
    const long long warpId = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long warpCount = (long long)gridDim.x * (blockDim.x >> 5);
    float gm[8], bt[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { gm[i] = act ? gamma[li * 8 + i] : 0.f; bt[i] = act ? beta[li * 8 + i] : 0.f; }
    const float invc = 1.f / (float)c;
    pdlWait();  // gamma / beta are constants; x comes from the preceding kernel
    for (long long t0 = warpId * TPW; t0 < tokens; t0 += warpCount * TPW) {
        const long long tok = t0 + sub;
        const bool ok = act && tok < tokens;
        float v[8];
        if (ok) {
            const uint4 raw = *reinterpret_cast<const uint4*>(x + tok * c + li * 8);
            const __half2* h = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
            for (int i = 0; i < 4; ++i) { const float2 f = __half22float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = 0.f;
        }
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) sum += v[i];
        // all-lanes butterfly over the token's LANES lanes: gather via shuffles from the token's first lane range
        float tot = 0.f;
#pragma unroll
        for (int j = 0; j < LANES; ++j) tot += __shfl_sync(0xffffffffu, sum, sub * LANES + j);
        const float mean = tot * invc;
        float sq = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) { const float d = v[i] - mean; sq += d * d; }
        float vt = 0.f;
#pragma unroll
        for (int j = 0; j < LANES; ++j) vt += __shfl_sync(0xffffffffu, sq, sub * LANES + j);
        const float rstd = rsqrtf(vt * invc + eps);
        if (ok) {
            uint4 o;
            __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
            for (int i = 0; i < 4; ++i)
                oh[i] = __floats2half2_rn((v[2 * i] - mean) * rstd * gm[2 * i] + bt[2 * i], (v[2 * i + 1] - mean) * rstd * gm[2 * i + 1] + bt[2 * i + 1]);
            *reinterpret_cast<uint4*>(y + tok * c + li * 8) = o;
        }
    }
}

void launchLayerNorm(const __half* x, __half* y, long long tokens, int c, const float* gamma, const float* beta, float eps, cudaStream_t s) {
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0) sms = 148;
    }
    const int grid = sms * 8;
    // contract: c in {96, 192}
    if (c == 96) launchPdl(layernorm_kernel<12>, dim3(grid), dim3(256), 0, s, x, y, tokens, c, gamma, beta, eps);
    else launchPdl(layernorm_kernel<24>, dim3(grid), dim3(256), 0, s, x, y, tokens, c, gamma, beta, eps);
}

// One warp per (window, head); window = 6 (36 tokens).  Both GEMMs of the attention run on warp-level mma.sync
// (m16n8k16, fp16 in / fp32 accumulate): S = Q K^T as 3 row tiles x 5 column tiles (36 -> 48 x 40, padding masked), the
// softmax works on the accumulator layout (quad shuffles), and the probabilities are re-used in registers as the A operand
// of O = P V (the accumulator layout of two adjacent 8-column tiles IS the A fragment of one k-step).  Q/K/V fragments are
// loaded straight from the [token][3C] tensor (4-byte loads, L1-resident): no shared-memory staging, no re-layout.
__device__ __forceinline__ void mma16816h(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t packHalf2(float a, float b) {
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
}

template <int HD>
__global__ void __launch_bounds__(128) window_attention_kernel(const __half* __restrict__ qkv, __half* __restrict__ out, int n, int h, int w, int c,
                                                               int heads, int shift, const float* __restrict__ relpos, long long totalUnits) {
    constexpr int WIN = 6, NT = 36, KS = HD / 16, DT = HD / 8;
    __shared__ int stok[4][48];   // token index of each (padded) window position; padding points at position 0
    __shared__ int sreg[4][48];   // shift-mask region id
    pdlLaunchDependents();
    pdlWait();
    const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const long long unit = (long long)blockIdx.x * 4 + wib;
    if (unit >= totalUnits) return;  // whole warp exits together
    const int head = (int)(unit % heads);
    long long win = unit / heads;
    const int nwx = w / WIN, nwy = h / WIN;
    const int wx = (int)(win % nwx);
    win /= nwx;
    const int wy = (int)(win % nwy);
    const int img = (int)(win / nwy);
    for (int p = lane; p < 48; p += 32) {
        const int pp = p < NT ? p : 0;
        const int ys = wy * WIN + pp / WIN, xs = wx * WIN + pp % WIN;     // coordinates in the rolled grid
        const int y = (ys + shift) % h, x = (xs + shift) % w;             // torch.roll(x, -shift): rolled[p] = x[p + shift]
        stok[wib][p] = (img * h + y) * w + x;
        const int hid = ys < h - WIN ? 0 : (ys < h - shift ? 1 : 2);
        const int wid = xs < w - WIN ? 0 : (xs < w - shift ? 1 : 2);
        sreg[wib][p] = hid * 3 + wid;
    }
    __syncwarp();
    const long long rowStride = 3ll * c;
    const __half* qbase = qkv + head * HD;
    // ---- S = Q K^T ----
    float sacc[3][5][4];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 5; ++j) sacc[i][j][0] = sacc[i][j][1] = sacc[i][j][2] = sacc[i][j][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
        uint32_t bk[5][2];
#pragma unroll
        for (int j = 0; j < 5; ++j) {
            const __half* kp = qbase + (long long)stok[wib][8 * j + g] * rowStride + c + 16 * ks + 2 * t;
            bk[j][0] = *reinterpret_cast<const uint32_t*>(kp);
            bk[j][1] = *reinterpret_cast<const uint32_t*>(kp + 8);
        }
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const __half* q0 = qbase + (long long)stok[wib][16 * i + g] * rowStride + 16 * ks + 2 * t;
            const __half* q1 = qbase + (long long)stok[wib][16 * i + g + 8] * rowStride + 16 * ks + 2 * t;
            const uint32_t a0 = *reinterpret_cast<const uint32_t*>(q0), a1 = *reinterpret_cast<const uint32_t*>(q1);
            const uint32_t a2 = *reinterpret_cast<const uint32_t*>(q0 + 8), a3 = *reinterpret_cast<const uint32_t*>(q1 + 8);
#pragma unroll
            for (int j = 0; j < 5; ++j) mma16816h(sacc[i][j], a0, a1, a2, a3, bk[j][0], bk[j][1]);
        }
    }
    // ---- scale + relative position bias + shift mask + softmax (rows 16i+g and 16i+g+8; columns 8j+2t, 8j+2t+1) ----
    const float scale = rsqrtf((float)HD);
    const float* bias = relpos + (long long)head * NT * NT;
    uint32_t pa[3][5][2];  // probabilities as half2: [row tile][col tile][row g | row g+8]
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int r = 16 * i + g + 8 * half;
            const bool rowOk = r < NT;
            const int rr = rowOk ? r : 0;
            const int myReg = sreg[wib][rr];
            float v[5][2];
            float mx = -1e30f;
#pragma unroll
            for (int j = 0; j < 5; ++j) {
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int col = 8 * j + 2 * t + e;
                    float a = -1e30f;
                    if (col < NT) {
                        a = sacc[i][j][2 * half + e] * scale + __ldg(bias + rr * NT + col);
                        if (shift > 0 && sreg[wib][col] != myReg) a += -100.f;
                    }
                    v[j][e] = a;
                    mx = fmaxf(mx, a);
                }
            }
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
            float den = 0.f;
#pragma unroll
            for (int j = 0; j < 5; ++j) {
                v[j][0] = __expf(v[j][0] - mx);
                v[j][1] = __expf(v[j][1] - mx);
                den += v[j][0] + v[j][1];
            }
            den += __shfl_xor_sync(0xffffffffu, den, 1);
            den += __shfl_xor_sync(0xffffffffu, den, 2);
            const float inv = 1.f / den;
#pragma unroll
            for (int j = 0; j < 5; ++j) pa[i][j][half] = packHalf2(v[j][0] * inv, v[j][1] * inv);
        }
    }
    // ---- O = P V : k-steps over 16 tokens (column tiles 2s, 2s+1), n-tiles over the head dim ----
    float oacc[3][DT][4];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int jd = 0; jd < DT; ++jd) oacc[i][jd][0] = oacc[i][jd][1] = oacc[i][jd][2] = oacc[i][jd][3] = 0.f;
    const unsigned short* vbase = reinterpret_cast<const unsigned short*>(qbase + 2 * c);
#pragma unroll
    for (int s = 0; s < 3; ++s) {
        // V fragments: b0 = (tokens 16s+2t, +1 ; dim 8jd+g), b1 = (tokens 16s+2t+8, +9); padded tokens carry P = 0
        const long long r0 = (long long)stok[wib][16 * s + 2 * t] * rowStride, r1 = (long long)stok[wib][16 * s + 2 * t + 1] * rowStride;
        const long long r2 = (long long)stok[wib][min(16 * s + 2 * t + 8, 47)] * rowStride, r3 = (long long)stok[wib][min(16 * s + 2 * t + 9, 47)] * rowStride;
#pragma unroll
        for (int jd = 0; jd < DT; ++jd) {
            const int d = 8 * jd + g;
            const uint32_t b0 = (uint32_t)vbase[r0 + d] | ((uint32_t)vbase[r1 + d] << 16);
            const uint32_t b1 = (uint32_t)vbase[r2 + d] | ((uint32_t)vbase[r3 + d] << 16);
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const uint32_t a0 = pa[i][2 * s][0], a1 = pa[i][2 * s][1];
                const uint32_t a2 = (2 * s + 1 < 5) ? pa[i][(2 * s + 1 < 5) ? 2 * s + 1 : 0][0] : 0u;
                const uint32_t a3 = (2 * s + 1 < 5) ? pa[i][(2 * s + 1 < 5) ? 2 * s + 1 : 0][1] : 0u;
                mma16816h(oacc[i][jd], a0, a1, a2, a3, b0, b1);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int r = 16 * i + g + 8 * half;
            if (r < NT) {
                __half* op = out + (long long)stok[wib][r] * c + head * HD + 2 * t;
#pragma unroll
                for (int jd = 0; jd < DT; ++jd)
                    *reinterpret_cast<__half2*>(op + 8 * jd) = __floats2half2_rn(oacc[i][jd][2 * half], oacc[i][jd][2 * half + 1]);
            }
        }
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
    if (hd == 16) launchPdl(window_attention_kernel<16>, dim3(blocks), dim3(128), 0, s, qkv, out, n, h, w, c, heads, sh, relpos, units);
    else launchPdl(window_attention_kernel<32>, dim3(blocks), dim3(128), 0, s, qkv, out, n, h, w, c, heads, sh, relpos, units);
}

}  // namespace w2x
