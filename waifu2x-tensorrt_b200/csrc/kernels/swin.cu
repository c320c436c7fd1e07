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

// Each lane owns one 16-byte chunk (8 channels) of a token; a token occupies an aligned segment of SEG lanes (c = 96: 12 of 16
// lanes, two tokens per warp; c = 192: 24 of 32 lanes), so both reductions are log2(SEG)-step xor butterflies (idle lanes add 0).
template <int SEG>
__global__ void __launch_bounds__(256) layernorm_kernel(const __half* __restrict__ x, __half* __restrict__ y, long long tokens, int c,
                                                        const float* __restrict__ gamma, const float* __restrict__ beta, float eps) {
    constexpr int TPW = 32 / SEG;  // tokens per warp iteration (2 or 1)
    pdlLaunchDependents();
    const int lane = threadIdx.x & 31;
    const int sub = lane / SEG, li = lane - sub * SEG;
    const bool act = li * 8 < c;
    const long long warpId = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long warpCount = (long long)gridDim.x * (blockDim.x >> 5);
    float gm[8], bt[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { gm[i] = act ? gamma[li * 8 + i] : 0.f; bt[i] = act ? beta[li * 8 + i] : 0.f; }
    const float invc = 1.f / (float)c;
    pdlWait();  // gamma / beta are constants; x comes from the preceding kernel
    // two independent token groups per iteration: both 16-byte loads are in flight before either is reduced (one load per lane per
    // iteration left the kernel latency-bound at half the HBM rate)
    constexpr int U = 2;
    for (long long t0 = warpId * TPW * U; t0 < tokens; t0 += warpCount * TPW * U) {
        uint4 raw[U];
        bool ok[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long tok = t0 + u * TPW + sub;
            ok[u] = act && tok < tokens;
            raw[u] = make_uint4(0, 0, 0, 0);
            if (ok[u]) raw[u] = *reinterpret_cast<const uint4*>(x + tok * c + li * 8);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long tok = t0 + u * TPW + sub;
            float v[8];
            const __half2* h = reinterpret_cast<const __half2*>(&raw[u]);
#pragma unroll
            for (int i = 0; i < 4; ++i) { const float2 f = __half22float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
            float sum = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) sum += v[i];
#pragma unroll
            for (int o = SEG / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
            const float mean = sum * invc;
            float sq = 0.f;
            if (act) {
#pragma unroll
                for (int i = 0; i < 8; ++i) { const float d = v[i] - mean; sq += d * d; }
            }
#pragma unroll
            for (int o = SEG / 2; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
            const float rstd = rsqrtf(sq * invc + eps);
            if (ok[u]) {
                uint4 o;
                __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    oh[i] = __floats2half2_rn((v[2 * i] - mean) * rstd * gm[2 * i] + bt[2 * i], (v[2 * i + 1] - mean) * rstd * gm[2 * i + 1] + bt[2 * i + 1]);
                *reinterpret_cast<uint4*>(y + tok * c + li * 8) = o;
            }
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
    if (c == 96) launchPdl(layernorm_kernel<16>, dim3(grid), dim3(256), 0, s, x, y, tokens, c, gamma, beta, eps);
    else launchPdl(layernorm_kernel<32>, dim3(grid), dim3(256), 0, s, x, y, tokens, c, gamma, beta, eps);
}

// One warp per (window, head); window = 6 (36 tokens).  Both GEMMs of the attention run on warp-level mma.sync
// (m16n8k16, fp16 in / fp32 accumulate): S = Q K^T as 3 row tiles x 5 column tiles (36 -> 48 x 40, padding masked), the
// softmax works on the accumulator layout (quad shuffles), and the probabilities are re-used in registers as the A operand
// of O = P V (the accumulator layout of two adjacent 8-column tiles IS the A fragment of one k-step).
// The warp's Q, K and V slices (36 tokens x head dim each) are staged in shared memory with 16-byte cp.async copies (the
// roll + window partition is the gather's address math) and read back as MMA fragments with ldmatrix (.trans for V); rows are
// padded by 16 bytes so the eight row addresses of an 8x8 matrix fall into distinct banks.  The output tile goes back
// through the same buffer so that it leaves in 16-byte stores.
__device__ __forceinline__ void mma16816h(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t packHalf2(float a, float b) {
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
}
// 2^x, one MUFU (inputs here are <= 0 after the max subtraction; large negative inputs flush to 0)
__device__ __forceinline__ float ex2Approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void ldsm4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm2(uint32_t addr, uint32_t& r0, uint32_t& r1) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
__device__ __forceinline__ void ldsm4t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}

template <int HD>
__global__ void __launch_bounds__(128) window_attention_kernel(const __half* __restrict__ qkv, __half* __restrict__ out, int n, int h, int w, int c,
                                                               int heads, int shift, const float* __restrict__ relpos, long long totalUnits) {
    constexpr int WIN = 6, NT = 36, KS = HD / 16, DT = HD / 8, CH = HD / 8;
    constexpr int PITCH = HD * 2 + 16;   // bytes per staged row
    constexpr int MAT = 48 * PITCH;      // one of Q / K / V (48 rows: 36 tokens + zero padding up to the 3 x 16 row tiles)
    __shared__ int stok[4][48];   // token index of each (padded) window position; padding points at position 0
    __shared__ int sreg[4][48];   // shift-mask region id
    __shared__ __align__(16) uint8_t stage[4][3 * MAT];
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
    // ---- stage Q | K | V of this (window, head): 36 x 3 x CH chunks of 16 bytes, padding rows zeroed ----
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(&stage[wib][0]);
    const long long rowStride = 3ll * c;
    const __half* hbase = qkv + head * HD;
    for (int idx = lane; idx < NT * 3 * CH; idx += 32) {
        const int p = idx / (3 * CH), rem = idx - p * (3 * CH);
        const int which = rem / CH, ch = rem - which * CH;
        const __half* src = hbase + (long long)stok[wib][p] * rowStride + which * c + ch * 8;
        const uint32_t dst = sbase + (uint32_t)(which * MAT + p * PITCH + ch * 16);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    for (int idx = lane; idx < 12 * 3 * CH; idx += 32) {
        const int p = NT + idx / (3 * CH), rem = idx % (3 * CH);
        *reinterpret_cast<uint4*>(&stage[wib][(rem / CH) * MAT + p * PITCH + (rem % CH) * 16]) = make_uint4(0, 0, 0, 0);
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
    const uint32_t sQ = sbase, sK = sbase + MAT, sV = sbase + 2 * MAT;
    // ---- S = Q K^T ----
    float sacc[3][5][4];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 5; ++j) sacc[i][j][0] = sacc[i][j][1] = sacc[i][j][2] = sacc[i][j][3] = 0.f;
    uint32_t bk[5][KS][2];
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        // 8 key tokens x HD dims: matrix m of the ldmatrix = 16-byte chunk m of the rows -> (b0, b1) of k-step m / 2
        if (KS == 2) ldsm4(sK + (uint32_t)((8 * j + (lane & 7)) * PITCH + (lane >> 3) * 16), bk[j][0][0], bk[j][0][1], bk[j][KS - 1][0], bk[j][KS - 1][1]);
        else ldsm2(sK + (uint32_t)((8 * j + (lane & 7)) * PITCH + ((lane >> 3) & 1) * 16), bk[j][0][0], bk[j][0][1]);
    }
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            uint32_t a0, a1, a2, a3;
            ldsm4(sQ + (uint32_t)((16 * i + (lane & 7) + ((lane >> 3) & 1) * 8) * PITCH + (2 * ks + (lane >> 4)) * 16), a0, a1, a2, a3);
#pragma unroll
            for (int j = 0; j < 5; ++j) mma16816h(sacc[i][j], a0, a1, a2, a3, bk[j][ks][0], bk[j][ks][1]);
        }
    }
    // ---- scale + relative position bias + shift mask + softmax (rows 16i+g and 16i+g+8; columns 8j+2t, 8j+2t+1) ----
    // Everything is in log2 units: `relpos` arrives as [head][48][40], already multiplied by log2(e) and with the padded key columns
    // set to -1e30, so an element costs one FFMA, (the mask), a max, a subtract, an ex2 and an add.
    const float scale2 = rsqrtf((float)HD) * 1.4426950408889634f;
    const float* bias = relpos + (long long)head * 48 * 40 + 2 * t;
    int colReg[5][2];  // shift-mask region of this lane's ten key columns
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        colReg[j][0] = sreg[wib][min(8 * j + 2 * t, NT - 1)];
        colReg[j][1] = sreg[wib][min(8 * j + 2 * t + 1, NT - 1)];
    }
    uint32_t pa[3][5][2];  // probabilities as half2: [row tile][col tile][row g | row g+8]
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int r = 16 * i + g + 8 * half;
            const int myReg = sreg[wib][min(r, NT - 1)];
            const float* brow = bias + r * 40;
            float v[5][2];
            float mx = -1e30f;
#pragma unroll
            for (int j = 0; j < 5; ++j) {
                const float2 b = __ldg(reinterpret_cast<const float2*>(brow + 8 * j));
                float a0 = fmaf(sacc[i][j][2 * half], scale2, b.x), a1 = fmaf(sacc[i][j][2 * half + 1], scale2, b.y);
                if (shift > 0) {
                    if (colReg[j][0] != myReg) a0 -= 144.26950408889634f;  // -100 in natural-log units
                    if (colReg[j][1] != myReg) a1 -= 144.26950408889634f;
                }
                v[j][0] = a0;
                v[j][1] = a1;
                mx = fmaxf(mx, fmaxf(a0, a1));
            }
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
            float den = 0.f;
#pragma unroll
            for (int j = 0; j < 5; ++j) {
                v[j][0] = ex2Approx(v[j][0] - mx);
                v[j][1] = ex2Approx(v[j][1] - mx);
                den += v[j][0] + v[j][1];
            }
            den += __shfl_xor_sync(0xffffffffu, den, 1);
            den += __shfl_xor_sync(0xffffffffu, den, 2);
            const float inv = __fdividef(1.f, den);
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
#pragma unroll
    for (int s = 0; s < 3; ++s) {
#pragma unroll
        for (int jp = 0; jp < DT / 2; ++jp) {
            // V^T fragments of dims 16jp..16jp+15 for tokens 16s..16s+15 (padded tokens are zero rows and carry P = 0)
            uint32_t b00, b01, b10, b11;
            ldsm4t(sV + (uint32_t)((16 * s + (lane & 7) + ((lane >> 3) & 1) * 8) * PITCH + (2 * jp + (lane >> 4)) * 16), b00, b01, b10, b11);
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const uint32_t a0 = pa[i][2 * s][0], a1 = pa[i][2 * s][1];
                const uint32_t a2 = (2 * s + 1 < 5) ? pa[i][(2 * s + 1 < 5) ? 2 * s + 1 : 0][0] : 0u;
                const uint32_t a3 = (2 * s + 1 < 5) ? pa[i][(2 * s + 1 < 5) ? 2 * s + 1 : 0][1] : 0u;
                mma16816h(oacc[i][2 * jp], a0, a1, a2, a3, b00, b01);
                mma16816h(oacc[i][2 * jp + 1], a0, a1, a2, a3, b10, b11);
            }
        }
    }
    // ---- O -> the Q buffer (all Q reads are done) -> 16-byte stores ----
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int r = 16 * i + g + 8 * half;
#pragma unroll
            for (int jd = 0; jd < DT; ++jd)
                *reinterpret_cast<__half2*>(&stage[wib][r * PITCH + (8 * jd + 2 * t) * 2]) = __floats2half2_rn(oacc[i][jd][2 * half], oacc[i][jd][2 * half + 1]);
        }
    }
    __syncwarp();
    for (int idx = lane; idx < NT * CH; idx += 32) {
        const int p = idx / CH, ch = idx - p * CH;
        *reinterpret_cast<uint4*>(out + (long long)stok[wib][p] * c + head * HD + ch * 8) = *reinterpret_cast<const uint4*>(&stage[wib][p * PITCH + ch * 16]);
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
