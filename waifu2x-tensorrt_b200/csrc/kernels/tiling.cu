// Memory-bound tiling kernels: the fused replacements for the reference's cv::cuda library launches in
// /root/reference/src/tensorrt/img2img_render.cpp and img2img_infer.cpp.
//
//   unpack     : padRoi (:68-105) + applyAugmentation (:134-177) + cvtColor BGR2RGB (:227) + split/convertTo(1/255)
//                (infer.cpp:5-21) -> fp16 RGB0 NHWC4 tile batch, one pass, no intermediate mats.
//   stitch     : applyWeights (:107-121) + canvas add (:329-330) + convertTo(8U, 255) (:342) + RGB2BGR (:343).
//                Gather formulation: each output pixel is owned by exactly one thread, which sums the <= 4 covering
//                tiles in the reference's tile order (column-major), so no zero-filled f32 canvas and no RMW passes.
//   tta_reduce : reverseAugmentation (:179-222) + accumulate + x(1/8) (:305-318).
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "conv_params.h"

namespace w2x {

// augmented[r][c] == original[rr][cc] for an n x n tile (D4 ops by OpenCV flip code / NPP CCW rotation; see
// oracle/tiling.py::augment_src_index which tests pin against np.rot90 / slicing).
__device__ __forceinline__ void augSrcIndex(int k, int r, int c, int n, int& rr, int& cc) {
    const int m = n - 1;
    switch (k) {
        default: rr = r; cc = c; break;
        case 1: rr = m - r; cc = c; break;
        case 2: rr = r; cc = m - c; break;
        case 3: rr = c; cc = m - r; break;
        case 4: rr = m - r; cc = m - c; break;
        case 5: rr = m - c; cc = r; break;
        case 6: rr = m - c; cc = m - r; break;
        case 7: rr = c; cc = r; break;
    }
}

// deaugmented[r][c] == model_out[rr][cc]
__device__ __forceinline__ void revSrcIndex(int k, int r, int c, int n, int& rr, int& cc) {
    const int m = n - 1;
    switch (k) {
        default: rr = r; cc = c; break;
        case 1: rr = m - r; cc = c; break;
        case 2: rr = r; cc = m - c; break;
        case 3: rr = m - c; cc = r; break;
        case 4: rr = m - r; cc = m - c; break;
        case 5: rr = c; cc = m - r; break;
        case 6: rr = m - c; cc = m - r; break;
        case 7: rr = c; cc = r; break;
    }
}

// ---- unpack ---------------------------------------------------------------------------------------
// grid: (ceil(tile*tile / 256), nslots).  One thread per tile pixel: 3 byte loads (L1/L2-served; neighbouring
// threads read neighbouring bytes), one 8-byte store.  Stores are fully coalesced (2 KB per warp-row).
__global__ void __launch_bounds__(256) unpack_kernel(const uint8_t* __restrict__ frame, int w, int h, size_t pitch,
                                                     const TileSlot* __restrict__ slots, int tile,
                                                     __half* __restrict__ out) {
    const int slot = blockIdx.y;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= tile * tile) return;
    const TileSlot ts = slots[slot];
    Half4 o;
    if (!ts.valid) {
        o.a = __floats2half2_rn(0.f, 0.f);
        o.b = o.a;
    } else {
        const int r = idx / tile, c = idx - r * tile;
        int rr, cc;
        augSrcIndex(ts.aug, r, c, tile, rr, cc);
        const int sy = min(max(ts.y + rr, 0), h - 1);
        const int sx = min(max(ts.x + cc, 0), w - 1);
        const uint8_t* px = frame + (size_t)sy * pitch + (size_t)sx * 3;
        const float k = 1.0f / 255.0f;  // float(1.0/255.0), infer.cpp:19
        const float b = __fmul_rn((float)px[0], k), g = __fmul_rn((float)px[1], k), rch = __fmul_rn((float)px[2], k);
        o.a = __floats2half2_rn(rch, g);
        o.b = __floats2half2_rn(b, 0.f);
    }
    reinterpret_cast<Half4*>(out)[(size_t)slot * tile * tile + idx] = o;
}

void launchUnpack(const uint8_t* frame, int w, int h, size_t pitch, const TileSlot* slots, int nslots, int tile,
                  __half* out, cudaStream_t s) {
    if (nslots <= 0) return;
    dim3 grid((tile * tile + 255) / 256, nslots);
    unpack_kernel<<<grid, 256, 0, s>>>(frame, w, h, pitch, slots, tile, out);
}

// ---- stitch + pack ----------------------------------------------------------------------------------
template <bool F32>
__device__ __forceinline__ void loadTilePx(const void* tiles, size_t idx, float& r, float& g, float& b) {
    if (F32) {
        const float4 v = reinterpret_cast<const float4*>(tiles)[idx];
        r = v.x; g = v.y; b = v.z;
    } else {
        const Half4 v = reinterpret_cast<const Half4*>(tiles)[idx];
        const float2 p0 = __half22float2(v.a), p1 = __half22float2(v.b);
        r = p0.x; g = p0.y; b = p1.x;
    }
}

// One thread per output pixel.  Along each axis a pixel is covered by at most two tiles (stride = outT - ov > outT/2).
template <bool F32>
__global__ void __launch_bounds__(256) stitch_kernel(StitchParams p) {
    const int ox = blockIdx.x * 64 + (threadIdx.x & 63);
    const int oy = p.y_begin + blockIdx.y * 4 + (threadIdx.x >> 6);
    if (ox >= p.cw || oy >= p.y_end) return;
    const int stx = p.outT - p.ovx, sty = p.outT - p.ovy;
    int ti[2], tj[2], ni = 0, nj = 0;
    {
        const int i1 = min(ox / stx, p.nx - 1);
        if (i1 > 0 && ox - (i1 - 1) * stx < p.outT) ti[ni++] = i1 - 1;
        if (ox - i1 * stx < p.outT) ti[ni++] = i1;
        const int j1 = min(oy / sty, p.ny - 1);
        if (j1 > 0 && oy - (j1 - 1) * sty < p.outT) tj[nj++] = j1 - 1;
        if (oy - j1 * sty < p.outT) tj[nj++] = j1;
    }
    float accr = 0.f, accg = 0.f, accb = 0.f;
    for (int a = 0; a < ni; ++a) {
        const int i = ti[a];
        const int tx0 = i * stx;
        const int lx = ox - tx0;
        const int tw = tx0 + p.outT > p.cw ? p.cw - tx0 : p.outT;  // clipped rect width (render.cpp:59)
        // applyWeights predicates on the clipped rect (render.cpp:110-120); ramps per img2img_load.cpp:29-52
        float wl = 1.f, wr = 1.f;
        if (tx0 > 0 && lx < p.ovx) wl = p.rampx[lx];
        if (tx0 + tw < p.cw && p.outT - 1 - lx < p.ovx) wr = p.rampx[p.outT - 1 - lx];
        for (int bsel = 0; bsel < nj; ++bsel) {
            const int j = tj[bsel];
            const int ty0 = j * sty;
            const int ly = oy - ty0;
            const int th = ty0 + p.outT > p.ch ? p.ch - ty0 : p.outT;
            float wt = 1.f, wb = 1.f;
            if (ty0 > 0 && ly < p.ovy) wt = p.rampy[ly];
            if (ty0 + th < p.ch && p.outT - 1 - ly < p.ovy) wb = p.rampy[p.outT - 1 - ly];
            float r, g, b;
            const int slot = p.tile_map ? __ldg(p.tile_map + i * p.ny + j) : i * p.ny + j;
            loadTilePx<F32>(p.tiles, ((size_t)slot * p.outT + ly) * p.outT + lx, r, g, b);
            // sequential in-place multiplies in the reference order: left, top, right, bottom (x1.0 is exact)
            r = __fmul_rn(__fmul_rn(__fmul_rn(__fmul_rn(r, wl), wt), wr), wb);
            g = __fmul_rn(__fmul_rn(__fmul_rn(__fmul_rn(g, wl), wt), wr), wb);
            b = __fmul_rn(__fmul_rn(__fmul_rn(__fmul_rn(b, wl), wt), wr), wb);
            accr = __fadd_rn(accr, r);
            accg = __fadd_rn(accg, g);
            accb = __fadd_rn(accb, b);
        }
    }
    // convertTo(CV_8UC3, 255): saturate_cast<uchar>(rint(v * 255)) (render.cpp:342), then RGB -> BGR (:343)
    const int ir = min(max(__float2int_rn(__fmul_rn(accr, 255.f)), 0), 255);
    const int ig = min(max(__float2int_rn(__fmul_rn(accg, 255.f)), 0), 255);
    const int ib = min(max(__float2int_rn(__fmul_rn(accb, 255.f)), 0), 255);
    uint8_t* d = p.dst + (size_t)oy * p.pitch + (size_t)ox * 3;
    d[0] = (uint8_t)ib;
    d[1] = (uint8_t)ig;
    d[2] = (uint8_t)ir;
}

void launchStitch(const StitchParams& pin, cudaStream_t s) {
    StitchParams p = pin;
    if (p.y_end <= 0) { p.y_begin = 0; p.y_end = p.ch; }  // whole canvas
    if (p.y_end <= p.y_begin) return;
    dim3 grid((p.cw + 63) / 64, (p.y_end - p.y_begin + 3) / 4);
    if (p.f32) stitch_kernel<true><<<grid, 256, 0, s>>>(p);
    else stitch_kernel<false><<<grid, 256, 0, s>>>(p);
}

// ---- TTA reduce -------------------------------------------------------------------------------------
// acc = 0 + out_0; acc += rev_k(out_k), k = 1..7 (index order); acc *= 1/8   (render.cpp:305-318, mean per SURVEY q1)
__global__ void __launch_bounds__(256) tta_reduce_kernel(const __half* __restrict__ outs, int outT, float* __restrict__ mean) {
    const int tile = blockIdx.y;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= outT * outT) return;
    const int r = idx / outT, c = idx - r * outT;
    float ar = 0.f, ag = 0.f, ab = 0.f;
    const size_t plane = (size_t)outT * outT;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        int rr, cc;
        revSrcIndex(k, r, c, outT, rr, cc);
        const Half4 v = reinterpret_cast<const Half4*>(outs)[((size_t)tile * 8 + k) * plane + (size_t)rr * outT + cc];
        const float2 p0 = __half22float2(v.a), p1 = __half22float2(v.b);
        ar = __fadd_rn(ar, p0.x);
        ag = __fadd_rn(ag, p0.y);
        ab = __fadd_rn(ab, p1.x);
    }
    const float k8 = 0.125f;
    reinterpret_cast<float4*>(mean)[(size_t)tile * plane + idx] =
        make_float4(__fmul_rn(ar, k8), __fmul_rn(ag, k8), __fmul_rn(ab, k8), 0.f);
}

void launchTtaReduce(const __half* outs, int tiles, int outT, float* mean, cudaStream_t s) {
    if (tiles <= 0) return;
    dim3 grid((outT * outT + 255) / 256, tiles);
    tta_reduce_kernel<<<grid, 256, 0, s>>>(outs, outT, mean);
}

// ---- layout helpers for the Img2Img::infer-shaped entry point ---------------------------------------------
__global__ void nchw_to_nhwc4_kernel(const float* __restrict__ in, int t, __half* __restrict__ out) {
    const int n = blockIdx.y;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= t * t) return;
    const size_t plane = (size_t)t * t;
    const float* b = in + (size_t)n * 3 * plane;
    Half4 o{__floats2half2_rn(b[idx], b[plane + idx]), __floats2half2_rn(b[2 * plane + idx], 0.f)};
    reinterpret_cast<Half4*>(out)[(size_t)n * plane + idx] = o;
}

__global__ void nhwc4_to_nchw_kernel(const __half* __restrict__ in, int t, float* __restrict__ out) {
    const int n = blockIdx.y;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= t * t) return;
    const size_t plane = (size_t)t * t;
    const Half4 v = reinterpret_cast<const Half4*>(in)[(size_t)n * plane + idx];
    const float2 p0 = __half22float2(v.a), p1 = __half22float2(v.b);
    float* b = out + (size_t)n * 3 * plane;
    b[idx] = p0.x;
    b[plane + idx] = p0.y;
    b[2 * plane + idx] = p1.x;
}

void launchNchwToNhwc4(const float* in, int n, int t, __half* out, cudaStream_t s) {
    dim3 grid((t * t + 255) / 256, n);
    nchw_to_nhwc4_kernel<<<grid, 256, 0, s>>>(in, t, out);
}

void launchNhwc4ToNchw(const __half* in, int n, int t, float* out, cudaStream_t s) {
    dim3 grid((t * t + 255) / 256, n);
    nhwc4_to_nchw_kernel<<<grid, 256, 0, s>>>(in, t, out);
}

}  // namespace w2x
