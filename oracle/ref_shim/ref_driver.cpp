// TEST INFRASTRUCTURE ONLY -- C entry points of oracle/_ref/libw2xref.so (recipe: oracle/Makefile).
//
// The library is the reference's own src/tensorrt/{img2img_base,img2img_build,img2img_infer,img2img_load,img2img_render,
// logger}.cpp, compiled unmodified from /root/reference against the CPU mocks in this directory.  This file only exposes
// those translation units' functions with plain-C signatures so that tests/ and tests/golden/make_ref_goldens.py can call
// them through ctypes.  Every function below forwards to reference code; none restates it.
#include <cstring>
#include <memory>
#include <string>
#include <tuple>
#include <vector>

#include "tensorrt/img2img.h"

// free functions defined (non-static) in the reference's translation units
std::tuple<const int, std::vector<cv::Rect2i>, std::vector<cv::Rect2i>> calculateTiles(const cv::Rect2i& inputRect, const cv::Rect2i& outputRect,
                                                                                         const cv::Size2i& inputTileSize, const cv::Size2i& outputTileSize,
                                                                                         int scaling, const cv::Point2d& overlap);              // img2img_render.cpp:7
cv::cuda::GpuMat padRoi(const cv::cuda::GpuMat& input, const cv::Rect2i& roi, cv::cuda::Stream& stream);                                      // :68
void applyWeights(const cv::cuda::GpuMat& src, cv::cuda::GpuMat& dst, const cv::Rect2i& srcRect, const cv::Rect2i& dstRect,
                  std::array<cv::cuda::GpuMat, 4> weights, cv::cuda::Stream& stream);                                                          // :107
void applyAugmentation(const cv::cuda::GpuMat& src, cv::cuda::GpuMat& dst, const cv::Size2i& dstSize, cv::cuda::GpuMat& tmp, int augmentationIndex,
                       cv::cuda::Stream& stream);                                                                                               // :134
void reverseAugmentation(const cv::cuda::GpuMat& src, cv::cuda::GpuMat& dst, const cv::Size2i& dstSize, cv::cuda::GpuMat& tmp, int augmentationIndex,
                         cv::cuda::Stream& stream);                                                                                             // :179
void createTileWeights(std::array<cv::cuda::GpuMat, 4>& weights, const cv::Point2i& overlap, const cv::Size2i& size, cv::cuda::Stream& stream);  // img2img_load.cpp:29
bool isCompatible(const trt::RenderConfig& renderConfig, const trt::BuildConfig& buildConfig);                                                 // img2img_load.cpp:9
bool isOptimized(const trt::RenderConfig& renderConfig, const trt::BuildConfig& buildConfig);                                                  // :22
std::string getEnginePath(const std::string& modelPath, const trt::RenderConfig& config);                                                      // :79
std::string getConfigHash(const trt::BuildConfig& config);                                                                                     // img2img_build.cpp:8
void serializeConfig(const std::string& path, const trt::BuildConfig& config);                                                                 // :29
cv::cuda::GpuMat blobFromImages(const std::vector<cv::cuda::GpuMat>& images, cv::cuda::Stream& stream);                                        // img2img_infer.cpp:5

namespace {
cv::cuda::Stream gStream;

// dense host array <-> GpuMat
cv::cuda::GpuMat toGpu(const void* p, int rows, int cols, int type) {
    cv::Mat m(rows, cols, type, const_cast<void*>(p));
    cv::cuda::GpuMat g;
    g.upload(m, gStream);
    return g;
}
void fromGpu(const cv::cuda::GpuMat& g, void* p) {
    const size_t rowBytes = (size_t)g.cols * g.elemSize();
    for (int y = 0; y < g.rows; ++y) std::memcpy((uint8_t*)p + rowBytes * y, g.ptr(y), rowBytes);
}

// plain-int mirrors of trt::BuildConfig / trt::RenderConfig (same field order as include/w2x.h)
struct CBuild { int deviceId, precision, minB, optB, maxB, minC, optC, maxC, minW, optW, maxW, minH, optH, maxH; };
struct CRender { int deviceId, precision, batchSize, channels, height, width, scaling; double overlapX, overlapY; int tta; };

trt::BuildConfig toBuild(const CBuild* c) {
    trt::BuildConfig b;
    b.deviceId = c->deviceId;
    b.precision = c->precision == 0 ? trt::Precision::TF32 : trt::Precision::FP16;
    b.minBatchSize = c->minB; b.optBatchSize = c->optB; b.maxBatchSize = c->maxB;
    b.minChannels = c->minC; b.optChannels = c->optC; b.maxChannels = c->maxC;
    b.minWidth = c->minW; b.optWidth = c->optW; b.maxWidth = c->maxW;
    b.minHeight = c->minH; b.optHeight = c->optH; b.maxHeight = c->maxH;
    return b;
}
trt::RenderConfig toRender(const CRender* c) {
    trt::RenderConfig r;
    r.deviceId = c->deviceId;
    r.precision = c->precision == 0 ? trt::Precision::TF32 : trt::Precision::FP16;
    r.batchSize = c->batchSize; r.channels = c->channels; r.height = c->height; r.width = c->width; r.scaling = c->scaling;
    r.overlap = cv::Point2d(c->overlapX, c->overlapY);
    r.tta = c->tta != 0;
    return r;
}
void copyOut(const std::string& s, char* out, size_t cap) {
    if (!out || !cap) return;
    std::strncpy(out, s.c_str(), cap - 1);
    out[cap - 1] = 0;
}

struct Handle {
    trt::Img2Img img;
    std::string log;
};
}  // namespace

extern "C" {

// ---- mock configuration -------------------------------------------------------------------------------------------
void ref_shim_set_model(int scale, int outMinus, w2xshim::ModelFn fn, void* user) {
    auto& m = w2xshim::model();
    m.scale = scale; m.outMinus = outMinus; m.fn = fn; m.user = user;
}
void ref_shim_set_device_name(const char* name) { w2xshim::deviceName() = name; }
void ref_shim_set_pitch_align(int bytes) { w2xshim::pitchAlign() = bytes > 0 ? (size_t)bytes : 1; }

// ---- img2img_render.cpp free functions ----------------------------------------------------------------------------
// rects as 4 ints (x, y, w, h); returns tileCount (tiling.x * tiling.y, which may exceed what fits `cap`), or INT_MIN when the
// reference throws (vector::reserve of a negative count, for frames no larger than the overlap; render() reports that as a failure)
int ref_calculate_tiles(int inW, int inH, int outW, int outH, int tileW, int tileH, int outTileW, int outTileH, int scaling, double ovX, double ovY,
                        int* inRects, int* outRects, int cap) try {
    auto [count, in, out] = calculateTiles(cv::Rect2i(0, 0, inW, inH), cv::Rect2i(0, 0, outW, outH), cv::Size2i(tileW, tileH), cv::Size2i(outTileW, outTileH),
                                           scaling, cv::Point2d(ovX, ovY));
    for (size_t i = 0; i < in.size() && (int)i < cap; ++i) {
        inRects[4 * i] = in[i].x; inRects[4 * i + 1] = in[i].y; inRects[4 * i + 2] = in[i].width; inRects[4 * i + 3] = in[i].height;
        outRects[4 * i] = out[i].x; outRects[4 * i + 1] = out[i].y; outRects[4 * i + 2] = out[i].width; outRects[4 * i + 3] = out[i].height;
    }
    return count;
} catch (...) {
    return -2147483647 - 1;
}

// u8 HxWx3 image, roi (may lie outside) -> roiH x roiW x 3
int ref_pad_roi(const uint8_t* img, int w, int h, int rx, int ry, int rw, int rh, uint8_t* out) {
    try {
        cv::cuda::GpuMat g = toGpu(img, h, w, CV_8UC3);
        cv::cuda::GpuMat t = padRoi(g, cv::Rect2i(rx, ry, rw, rh), gStream);
        if (t.rows != rh || t.cols != rw) return 0;
        fromGpu(t, out);
        return 1;
    } catch (...) { return 0; }
}

// square u8 tile n x n x 3 (forward) / f32 n x n x 3 (reverse), aug 0..7
int ref_apply_augmentation(const uint8_t* tile, int n, int aug, uint8_t* out) {
    try {
        cv::cuda::GpuMat src = toGpu(tile, n, n, CV_8UC3), dst, tmp;
        applyAugmentation(src, dst, cv::Size2i(n, n), tmp, aug, gStream);
        fromGpu(dst, out);
        return 1;
    } catch (...) { return 0; }
}
// aliased != 0 reproduces the reference's call (img2img_render.cpp:310-311: dst and tmp are the same matrix)
int ref_reverse_augmentation(const float* tile, int n, int aug, int aliased, float* out) {
    try {
        cv::cuda::GpuMat src = toGpu(tile, n, n, CV_32FC3), dst(n, n, CV_32FC3), tmp(n, n, CV_32FC3);
        if (aliased) reverseAugmentation(src, dst, cv::Size2i(n, n), dst, aug, gStream);
        else reverseAugmentation(src, dst, cv::Size2i(n, n), tmp, aug, gStream);
        fromGpu(dst, out);
        return 1;
    } catch (...) { return 0; }
}

// 4 weight images [4][size][size][3] f32, order top, right, bottom, left (weights[0..3])
int ref_create_tile_weights(int ovX, int ovY, int sizeW, int sizeH, float* out) {
    try {
        std::array<cv::cuda::GpuMat, 4> w;
        createTileWeights(w, cv::Point2i(ovX, ovY), cv::Size2i(sizeW, sizeH), gStream);
        for (int i = 0; i < 4; ++i) fromGpu(w[i], out + (size_t)i * sizeW * sizeH * 3);
        return 1;
    } catch (...) { return 0; }
}

// tile (f32 size x size x 3) weighted in place as render() does for the clipped output rect (x, y, w, h) on a canvas
int ref_apply_weights(float* tile, int size, int ovX, int ovY, int rx, int ry, int rw, int rh, int canvasW, int canvasH) {
    try {
        std::array<cv::cuda::GpuMat, 4> w;
        createTileWeights(w, cv::Point2i(ovX, ovY), cv::Size2i(size, size), gStream);
        cv::cuda::GpuMat t = toGpu(tile, size, size, CV_32FC3);
        applyWeights(t, t, cv::Rect2i(rx, ry, rw, rh), cv::Rect2i(0, 0, canvasW, canvasH), w, gStream);
        fromGpu(t, tile);
        return 1;
    } catch (...) { return 0; }
}

// n u8 tiles (T x T x 3, RGB) -> blob [n][3][T][T] f32 exactly as infer() uploads it (row pitch of the blob ignored, q4)
int ref_blob_from_images(const uint8_t* tiles, int n, int t, float* out) {
    try {
        std::vector<cv::cuda::GpuMat> v;
        for (int i = 0; i < n; ++i) v.push_back(toGpu(tiles + (size_t)i * t * t * 3, t, t, CV_8UC3));
        cv::cuda::GpuMat blob = blobFromImages(v, gStream);
        std::memcpy(out, blob.ptr<void>(), (size_t)n * 3 * t * t * sizeof(float));  // img2img_infer.cpp:76 copies linearly
        return 1;
    } catch (...) { return 0; }
}

// ---- img2img_build.cpp / img2img_load.cpp host logic -----------------------------------------------------------------
void ref_config_hash(const CBuild* cfg, char out65[65]) { copyOut(getConfigHash(toBuild(cfg)), out65, 65); }
int ref_serialize_config(const char* path, const CBuild* cfg) {
    try { serializeConfig(path, toBuild(cfg)); return 1; } catch (...) { return 0; }
}
int ref_is_compatible(const CRender* r, const CBuild* b) { return isCompatible(toRender(r), toBuild(b)) ? 1 : 0; }
int ref_is_optimized(const CRender* r, const CBuild* b) { return isOptimized(toRender(r), toBuild(b)) ? 1 : 0; }
// returns 1 + path, or 0 + the exception text
int ref_get_engine_path(const char* modelPath, const CRender* r, char* out, size_t cap) {
    try { copyOut(getEnginePath(modelPath, toRender(r)), out, cap); return 1; }
    catch (const std::exception& e) { copyOut(e.what(), out, cap); return 0; }
}

// ---- trt::Img2Img ----------------------------------------------------------------------------------------------------
void* ref_create() {
    auto* h = new Handle();
    h->img.setMessageCallback([h](trt::Severity sev, const std::string& msg) { h->log += std::to_string((int)sev) + "|" + msg + "\n"; });
    return h;
}
void ref_destroy(void* p) { delete (Handle*)p; }
const char* ref_log(void* p) { return ((Handle*)p)->log.c_str(); }
int ref_build(void* p, const char* onnxPath, const CBuild* cfg) { return ((Handle*)p)->img.build(onnxPath, toBuild(cfg)) ? 1 : 0; }
int ref_load(void* p, const char* onnxPath, const CRender* cfg) { return ((Handle*)p)->img.load(onnxPath, toRender(cfg)) ? 1 : 0; }
// src BGR u8 h x w x 3 (dense) -> dst BGR u8 (h*s) x (w*s) x 3 (dense); dstCap in bytes
int ref_render(void* p, const uint8_t* src, int w, int h, uint8_t* dst, size_t dstCap, int* outW, int* outH) {
    cv::Mat s(h, w, CV_8UC3, const_cast<uint8_t*>(src)), d;
    if (!((Handle*)p)->img.render(s, d)) return 0;
    if (d.type() != CV_8UC3 || (size_t)d.rows * d.cols * 3 > dstCap) return 0;
    for (int y = 0; y < d.rows; ++y) std::memcpy(dst + (size_t)y * d.cols * 3, d.ptr(y), (size_t)d.cols * 3);
    *outW = d.cols; *outH = d.rows;
    return 1;
}

}  // extern "C"
