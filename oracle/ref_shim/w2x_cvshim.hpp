// TEST INFRASTRUCTURE ONLY -- part of oracle/_ref (see oracle/Makefile); never linked into the product.
//
// CPU stand-in for the slice of OpenCV (core + cudaarithm / cudawarping / cudaimgproc) that the reference's
// src/tensorrt/*.cpp uses, so that those files compile UNMODIFIED, where they lie under /root/reference, and their host
// logic (calculateTiles, padRoi, applyWeights, apply/reverseAugmentation, createTileWeights, Img2Img::load/render/infer,
// blobFromImages/imagesFromBlob) runs on the CPU.  "GpuMat" is host memory; every cv::cuda:: function below is a plain
// loop that follows the documented OpenCV-CUDA semantics of the call:
//   * GpuMat allocation is pitched like cudaMallocPitch when rows > 1 && cols > 1 (OpenCV DefaultAllocator), pitch
//     alignment w2xshim::pitchAlign() (512 B default) -- this is what makes the reference's blob code (q4) misbehave for
//     batch > 1 when 3*T*T is not pitch-aligned, exactly as its README says;
//   * create() keeps the buffer when size and type already match (so cv::cuda::split writes through headers that wrap
//     blob memory), reallocates otherwise;
//   * convertTo: saturate_cast<D>(float(alpha) * src) with round-half-even float -> uchar (cvt.rni + clamp), the number of
//     channels of rtype is ignored;
//   * multiply / add on CV_32F: one IEEE f32 operation per element;
//   * flip: code 0 reverses rows, code > 0 reverses columns;
//   * rotate: NPP nppiRotate mapping  x' = cos(a) x + sin(a) y + xShift,  y' = -sin(a) x + cos(a) y + yShift,
//     nearest neighbour, destination cleared first (cv::cuda::rotate does dst.setTo(0)); only multiples of 90 degrees;
//   * copyMakeBorder: BORDER_REPLICATE only.
// The mock is checked against the real OpenCV (cv2, CPU build, present in the image) in tests/test_ref_shim.py for every
// op cv2 also has (flip, copyMakeBorder, multiply, add, convertTo rounding, cvtColor, split/merge).
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "w2x_cudashim.hpp"

typedef unsigned char uchar;

#define CV_8U 0
#define CV_32F 5
#define CV_CN_SHIFT 3
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn)-1) << CV_CN_SHIFT))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_32FC3 CV_MAKETYPE(CV_32F, 3)
#define CV_MAT_DEPTH(t) ((t) & 7)
#define CV_MAT_CN(t) ((((t) >> CV_CN_SHIFT) & 511) + 1)

namespace w2xshim {
inline size_t& pitchAlign() {
    static size_t v = 512;
    return v;
}
}  // namespace w2xshim

namespace cv {

template <class T>
struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T x_, T y_) : x(x_), y(y_) {}
};
typedef Point_<int> Point2i;
typedef Point_<double> Point2d;
typedef Point2i Point;

template <class T>
struct Size_ {
    T width, height;
    Size_() : width(0), height(0) {}
    Size_(T w, T h) : width(w), height(h) {}
    bool operator==(const Size_& o) const { return width == o.width && height == o.height; }
    bool operator!=(const Size_& o) const { return !(*this == o); }
};
typedef Size_<int> Size2i;
typedef Size2i Size;

template <class T>
struct Rect_ {
    T x, y, width, height;
    Rect_() : x(0), y(0), width(0), height(0) {}
    Rect_(T x_, T y_, T w, T h) : x(x_), y(y_), width(w), height(h) {}
};
typedef Rect_<int> Rect2i;
typedef Rect2i Rect;

struct Scalar {
    double val[4];
    Scalar() : val{0, 0, 0, 0} {}
    Scalar(double v0, double v1 = 0, double v2 = 0, double v3 = 0) : val{v0, v1, v2, v3} {}
    static Scalar all(double v) { return Scalar(v, v, v, v); }
    double operator[](int i) const { return val[i]; }
};

enum { BORDER_CONSTANT = 0, BORDER_REPLICATE = 1 };
enum { INTER_NEAREST = 0, INTER_LINEAR = 1 };
enum { COLOR_BGR2RGB = 4, COLOR_RGB2BGR = 4 };

inline int elemSizeOf(int type) { return (CV_MAT_DEPTH(type) == CV_8U ? 1 : 4) * CV_MAT_CN(type); }

// Common header of Mat and GpuMat: a (possibly strided) view of a shared host buffer.
struct MatBase {
    int flags = 0;
    int rows = 0, cols = 0;
    size_t step = 0;
    uchar* data = nullptr;
    std::shared_ptr<std::vector<uchar>> owner;

    int type() const { return flags; }
    int depth() const { return CV_MAT_DEPTH(flags); }
    int channels() const { return CV_MAT_CN(flags); }
    size_t elemSize() const { return (size_t)elemSizeOf(flags); }
    size_t elemSize1() const { return depth() == CV_8U ? 1 : 4; }
    Size size() const { return Size(cols, rows); }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    bool isContinuous() const { return rows == 1 || step == (size_t)cols * elemSize(); }
    uchar* ptr(int y = 0) { return data + step * (size_t)y; }
    const uchar* ptr(int y = 0) const { return data + step * (size_t)y; }
    template <class T>
    T* ptr(int y = 0) { return (T*)(data + step * (size_t)y); }
    template <class T>
    const T* ptr(int y = 0) const { return (const T*)(data + step * (size_t)y); }

    void release() {
        owner.reset();
        data = nullptr;
        rows = cols = 0;
        step = 0;
    }

protected:
    void allocate(int r, int c, int t, bool pitched) {
        if (r < 0 || c < 0) throw std::runtime_error("cvshim: negative matrix size");
        flags = t;
        rows = r;
        cols = c;
        const size_t rowBytes = (size_t)c * elemSizeOf(t);
        const size_t a = w2xshim::pitchAlign();
        step = (pitched && r > 1 && c > 1 && a > 1) ? (rowBytes + a - 1) / a * a : rowBytes;
        owner = std::make_shared<std::vector<uchar>>(std::max<size_t>(step * (size_t)r, 1), (uchar)0);
        data = owner->data();
    }
    void wrap(int r, int c, int t, void* d, size_t s) {
        flags = t;
        rows = r;
        cols = c;
        step = s ? s : (size_t)c * elemSizeOf(t);
        owner.reset();
        data = (uchar*)d;
    }
    void viewOf(const MatBase& m, const Rect2i& roi) {
        if (roi.x < 0 || roi.y < 0 || roi.width < 0 || roi.height < 0 || roi.x + roi.width > m.cols || roi.y + roi.height > m.rows)
            throw std::runtime_error("cvshim: ROI outside the matrix (OpenCV would assert)");
        flags = m.flags;
        rows = roi.height;
        cols = roi.width;
        step = m.step;
        owner = m.owner;
        data = m.data + m.step * (size_t)roi.y + (size_t)roi.x * m.elemSize();
    }
};

class Mat : public MatBase {
public:
    static constexpr size_t AUTO_STEP = 0;
    Mat() {}
    Mat(int r, int c, int t) { allocate(r, c, t, false); }
    Mat(Size s, int t) { allocate(s.height, s.width, t, false); }
    Mat(int r, int c, int t, void* d, size_t s = AUTO_STEP) { wrap(r, c, t, d, s); }
    void create(int r, int c, int t) {
        if (data && r == rows && c == cols && t == flags) return;
        allocate(r, c, t, false);
    }
    void create(Size s, int t) { create(s.height, s.width, t); }
    Mat operator()(const Rect2i& roi) const {
        Mat v;
        v.viewOf(*this, roi);
        return v;
    }
};

namespace cuda {

class Stream {
public:
    Stream() {}
    explicit Stream(unsigned) {}
    void waitForCompletion() {}
    static Stream& Null() {
        static Stream s;
        return s;
    }
};

struct StreamAccessor {
    static cudaStream_t getStream(const Stream&) { return nullptr; }
};

class GpuMat;

// OutputArray stand-in: binds lvalues and temporaries (OpenCV's _OutputArray is a const-ref proxy as well)
struct OutArr {
    GpuMat* m;
    OutArr(const GpuMat& g) : m(const_cast<GpuMat*>(&g)) {}
};

class GpuMat : public MatBase {
public:
    static constexpr size_t AUTO_STEP = 0;
    GpuMat() {}
    GpuMat(int r, int c, int t) { allocate(r, c, t, true); }
    GpuMat(Size s, int t) { allocate(s.height, s.width, t, true); }
    GpuMat(int r, int c, int t, Scalar v) {
        allocate(r, c, t, true);
        fill(v);
    }
    GpuMat(Size s, int t, Scalar v) {
        allocate(s.height, s.width, t, true);
        fill(v);
    }
    GpuMat(int r, int c, int t, void* d, size_t s = AUTO_STEP) { wrap(r, c, t, d, s); }
    GpuMat(Size sz, int t, void* d, size_t s = AUTO_STEP) { wrap(sz.height, sz.width, t, d, s); }

    void create(int r, int c, int t) {
        if (data && r == rows && c == cols && t == flags) return;
        allocate(r, c, t, true);
    }
    void create(Size s, int t) { create(s.height, s.width, t); }

    GpuMat operator()(const Rect2i& roi) const {
        GpuMat v;
        v.viewOf(*this, roi);
        return v;
    }
    GpuMat row(int y) const { return (*this)(Rect2i(0, y, cols, 1)); }
    GpuMat col(int x) const { return (*this)(Rect2i(x, 0, 1, rows)); }

    void upload(const Mat& src, Stream& = Stream::Null()) {
        create(src.rows, src.cols, src.type());
        const size_t rowBytes = (size_t)cols * elemSize();
        for (int y = 0; y < rows; ++y) std::memcpy(ptr(y), src.ptr(y), rowBytes);
    }
    void download(Mat& dst, Stream& = Stream::Null()) const {
        dst.create(rows, cols, type());
        const size_t rowBytes = (size_t)cols * elemSize();
        for (int y = 0; y < rows; ++y) std::memcpy(dst.ptr(y), ptr(y), rowBytes);
    }
    void copyTo(OutArr dst, Stream& = Stream::Null()) const {
        GpuMat src = *this;
        dst.m->create(src.rows, src.cols, src.type());
        const size_t rowBytes = (size_t)src.cols * src.elemSize();
        for (int y = 0; y < src.rows; ++y) std::memmove(dst.m->ptr(y), src.ptr(y), rowBytes);
    }
    GpuMat& setTo(Scalar v, Stream& = Stream::Null()) {
        fill(v);
        return *this;
    }
    // GpuMat::convertTo(dst, rtype, alpha, stream): dst = saturate_cast<D>(float(alpha) * src); channels follow the source
    void convertTo(OutArr dst, int rtype, double alpha, Stream& s = Stream::Null()) const { convertTo(dst, rtype, alpha, 0.0, s); }
    void convertTo(OutArr dst, int rtype, double alpha, double beta, Stream& = Stream::Null()) const {
        GpuMat src = *this;  // keeps the source buffer alive when dst is this very matrix (in-place with a type change)
        const int ddepth = CV_MAT_DEPTH(rtype), cn = src.channels();
        GpuMat out;
        if (dst.m->data && dst.m->rows == src.rows && dst.m->cols == src.cols && dst.m->type() == CV_MAKETYPE(ddepth, cn) && dst.m->data != src.data)
            out = *dst.m;
        else
            out = GpuMat(src.rows, src.cols, CV_MAKETYPE(ddepth, cn));
        const float a = (float)alpha, b = (float)beta;
        const int n = src.cols * cn;
        for (int y = 0; y < src.rows; ++y) {
            for (int i = 0; i < n; ++i) {
                const float v = src.depth() == CV_8U ? (float)src.ptr(y)[i] : src.ptr<float>(y)[i];
                const float r = std::fmaf(a, v, b);
                if (ddepth == CV_8U) {
                    // saturate_cast<uchar>(float) on the device: __float2int_rn (round half to even), then clamp
                    float q = std::nearbyintf(r);
                    if (!(q > 0.f)) q = 0.f;  // also maps NaN to 0
                    if (q > 255.f) q = 255.f;
                    out.ptr(y)[i] = (uchar)q;
                } else {
                    out.ptr<float>(y)[i] = r;
                }
            }
        }
        *dst.m = out;
    }

private:
    void fill(const Scalar& v) {
        const int cn = channels();
        for (int y = 0; y < rows; ++y)
            for (int x = 0; x < cols; ++x)
                for (int c = 0; c < cn; ++c) {
                    if (depth() == CV_8U) {
                        double q = std::nearbyint(v.val[c & 3]);
                        q = std::min(255.0, std::max(0.0, q));
                        ptr(y)[x * cn + c] = (uchar)q;
                    } else {
                        ptr<float>(y)[x * cn + c] = (float)v.val[c & 3];
                    }
                }
    }
};

struct NoArray {};

inline void requireF32(const GpuMat& m, const char* what) {
    if (m.depth() != CV_32F) throw std::runtime_error(std::string("cvshim: ") + what + " is only mocked for CV_32F");
}

// cv::cuda::multiply(src1, src2, dst, scale = 1, dtype = -1): per-element f32 product
inline void multiply(const GpuMat& a, const GpuMat& b, OutArr dst, double scale = 1, int = -1, Stream& = Stream::Null()) {
    requireF32(a, "multiply");
    if (a.size() != b.size() || a.type() != b.type()) throw std::runtime_error("cvshim: multiply size/type mismatch");
    GpuMat s1 = a, s2 = b;
    dst.m->create(a.rows, a.cols, a.type());
    const int n = a.cols * a.channels();
    const float sc = (float)scale;
    for (int y = 0; y < a.rows; ++y) {
        const float* pa = s1.ptr<float>(y);
        const float* pb = s2.ptr<float>(y);
        float* pd = dst.m->ptr<float>(y);
        for (int i = 0; i < n; ++i) pd[i] = scale == 1 ? pa[i] * pb[i] : sc * pa[i] * pb[i];
    }
}
// scalar second operand: a double converts to Scalar(v, 0, 0, 0) exactly as cv::InputArray(double) does, one value per channel
inline void multiply(const GpuMat& a, double v, OutArr dst, double = 1, int = -1, Stream& = Stream::Null()) {
    requireF32(a, "multiply");
    GpuMat s1 = a;
    dst.m->create(a.rows, a.cols, a.type());
    const Scalar sv(v);
    const int cn = a.channels();
    for (int y = 0; y < a.rows; ++y)
        for (int x = 0; x < a.cols; ++x)
            for (int c = 0; c < cn; ++c) dst.m->ptr<float>(y)[x * cn + c] = s1.ptr<float>(y)[x * cn + c] * (float)sv.val[c & 3];
}
inline void add(const GpuMat& a, const GpuMat& b, OutArr dst, NoArray = NoArray(), int = -1, Stream& = Stream::Null()) {
    requireF32(a, "add");
    if (a.size() != b.size() || a.type() != b.type()) throw std::runtime_error("cvshim: add size/type mismatch");
    GpuMat s1 = a, s2 = b;
    dst.m->create(a.rows, a.cols, a.type());
    const int n = a.cols * a.channels();
    for (int y = 0; y < a.rows; ++y) {
        const float* pa = s1.ptr<float>(y);
        const float* pb = s2.ptr<float>(y);
        float* pd = dst.m->ptr<float>(y);
        for (int i = 0; i < n; ++i) pd[i] = pa[i] + pb[i];
    }
}

// flipCode 0: around the x axis (rows reversed); > 0: around the y axis (columns reversed); < 0: both
inline void flip(const GpuMat& src, OutArr dst, int flipCode, Stream& = Stream::Null()) {
    GpuMat s = src, tmp(src.rows, src.cols, src.type());
    const size_t es = s.elemSize();
    for (int y = 0; y < s.rows; ++y) {
        const int sy = flipCode <= 0 ? s.rows - 1 - y : y;
        for (int x = 0; x < s.cols; ++x) {
            const int sx = flipCode != 0 ? s.cols - 1 - x : x;
            std::memcpy(tmp.ptr(y) + (size_t)x * es, s.ptr(sy) + (size_t)sx * es, es);
        }
    }
    dst.m->create(s.rows, s.cols, s.type());
    for (int y = 0; y < s.rows; ++y) std::memcpy(dst.m->ptr(y), tmp.ptr(y), (size_t)s.cols * es);
}

// cv::cuda::rotate -> nppiRotate: rotation about the origin by `angle` degrees, then the shift; nearest neighbour
inline void rotate(const GpuMat& src, OutArr dst, Size dsize, double angle, double xShift = 0, double yShift = 0, int = INTER_LINEAR,
                   Stream& = Stream::Null()) {
    const int q = (int)std::lround(angle / 90.0);
    if (std::fabs(angle - 90.0 * q) > 1e-9) throw std::runtime_error("cvshim: rotate is only mocked for multiples of 90 degrees");
    static const int cosT[4] = {1, 0, -1, 0}, sinT[4] = {0, 1, 0, -1};
    const int c = cosT[((q % 4) + 4) % 4], s = sinT[((q % 4) + 4) % 4];
    const int sx = (int)std::lround(xShift), sy = (int)std::lround(yShift);
    GpuMat in = src, tmp(dsize.height, dsize.width, src.type());
    const size_t es = in.elemSize();
    for (int y = 0; y < in.rows; ++y)
        for (int x = 0; x < in.cols; ++x) {
            const int dx = c * x + s * y + sx, dy = -s * x + c * y + sy;
            if (dx < 0 || dy < 0 || dx >= dsize.width || dy >= dsize.height) continue;
            std::memcpy(tmp.ptr(dy) + (size_t)dx * es, in.ptr(y) + (size_t)x * es, es);
        }
    dst.m->create(dsize.height, dsize.width, src.type());
    for (int y = 0; y < dsize.height; ++y) std::memcpy(dst.m->ptr(y), tmp.ptr(y), (size_t)dsize.width * es);
}

inline void copyMakeBorder(const GpuMat& src, OutArr dst, int top, int bottom, int left, int right, int borderType, Scalar = Scalar(),
                           Stream& = Stream::Null()) {
    if (borderType != BORDER_REPLICATE) throw std::runtime_error("cvshim: copyMakeBorder is only mocked for BORDER_REPLICATE");
    GpuMat s = src, out(src.rows + top + bottom, src.cols + left + right, src.type());
    const size_t es = s.elemSize();
    for (int y = 0; y < out.rows; ++y) {
        const int sy = std::min(std::max(y - top, 0), s.rows - 1);
        for (int x = 0; x < out.cols; ++x) {
            const int sx = std::min(std::max(x - left, 0), s.cols - 1);
            std::memcpy(out.ptr(y) + (size_t)x * es, s.ptr(sy) + (size_t)sx * es, es);
        }
    }
    *dst.m = out;
}

// only the 3-channel channel swap (COLOR_BGR2RGB == COLOR_RGB2BGR) is used by the reference
inline void cvtColor(const GpuMat& src, OutArr dst, int code, int = 0, Stream& = Stream::Null()) {
    if (code != COLOR_BGR2RGB || src.channels() != 3) throw std::runtime_error("cvshim: cvtColor is only mocked for the 3-channel swap");
    GpuMat s = src;
    dst.m->create(s.rows, s.cols, s.type());
    const size_t e1 = s.elemSize1();
    for (int y = 0; y < s.rows; ++y)
        for (int x = 0; x < s.cols; ++x) {
            uchar px[12];
            std::memcpy(px, s.ptr(y) + (size_t)x * 3 * e1, 3 * e1);
            uchar* d = dst.m->ptr(y) + (size_t)x * 3 * e1;
            std::memcpy(d, px + 2 * e1, e1);
            std::memcpy(d + e1, px + e1, e1);
            std::memcpy(d + 2 * e1, px, e1);
        }
}

// split into a caller-provided vector of headers: each is create()d (kept when it already has the right size and depth)
inline void split(const GpuMat& src, std::vector<GpuMat>& dst, Stream& = Stream::Null()) {
    const int cn = src.channels();
    dst.resize(cn);
    const size_t e1 = src.elemSize1();
    for (int c = 0; c < cn; ++c) dst[c].create(src.rows, src.cols, CV_MAKETYPE(src.depth(), 1));
    for (int y = 0; y < src.rows; ++y)
        for (int x = 0; x < src.cols; ++x)
            for (int c = 0; c < cn; ++c) std::memcpy(dst[c].ptr(y) + (size_t)x * e1, src.ptr(y) + ((size_t)x * cn + c) * e1, e1);
}

inline void merge(const std::vector<GpuMat>& src, OutArr dst, Stream& = Stream::Null()) {
    if (src.empty()) throw std::runtime_error("cvshim: merge of nothing");
    const int cn = (int)src.size();
    const size_t e1 = src[0].elemSize1();
    dst.m->create(src[0].rows, src[0].cols, CV_MAKETYPE(src[0].depth(), cn));
    for (int y = 0; y < src[0].rows; ++y)
        for (int x = 0; x < src[0].cols; ++x)
            for (int c = 0; c < cn; ++c) std::memcpy(dst.m->ptr(y) + ((size_t)x * cn + c) * e1, src[c].ptr(y) + (size_t)x * e1, e1);
}

}  // namespace cuda

inline cuda::NoArray noArray() { return cuda::NoArray(); }

}  // namespace cv
