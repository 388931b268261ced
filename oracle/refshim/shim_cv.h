// shim_cv.h -- a minimal stand-in for the subset of OpenCV (2.4 API) that
// /root/reference/include/RegisterPhotoICP.h uses, so the reference header can be compiled here.
//
// *** TEST INFRASTRUCTURE ONLY. ***  Written from scratch (C++98).
//
// The three OpenCV operations whose arithmetic is on the spherical registration path are restated
// and pinned against python cv2 4.13 golden vectors (tests/golden/cv2_vectors.npz):
//   cvtColor(CV_RGB2GRAY) 8UC3 : (R*9798 + G*19235 + B*3735 + 2^14) >> 15          bit-exact vs cv2
//   Mat::convertTo(CV_32F, s)  : (float)src * (float)s                               bit-exact vs cv2
//   pyrDown 32FC1              : separable [1 4 6 4 1], BORDER_REFLECT_101, /256     <= 4 ulp vs cv2
// Display / file functions (imshow, waitKey, imwrite, ...) are no-ops; filters that the spherical
// path never calls (Sobel, Scharr) abort.
#pragma once
#include <cassert>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <string>
#include <vector>
#include <iostream>
#include <sys/time.h>
#include <limits>

typedef unsigned char uchar;    // OpenCV declares these at global scope
typedef unsigned short ushort;

#define CV_8U 0
#define CV_16U 2
#define CV_32F 5
#define CV_64F 6
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn)-1) << 3))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_16UC1 CV_MAKETYPE(CV_16U, 1)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_32FC3 CV_MAKETYPE(CV_32F, 3)
#define CV_64FC1 CV_MAKETYPE(CV_64F, 1)
#define CV_RGB2GRAY 7
#define CV_BGR2GRAY 6
#define CV_WINDOW_AUTOSIZE 1

namespace cv {

using ::uchar;
using ::ushort;

template <typename T, int N>
struct Vec {
    T val[N];
    T& operator[](int i) { return val[i]; }
    const T& operator[](int i) const { return val[i]; }
};
typedef Vec<uchar, 3> Vec3b;
typedef Vec<float, 3> Vec3f;

struct Size {
    int width, height;
    Size() : width(0), height(0) {}
    Size(int w, int h) : width(w), height(h) {}
};
struct Rect {
    int x, y, width, height;
    Rect() : x(0), y(0), width(0), height(0) {}
    Rect(int x_, int y_, int w, int h) : x(x_), y(y_), width(w), height(h) {}
};
struct Scalar {
    double val[4];
    Scalar(double a = 0, double b = 0, double c = 0, double d = 0) { val[0] = a; val[1] = b; val[2] = c; val[3] = d; }
    double& operator[](int i) { return val[i]; }
    const double& operator[](int i) const { return val[i]; }
};

struct ShimBuf {   // reference-counted pixel storage (cv::Mat headers share it)
    uchar* p;
    int refs;
};

struct MatExpr;   // only Mat::zeros

class Mat {
public:
    int rows, cols;
    uchar* data;
    size_t step;   // bytes per row
    int type_;
    ShimBuf* buf;

    static int elemSizeOf(int type) {
        static const int dsz[8] = { 1, 1, 2, 2, 4, 4, 8, 0 };
        return dsz[type & 7] * ((type >> 3) + 1);
    }
    Mat() : rows(0), cols(0), data(0), step(0), type_(0), buf(0) {}
    Mat(int r, int c, int type) : rows(0), cols(0), data(0), step(0), type_(0), buf(0) { create(r, c, type); }
    Mat(Size s, int type) : rows(0), cols(0), data(0), step(0), type_(0), buf(0) { create(s.height, s.width, type); }
    Mat(int r, int c, int type, const Scalar& v) : rows(0), cols(0), data(0), step(0), type_(0), buf(0) { create(r, c, type); setTo(v); }
    // external data, not owned (as cv::Mat(rows, cols, type, void*))
    Mat(int r, int c, int type, void* ext, size_t step_ = 0) : rows(r), cols(c), data((uchar*)ext), step(step_ ? step_ : (size_t)c * elemSizeOf(type)), type_(type), buf(0) {}
    Mat(const Mat& o) : rows(o.rows), cols(o.cols), data(o.data), step(o.step), type_(o.type_), buf(o.buf) { if (buf) ++buf->refs; }
    Mat(const MatExpr& e);
    ~Mat() { release(); }
    Mat& operator=(const Mat& o) {
        if (this == &o) return *this;
        if (o.buf) ++o.buf->refs;
        release();
        rows = o.rows; cols = o.cols; data = o.data; step = o.step; type_ = o.type_; buf = o.buf;
        return *this;
    }
    Mat& operator=(const MatExpr& e);
    void release() {
        if (buf && --buf->refs == 0) { free(buf->p); delete buf; }
        buf = 0; data = 0; rows = cols = 0;
    }
    void create(int r, int c, int type) {
        if (data && r == rows && c == cols && type == type_) return;   // as cv::Mat::create
        release();
        rows = r; cols = c; type_ = type;
        step = (size_t)c * elemSizeOf(type);
        buf = new ShimBuf;
        buf->p = (uchar*)malloc(step * (size_t)(r > 0 ? r : 1) + 64);
        buf->refs = 1;
        data = buf->p;
    }
    int type() const { return type_; }
    int depth() const { return type_ & 7; }
    int channels() const { return (type_ >> 3) + 1; }
    size_t elemSize() const { return elemSizeOf(type_); }
    bool empty() const { return data == 0 || rows * cols == 0; }
    Size size() const { return Size(cols, rows); }
    size_t total() const { return (size_t)rows * cols; }
    bool isContinuous() const { return step == (size_t)cols * elemSize(); }

    template <typename T> T& at(int r, int c) { return *(T*)(data + (size_t)r * step + (size_t)c * sizeof(T)); }
    template <typename T> const T& at(int r, int c) const { return *(const T*)(data + (size_t)r * step + (size_t)c * sizeof(T)); }
    template <typename T> T& at(int i) { return at<T>(i / cols, i % cols); }
    template <typename T> const T& at(int i) const { return at<T>(i / cols, i % cols); }
    template <typename T> T* ptr(int r = 0) { return (T*)(data + (size_t)r * step); }
    template <typename T> const T* ptr(int r = 0) const { return (const T*)(data + (size_t)r * step); }

    Mat operator()(const Rect& roi) const {
        Mat m(*this);
        m.data = data + (size_t)roi.y * step + (size_t)roi.x * elemSize();
        m.rows = roi.height; m.cols = roi.width;
        return m;
    }
    void setTo(const Scalar& v) {
        const int cn = channels();
        for (int r = 0; r < rows; ++r)
            for (int c = 0; c < cols; ++c)
                for (int k = 0; k < cn; ++k) {
                    uchar* e = data + (size_t)r * step + ((size_t)c * cn + k) * (elemSize() / cn);
                    switch (depth()) {
                        case CV_8U: *e = (uchar)v.val[k]; break;
                        case CV_16U: *(ushort*)e = (ushort)v.val[k]; break;
                        case CV_32F: *(float*)e = (float)v.val[k]; break;
                        case CV_64F: *(double*)e = v.val[k]; break;
                        default: abort();
                    }
                }
    }
    Mat clone() const { Mat m; copyTo(m); return m; }
    void copyTo(Mat& dst) const {
        dst.create(rows, cols, type_);
        for (int r = 0; r < rows; ++r) memcpy(dst.data + (size_t)r * dst.step, data + (size_t)r * step, (size_t)cols * elemSize());
    }
    void copyTo(const Mat& roi) const { Mat d(roi); copyTo(d); }   // copy into an ROI temporary
    // dst(x) = saturate_cast<dst type>( src(x) * alpha + beta ).  For a 32F destination OpenCV's
    // cvt path computes in float: (float)src * (float)alpha + (float)beta; beta == 0 here.
    void convertTo(Mat& dst, int rtype, double alpha = 1, double beta = 0) const {
        const int ddepth = rtype & 7;
        const int cn = channels();
        Mat out(rows, cols, CV_MAKETYPE(ddepth, cn));
        const float a = (float)alpha, b = (float)beta;
        for (int r = 0; r < rows; ++r)
            for (int c = 0; c < cols * cn; ++c) {
                float v;
                switch (depth()) {
                    case CV_8U: v = (float)ptr<uchar>(r)[c]; break;
                    case CV_16U: v = (float)ptr<ushort>(r)[c]; break;
                    case CV_32F: v = ptr<float>(r)[c]; break;
                    default: abort();
                }
                if (alpha != 1 || beta != 0) { v = v * a; if (beta != 0) v = v + b; }
                switch (ddepth) {
                    case CV_32F: out.ptr<float>(r)[c] = v; break;
                    case CV_8U: { long q = lrintf(v); out.ptr<uchar>(r)[c] = (uchar)(q < 0 ? 0 : q > 255 ? 255 : q); } break;
                    case CV_16U: { long q = lrintf(v); out.ptr<ushort>(r)[c] = (ushort)(q < 0 ? 0 : q > 65535 ? 65535 : q); } break;
                    default: abort();
                }
            }
        dst = out;
    }
    static MatExpr zeros(int r, int c, int type);
    static MatExpr zeros(Size s, int type);
};

struct MatExpr {
    int rows, cols, type;
};
inline MatExpr Mat::zeros(int r, int c, int type) { MatExpr e; e.rows = r; e.cols = c; e.type = type; return e; }
inline MatExpr Mat::zeros(Size s, int type) { return zeros(s.height, s.width, type); }
inline Mat::Mat(const MatExpr& e) : rows(0), cols(0), data(0), step(0), type_(0), buf(0) { *this = e; }
// cv::Mat::operator=(const MatExpr&): create() (a no-op when size and type already match, so a
// ROI header is filled IN PLACE -- the sensor-joint mask of RPI.h:4541-4549 relies on this), then fill.
inline Mat& Mat::operator=(const MatExpr& e) {
    create(e.rows, e.cols, e.type);
    for (int r = 0; r < rows; ++r) memset(data + (size_t)r * step, 0, (size_t)cols * elemSize());
    return *this;
}

inline int shim_reflect101(int i, int n) {
    if (n == 1) return 0;
    while (i < 0 || i >= n) { if (i < 0) i = -i; if (i >= n) i = 2 * n - 2 - i; }
    return i;
}

inline void cvtColor(const Mat& src, Mat& dst, int code, int = 0) {
    assert(src.type() == CV_8UC3);
    // RGB2GRAY takes channel 0 as R; BGR2GRAY takes channel 0 as B
    const int c0 = (code == CV_RGB2GRAY) ? 9798 : 3735, c2 = (code == CV_RGB2GRAY) ? 3735 : 9798;
    Mat out(src.rows, src.cols, CV_8UC1);
    for (int r = 0; r < src.rows; ++r) {
        const uchar* s = src.ptr<uchar>(r);
        uchar* d = out.ptr<uchar>(r);
        for (int c = 0; c < src.cols; ++c)
            d[c] = (uchar)((s[3 * c] * c0 + s[3 * c + 1] * 19235 + s[3 * c + 2] * c2 + (1 << 14)) >> 15);
    }
    dst = out;
}

// pyrDown, OpenCV 2.4 operation order: horizontal  6*s2 + 4*(s1+s3) + s0 + s4  (PyrDownVec-less scalar
// row filter), vertical  ((r0+r4) + 2*r2) + 4*((r1+r3)+r2)  then * 1/256  (the SSE column filter).
inline void pyrDown(const Mat& src, Mat& dst, const Size& dsz = Size(), int = 4) {
    const int h = dsz.height ? dsz.height : (src.rows + 1) / 2, w = dsz.width ? dsz.width : (src.cols + 1) / 2;
    if (src.type() == CV_32FC1) {
        std::vector<float> hb((size_t)src.rows * w);
        for (int r = 0; r < src.rows; ++r) {
            const float* s = src.ptr<float>(r);
            for (int x = 0; x < w; ++x) {
                const float s0 = s[shim_reflect101(2 * x - 2, src.cols)], s1 = s[shim_reflect101(2 * x - 1, src.cols)],
                            s2 = s[shim_reflect101(2 * x, src.cols)], s3 = s[shim_reflect101(2 * x + 1, src.cols)],
                            s4 = s[shim_reflect101(2 * x + 2, src.cols)];
                hb[(size_t)r * w + x] = s2 * 6 + (s1 + s3) * 4 + s0 + s4;
            }
        }
        Mat out(h, w, CV_32FC1);
        for (int y = 0; y < h; ++y) {
            const float* r0 = &hb[(size_t)shim_reflect101(2 * y - 2, src.rows) * w];
            const float* r1 = &hb[(size_t)shim_reflect101(2 * y - 1, src.rows) * w];
            const float* r2 = &hb[(size_t)shim_reflect101(2 * y, src.rows) * w];
            const float* r3 = &hb[(size_t)shim_reflect101(2 * y + 1, src.rows) * w];
            const float* r4 = &hb[(size_t)shim_reflect101(2 * y + 2, src.rows) * w];
            float* d = out.ptr<float>(y);
            for (int x = 0; x < w; ++x) {
                const float t0 = (r0[x] + r4[x]) + (r2[x] + r2[x]);
                const float t1 = ((r1[x] + r3[x]) + r2[x]) * 4.f;
                d[x] = (t0 + t1) * (1.f / 256);
            }
        }
        dst = out;
    } else if (src.depth() == CV_8U) {   // colour pyramid of setSourceFrame (RPI.h:495): visualisation only
        const int cn = src.channels();
        Mat out(h, w, src.type());
        for (int y = 0; y < h; ++y)
            for (int x = 0; x < w; ++x)
                for (int k = 0; k < cn; ++k) {
                    static const int wt[5] = { 1, 4, 6, 4, 1 };
                    int acc = 0;
                    for (int i = -2; i <= 2; ++i)
                        for (int j = -2; j <= 2; ++j)
                            acc += wt[i + 2] * wt[j + 2] *
                                   src.ptr<uchar>(shim_reflect101(2 * y + i, src.rows))[shim_reflect101(2 * x + j, src.cols) * cn + k];
                    out.ptr<uchar>(y)[x * cn + k] = (uchar)((acc + 128) >> 8);
                }
        dst = out;
    } else {
        std::cerr << "refshim: pyrDown type " << src.type() << " not implemented\n";
        abort();
    }
}

inline void absdiff(const Mat& a, const Mat& b, Mat& dst) {
    Mat out(a.rows, a.cols, a.type());
    assert(a.depth() == CV_32F && b.type() == a.type());
    for (int r = 0; r < a.rows; ++r)
        for (int c = 0; c < a.cols * a.channels(); ++c) out.ptr<float>(r)[c] = std::fabs(a.ptr<float>(r)[c] - b.ptr<float>(r)[c]);
    dst = out;
}
inline Scalar mean(const Mat& m) {
    Scalar s;
    const int cn = m.channels();
    for (int r = 0; r < m.rows; ++r)
        for (int c = 0; c < m.cols; ++c)
            for (int k = 0; k < cn; ++k) {
                switch (m.depth()) {
                    case CV_8U: s.val[k] += m.ptr<uchar>(r)[c * cn + k]; break;
                    case CV_16U: s.val[k] += m.ptr<ushort>(r)[c * cn + k]; break;
                    case CV_32F: s.val[k] += m.ptr<float>(r)[c * cn + k]; break;
                    default: abort();
                }
            }
    const double n = (double)m.rows * m.cols;
    for (int k = 0; k < 4; ++k) s.val[k] = n > 0 ? s.val[k] / n : 0;
    return s;
}

enum { BORDER_DEFAULT = 4, BORDER_REFLECT_101 = 4, WINDOW_AUTOSIZE = 1 };
inline void shim_unsupported(const char* what) { std::cerr << "refshim: cv::" << what << " is not implemented\n"; abort(); }
inline void Sobel(const Mat&, Mat&, int, int, int, int = 3, double = 1, double = 0, int = 4) { shim_unsupported("Sobel"); }
inline void Scharr(const Mat&, Mat&, int, int, int, double = 1, double = 0, int = 4) { shim_unsupported("Scharr"); }
inline void imshow(const std::string&, const Mat&) {}
inline int waitKey(int = 0) { return -1; }
inline void namedWindow(const std::string&, int = 1) {}
inline void destroyWindow(const std::string&) {}
inline bool imwrite(const std::string&, const Mat&) { return true; }

class TickMeter {
    double t0, acc;
    static double now() { timeval tv; gettimeofday(&tv, 0); return tv.tv_sec + 1e-6 * tv.tv_usec; }
public:
    TickMeter() : t0(0), acc(0) {}
    void start() { t0 = now(); }
    void stop() { acc += now() - t0; }
    double getTimeSec() const { return acc; }
    double getTimeMilli() const { return acc * 1e3; }
    void reset() { acc = 0; }
};

}  // namespace cv
