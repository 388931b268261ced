// rpi_oracle.cpp -- CPU restatement of RegisterPhotoICP's spherical dense registration.
//
// *** TEST INFRASTRUCTURE ONLY. ***  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may load this library.  The product
// (rgbd360_b200/, include/) never links, imports or calls it.
//
// PARITY PIN: the reference (EduFdez/rgbd360) ships no tests, golden vectors or recorded outputs for
// this path, and its own build needs Eigen, OpenCV, PCL, MRPT and Boost (none present, no network).
// Its header /root/reference/include/RegisterPhotoICP.h is nevertheless compiled here, unmodified
// and from where it lies, against from-scratch stand-ins for the slice of those libraries it uses
// (oracle/refshim/, oracle/ref_harness.cpp -> oracle/_ref/).  This restatement is BIT-IDENTICAL to
// that build (planes, LUT, numValidPts of every evaluation, iteration counts, Hessian, gradient,
// pose at one OpenMP thread; tests/test_reference.py, recorded in
// tests/golden/reference_outputs.json) on synthetic pairs and on the reference's own sample pair.
// What stays a restatement on BOTH sides is the third-party arithmetic itself: OpenCV
// cvtColor/convertTo/pyrDown (pinned against python cv2 4.13 golden vectors, tests/test_oracle.py),
// Eigen fixed-size products / norm / 6x6 inverse, MRPT rank() and CPose3D::exp (unpinned: no
// version is named upstream and the libraries are absent).  Citations "RPI.h:N" =
// /root/reference/include/RegisterPhotoICP.h.
//
// Two arithmetic modes (orc_set_math):
//   0 PINNED : asin/atan2/sin/cos come from rgbd360_b200/csrc/sphere_math.h -- the exact
//              operation sequences the GPU kernels execute (bit-exact index maps).
//   1 LIBM   : the same code with glibc asinf/atan2f/sinf/cosf/sin/cos, i.e. what a g++/x86-64
//              build of the reference would call.  The two modes differ ONLY in those calls.
// Two accumulation modes for calcHessGrad_sphere:
//   0 FAITHFUL : as the reference -- per-pixel Jacobian rows stored in N x 6 column-major heap
//                matrices, then two more OpenMP passes with 27 float accumulators
//                (RPI.h:2761-2767, 3117-3224).  This is the mode timed as "reference CPU path".
//   1 STABLE   : same per-pixel floats, products accumulated in double in a single pass.
//
// Build: see oracle/Makefile (g++ -O3 -fopenmp -ffp-contract=off -mfma).
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <chrono>
#include "../include/r360.h"
#include "../rgbd360_b200/csrc/sphere_math.h"
#include "../rgbd360_b200/csrc/gn_math.h"
#include "../rgbd360_b200/csrc/synth.h"
#include "../rgbd360_b200/csrc/stitch_math.h"

#define ORC_INVALID_POINT (-10000.0f)   // RPI.h:40

namespace {

struct MathPinned {
    static float asin_(float x) { return r360_asinf(x); }
    static float atan2_(float y, float x) { return r360_atan2f(y, x); }
    static float sin_(float x) { return r360_sinf(x); }
    static float cos_(float x) { return r360_cosf(x); }
    static void sincos_d(double x, double* s, double* c) { r360_sincos(x, s, c); }
};
struct MathLibm {
    static float asin_(float x) { return asinf(x); }
    static float atan2_(float y, float x) { return atan2f(y, x); }
    static float sin_(float x) { return sinf(x); }
    static float cos_(float x) { return cosf(x); }
    static void sincos_d(double x, double* s, double* c) { *s = sin(x); *c = cos(x); }
};

int g_math_mode = 0;

struct Frame {
    int rows = 0, cols = 0, levels = 0;
    bool has_grad = false;
    std::vector<std::vector<float>> gray, depth, ggx, ggy, dgx, dgy;
};

// ------------------------------------------------------------------ a1: gray conversion
// cv::cvtColor(CV_RGB2GRAY) on 8UC3 (RPI.h:485,502) -- fixed-point luma, verified bit-exact
// against cv2 4.13 -- then convertTo(CV_32FC1, 1./255) (RPI.h:486,503) = u8 * (float)(1/255).
void gray_level0(const uint8_t* rgb, int n, std::vector<float>& out) {
    out.resize(n);
    const float scale = (float)(1. / 255);
    for (int i = 0; i < n; ++i) {
        int g = (rgb[3 * i] * 9798 + rgb[3 * i + 1] * 19235 + rgb[3 * i + 2] * 3735 + 16384) >> 15;
        out[i] = (float)g * scale;
    }
}

inline int reflect101(int i, int n) {
    if (i < 0) i = -i;
    if (i >= n) i = 2 * n - 2 - i;
    return i;
}

// a2: cv::pyrDown (RPI.h:303): separable [1 4 6 4 1], BORDER_REFLECT_101, even samples,
// dst = (cols/2, rows/2).  Horizontal pass is OpenCV's scalar formula, vertical pass its SSE
// formula (OpenCV 2.4: (r0+r4) + (r2+r2) + 4*((r1+r3)+r2), then *1/256).
void pyr_down(const std::vector<float>& src, int rows, int cols, std::vector<float>& dst) {
    const int h = rows / 2, w = cols / 2;
    std::vector<float> hbuf((size_t)rows * w);
#pragma omp parallel for
    for (int r = 0; r < rows; ++r) {
        const float* s = &src[(size_t)r * cols];
        for (int x = 0; x < w; ++x) {
            int c0 = reflect101(2 * x - 2, cols), c1 = reflect101(2 * x - 1, cols), c2 = 2 * x,
                c3 = reflect101(2 * x + 1, cols), c4 = reflect101(2 * x + 2, cols);
            hbuf[(size_t)r * w + x] = s[c2] * 6 + (s[c1] + s[c3]) * 4 + s[c0] + s[c4];
        }
    }
    dst.resize((size_t)h * w);
#pragma omp parallel for
    for (int y = 0; y < h; ++y) {
        const float* r0 = &hbuf[(size_t)reflect101(2 * y - 2, rows) * w];
        const float* r1 = &hbuf[(size_t)reflect101(2 * y - 1, rows) * w];
        const float* r2 = &hbuf[(size_t)(2 * y) * w];
        const float* r3 = &hbuf[(size_t)reflect101(2 * y + 1, rows) * w];
        const float* r4 = &hbuf[(size_t)reflect101(2 * y + 2, rows) * w];
        for (int x = 0; x < w; ++x) {
            float a = (r0[x] + r4[x]) + (r2[x] + r2[x]);
            float b = ((r1[x] + r3[x]) + r2[x]) * 4.0f;
            dst[(size_t)y * w + x] = (a + b) * (1.f / 256);
        }
    }
}

// a3: buildPyramidRange level step (RPI.h:322-350): mean of the 2x2 parents inside
// (minDepth, maxDepth), else 0.
void range_down(const std::vector<float>& src, int rows, int cols, float minD, float maxD,
                std::vector<float>& dst) {
    const int h = rows / 2, w = cols / 2;
    dst.assign((size_t)h * w, 0.f);
#pragma omp parallel for
    for (int r = 0; r < 2 * h; r += 2)
        for (int c = 0; c < 2 * w; c += 2) {
            float av = 0.f;
            unsigned n = 0;
            for (int i = 0; i < 2; ++i)
                for (int j = 0; j < 2; ++j) {
                    float z = src[(size_t)(r + i) * cols + c + j];
                    if (z > minD && z < maxD) { av += z; ++n; }
                }
            if (n > 0) dst[(size_t)(r / 2) * w + c / 2] = av / n;
        }
}

// a4: calcGradientXY (RPI.h:365-398).  Serial in the reference; rows are independent, so the
// OpenMP loop here changes nothing but time.
void grad_xy(const std::vector<float>& s, int rows, int cols, std::vector<float>& gx,
             std::vector<float>& gy) {
    gx.assign((size_t)rows * cols, 0.f);
    gy.assign((size_t)rows * cols, 0.f);
#pragma omp parallel for
    for (int r = 1; r < rows - 1; ++r)
        for (int c = 1; c < cols - 1; ++c) {
            const size_t i = (size_t)r * cols + c;
            const float v = s[i], e = s[i + 1], w = s[i - 1], d = s[i + cols], u = s[i - cols];
            if ((v > e && v < w) || (v < e && v > w)) gx[i] = 2.f / (1 / (e - v) + 1 / (v - w));
            if ((v > d && v < u) || (v < d && v > u)) gy[i] = 2.f / (1 / (d - v) + 1 / (v - u));
        }
}

// a5: sensor-joint mask (RPI.h:4537-4549): zero columns k*W/8-1 and k*W/8, k = 1..7.
void joint_mask(std::vector<float>& p, int rows, int cols, int n_sensors) {
    if (n_sensors <= 1) return;
    const int ws = cols / n_sensors;
    for (int k = 1; k < n_sensors; ++k)
        for (int r = 0; r < rows; ++r) {
            p[(size_t)r * cols + k * ws - 1] = 0.f;
            p[(size_t)r * cols + k * ws] = 0.f;
        }
}

Frame* build_frame(const uint8_t* rgb, const uint16_t* depth_mm, const float* depth_m, int rows,
                   int cols, const r360_params* P, int with_grad) {
    Frame* f = new Frame;
    f->rows = rows; f->cols = cols; f->levels = P->n_levels; f->has_grad = with_grad != 0;
    const int L = P->n_levels;
    f->gray.resize(L); f->depth.resize(L);
    gray_level0(rgb, rows * cols, f->gray[0]);
    f->depth[0].resize((size_t)rows * cols);
    if (depth_mm) {
        const float scale = (float)0.001;            // convertTo(CV_32FC1, 0.001) RPI.h:317
        for (int i = 0; i < rows * cols; ++i) f->depth[0][i] = (float)depth_mm[i] * scale;
    } else {
        memcpy(f->depth[0].data(), depth_m, sizeof(float) * rows * cols);   // RPI.h:319
    }
    for (int l = 1; l < L; ++l) {
        pyr_down(f->gray[l - 1], rows >> (l - 1), cols >> (l - 1), f->gray[l]);
        range_down(f->depth[l - 1], rows >> (l - 1), cols >> (l - 1), P->min_depth, P->max_depth,
                   f->depth[l]);
    }
    if (with_grad) {
        f->ggx.resize(L); f->ggy.resize(L); f->dgx.resize(L); f->dgy.resize(L);
        for (int l = 0; l < L; ++l) {
            const int r = rows >> l, c = cols >> l;
            grad_xy(f->gray[l], r, c, f->ggx[l], f->ggy[l]);     // RPI.h:448
            grad_xy(f->depth[l], r, c, f->dgx[l], f->dgy[l]);    // RPI.h:450
            joint_mask(f->ggx[l], r, c, P->n_sensors_mask);
            joint_mask(f->ggy[l], r, c, P->n_sensors_mask);
            joint_mask(f->dgx[l], r, c, P->n_sensors_mask);
            joint_mask(f->dgy[l], r, c, P->n_sensors_mask);
        }
    }
    return f;
}

// ------------------------------------------------------------------ level constants + LUT (a6)
struct LevelK {
    int rows, cols;
    float res, res_inv, half_rows;
};
LevelK level_consts(const Frame* f, int level) {
    LevelK k;
    k.rows = f->rows >> level;
    k.cols = f->cols >> level;
    k.res = (float)(2 * R360_PI_D / k.cols);          // RPI.h:2553
    k.res_inv = 1 / k.res;                            // RPI.h:2554
    k.half_rows = (float)(0.5 * k.rows - 0.5);        // RPI.h:2556
    return k;
}

// LUT_xyz_sphere (RPI.h:4553-4587), 3 floats per pixel, x = INVALID_POINT when out of range.
template <class M>
void build_lut(const Frame* src, int level, const r360_params* P, std::vector<float>& lut) {
    const LevelK k = level_consts(src, level);
    std::vector<float> st(k.cols), ct(k.cols);
    for (int c = 0; c < k.cols; ++c) {
        float theta = c * k.res;
        st[c] = M::sin_(theta);
        ct[c] = M::cos_(theta);
    }
    lut.resize((size_t)3 * k.rows * k.cols);
    const std::vector<float>& D = src->depth[level];
#pragma omp parallel for
    for (int r = 0; r < k.rows; ++r) {
        float phi = (k.half_rows - r) * k.res;
        float sp = M::sin_(phi), cp = M::cos_(phi);
        for (int c = 0; c < k.cols; ++c) {
            const size_t i = (size_t)r * k.cols + c;
            float d = D[i];
            if (P->min_depth < d && d < P->max_depth) {
                lut[3 * i + 0] = d * sp;
                lut[3 * i + 1] = -d * cp * st[c];
                lut[3 * i + 2] = -d * cp * ct[c];
            } else {
                lut[3 * i + 0] = ORC_INVALID_POINT;
            }
        }
    }
}

// ------------------------------------------------------------------ per-pixel warp (a8/a9 core)
struct Warped {
    float px, py, pz, dist, dinv;
    int r, c;
};
// RPI.h:2672-2680 == 2973-2981.  Eigen 3x3*3x1 coefficient product sums left to right, then
// + translation; norm() = sqrt(x^2 + (y^2 + z^2)) (Eigen's unrolled redux splits 3 as 1 + 2).
template <class M>
inline bool warp_point(const float* T, const float* X, const LevelK& k, Warped& w) {
    w.px = ((T[0] * X[0] + T[4] * X[1]) + T[8] * X[2]) + T[12];
    w.py = ((T[1] * X[0] + T[5] * X[1]) + T[9] * X[2]) + T[13];
    w.pz = ((T[2] * X[0] + T[6] * X[1]) + T[10] * X[2]) + T[14];
    w.dist = sqrtf(w.px * w.px + (w.py * w.py + w.pz * w.pz));
    w.dinv = 1.f / w.dist;
    float phi = M::asin_(w.px * w.dinv);
    float theta = (float)((double)M::atan2_(w.py, w.pz) + R360_PI_D);
    w.r = r360_round_to_int(k.half_rows - phi * k.res_inv);
    w.c = r360_round_to_int(theta * k.res_inv);
    return (w.r >= 0 && w.r < k.rows) && w.c < k.cols;        // RPI.h:2683 (no c >= 0 test)
}

// ------------------------------------------------------------------ a8 errorPhotoICP_sphere
template <class M>
void error_sphere(const Frame* src, const Frame* trg, int level, const float* T,
                  const r360_params* P, const std::vector<float>& lut, double* err2_out,
                  int* nvalid_out) {
    const LevelK k = level_consts(src, level);
    const int N = k.rows * k.cols;
    const double stdDevPhoto_inv = 1. / P->std_photo;          // RPI.h:2561 (double)
    const float* Is = src->gray[level].data();
    const float* It = trg->gray[level].data();
    const float* Dt = trg->depth[level].data();
    const float* Ix = trg->ggx[level].data();
    const float* Iy = trg->ggy[level].data();
    const float* Dx = trg->dgx[level].data();
    const float* Dy = trg->dgy[level].data();
    const int method = P->method;
    double error2 = 0.0;
    int numValidPts = 0;
#pragma omp parallel for reduction(+ : error2, numValidPts)
    for (int i = 0; i < N; ++i) {
        if (lut[3 * (size_t)i] == ORC_INVALID_POINT) continue;
        Warped w;
        if (!warp_point<M>(T, &lut[3 * (size_t)i], k, w)) continue;
        const size_t j = (size_t)w.r * k.cols + w.c;
        if (method == R360_PHOTO_CONSISTENCY || method == R360_PHOTO_DEPTH) {
            if (fabsf(Ix[j]) < P->thres_sal_int && fabsf(Iy[j]) < P->thres_sal_int) continue;
            float photoDiff = It[j] - Is[i];
            double weight_photo = r360_huber(photoDiff, P->std_photo) * stdDevPhoto_inv;
            float werr = (float)(weight_photo * photoDiff);
            error2 += werr * werr;
            ++numValidPts;
        }
        if (method == R360_DEPTH_CONSISTENCY || method == R360_PHOTO_DEPTH) {
            float depth2 = Dt[j];
            if (std::isfinite(depth2)) {
                if (fabsf(Dx[j]) < P->thres_sal_depth && fabsf(Dy[j]) < P->thres_sal_depth) continue;
                float depthDiff = depth2 - w.dist;
                float sd = P->std_depth * depth2;
                double weight_depth = r360_huber(depthDiff, sd) / sd;
                float werr = (float)(weight_depth * depthDiff);
                error2 += werr * werr;
                ++numValidPts;
            }
        }
    }
    *err2_out = error2;
    *nvalid_out = numValidPts;
}

// ------------------------------------------------------------------ a9 per-pixel Jacobian rows
// RPI.h:2991-3088.  Returns bit 0: photo row valid, bit 1: depth row valid.
inline int jacobian_rows(const Warped& w, const LevelK& k, const r360_params* P, float Is,
                         float It, float Dt, float Ix, float Iy, float Dx, float Dy,
                         float stdDevPhoto_inv, float* Jp, float* rp, float* Jd, float* rd,
                         bool occ_store = false) {
    // occ_store: calcHessGrad_sphereOcc1/Occ2 assign the rows in a block AFTER both terms
    // (RPI.h:3573-3599, 4095-4121), so the depth saliency `continue` (RPI.h:3556-3557, 4078-4079)
    // drops the photo row of that pixel as well.
    const int method = P->method;
    const float x = w.px, y = w.py, z = w.pz;
    // jacobianProj23
    float z_inv = 1.f / z;
    float z_inv2 = z_inv * z_inv;
    float D_atan = 1.f / (1 + y * y * z_inv2) * k.res_inv;
    float P01 = D_atan * z_inv;
    float P02 = -y * z_inv2 * D_atan;
    float dinv2 = w.dinv * w.dinv;
    float xd = x * dinv2;
    float D_asin = 1.f / sqrtf(1 - x * xd) * k.res_inv;
    float P10 = -D_asin * w.dinv * (1 - x * xd);
    float P11 = D_asin * (xd * y * w.dinv);
    float P12 = D_asin * (xd * z * w.dinv);
    // jacobianWarpRt = jacobianProj23 * [I | -skew(p)], coefficient sums k = 0,1,2 left to right
    float Jw0[6], Jw1[6];
    const float P00 = 0.f;
    Jw0[0] = (P00 * 1 + P01 * 0) + P02 * 0;  Jw0[1] = (P00 * 0 + P01 * 1) + P02 * 0;
    Jw0[2] = (P00 * 0 + P01 * 0) + P02 * 1;
    Jw0[3] = (P00 * 0 + P01 * (-z)) + P02 * y;
    Jw0[4] = (P00 * z + P01 * 0) + P02 * (-x);
    Jw0[5] = (P00 * (-y) + P01 * x) + P02 * 0;
    Jw1[0] = (P10 * 1 + P11 * 0) + P12 * 0;  Jw1[1] = (P10 * 0 + P11 * 1) + P12 * 0;
    Jw1[2] = (P10 * 0 + P11 * 0) + P12 * 1;
    Jw1[3] = (P10 * 0 + P11 * (-z)) + P12 * y;
    Jw1[4] = (P10 * z + P11 * 0) + P12 * (-x);
    Jw1[5] = (P10 * (-y) + P11 * x) + P12 * 0;
    int valid = 0;
    if (method == R360_PHOTO_CONSISTENCY || method == R360_PHOTO_DEPTH) {
        if (fabsf(Ix) < P->thres_sal_int && fabsf(Iy) < P->thres_sal_int) return 0;   // `continue`
        float photoDiff = It - Is;
        float weight_photo = r360_huber(photoDiff, P->std_photo) * stdDevPhoto_inv;
        *rp = weight_photo * photoDiff;
        float a = weight_photo * Ix, b = weight_photo * Iy;
        for (int q = 0; q < 6; ++q) Jp[q] = a * Jw0[q] + b * Jw1[q];
        valid |= 1;
    }
    if (method == R360_DEPTH_CONSISTENCY || method == R360_PHOTO_DEPTH) {
        if (std::isfinite(Dt)) {
            if (fabsf(Dx) < P->thres_sal_depth && fabsf(Dy) < P->thres_sal_depth) return occ_store ? 0 : valid;
            float depthDiff = Dt - w.dist;
            float sd = P->std_depth * Dt;
            float weight_depth = r360_huber(depthDiff, sd) / sd;
            *rd = weight_depth * depthDiff;
            const float n0 = x * w.dinv, n1 = y * w.dinv, n2 = z * w.dinv;
            float nT[6];
            nT[0] = (n0 * 1 + n1 * 0) + n2 * 0;  nT[1] = (n0 * 0 + n1 * 1) + n2 * 0;
            nT[2] = (n0 * 0 + n1 * 0) + n2 * 1;
            nT[3] = (n0 * 0 + n1 * (-z)) + n2 * y;
            nT[4] = (n0 * z + n1 * 0) + n2 * (-x);
            nT[5] = (n0 * (-y) + n1 * x) + n2 * 0;
            for (int q = 0; q < 6; ++q) Jd[q] = weight_depth * ((Dx * Jw0[q] + Dy * Jw1[q]) - nT[q]);
            valid |= 2;
        }
    }
    return valid;
}

struct HessOut {
    float H[36];
    float g[6];
    double Hd[21], gd[6];
    int n_visible;
    int n_photo, n_depth;
};

template <class M>
void hessgrad_sphere(const Frame* src, const Frame* trg, int level, const float* T,
                     const r360_params* P, const std::vector<float>& lut, int accum_mode,
                     HessOut* out) {
    const LevelK k = level_consts(src, level);
    const int N = k.rows * k.cols;
    const float stdDevPhoto_inv = (float)(1. / P->std_photo);      // RPI.h:2774 (float)
    const float* Is = src->gray[level].data();
    const float* It = trg->gray[level].data();
    const float* Dt = trg->depth[level].data();
    const float* Ix = trg->ggx[level].data();
    const float* Iy = trg->ggy[level].data();
    const float* Dx = trg->dgx[level].data();
    const float* Dy = trg->dgy[level].data();
    int numVisible = 0, nPhoto = 0, nDepth = 0;
    double acc[27];
    for (int q = 0; q < 27; ++q) acc[q] = 0.0;

    if (accum_mode == 0) {
        // FAITHFUL: heap Jacobian store (column-major N x 6), then float reductions.
        float* JP = (float*)malloc(sizeof(float) * 6 * (size_t)N);       // MatrixXf(imgSize,6), uninitialised
        float* JD = (float*)malloc(sizeof(float) * 6 * (size_t)N);
        float* RP = (float*)calloc(N, sizeof(float));
        float* RD = (float*)calloc(N, sizeof(float));
        int* VP = (int*)calloc(N, sizeof(int));
        int* VD = (int*)calloc(N, sizeof(int));
        float* ZB = (float*)calloc(N, sizeof(float));                    // invDepthBuffer (unused, RPI.h:2767)
#pragma omp parallel for reduction(+ : numVisible)
        for (int i = 0; i < N; ++i) {
            if (lut[3 * (size_t)i] == ORC_INVALID_POINT) continue;
            Warped w;
            if (!warp_point<M>(T, &lut[3 * (size_t)i], k, w)) continue;
            ++numVisible;
            const size_t j = (size_t)w.r * k.cols + w.c;
            float Jp[6], Jd[6], rp = 0.f, rd = 0.f;
            int v = jacobian_rows(w, k, P, Is[i], It[j], Dt[j], Ix[j], Iy[j], Dx[j], Dy[j],
                                  stdDevPhoto_inv, Jp, &rp, Jd, &rd);
            if (v & 1) { for (int q = 0; q < 6; ++q) JP[(size_t)q * N + i] = Jp[q]; RP[i] = rp; VP[i] = 1; }
            if (v & 2) { for (int q = 0; q < 6; ++q) JD[(size_t)q * N + i] = Jd[q]; RD[i] = rd; VD[i] = 1; }
        }
        float h11 = 0, h12 = 0, h13 = 0, h14 = 0, h15 = 0, h16 = 0, h22 = 0, h23 = 0, h24 = 0, h25 = 0,
              h26 = 0, h33 = 0, h34 = 0, h35 = 0, h36 = 0, h44 = 0, h45 = 0, h46 = 0, h55 = 0, h56 = 0,
              h66 = 0, g1 = 0, g2 = 0, g3 = 0, g4 = 0, g5 = 0, g6 = 0;
        for (int pass = 0; pass < 2; ++pass) {
            const float* J = pass == 0 ? JP : JD;
            const float* R = pass == 0 ? RP : RD;
            const int* V = pass == 0 ? VP : VD;
            const bool on = pass == 0 ? (P->method != R360_DEPTH_CONSISTENCY)
                                      : (P->method != R360_PHOTO_CONSISTENCY);
            if (!on) continue;
            int cnt = 0;
#pragma omp parallel for reduction(+ : h11, h12, h13, h14, h15, h16, h22, h23, h24, h25, h26, h33, h34, h35, h36, h44, h45, h46, h55, h56, h66, g1, g2, g3, g4, g5, g6, cnt)
            for (int i = 0; i < N; ++i)
                if (V[i]) {
                    const float j0 = J[i], j1 = J[(size_t)N + i], j2 = J[2 * (size_t)N + i],
                                j3 = J[3 * (size_t)N + i], j4 = J[4 * (size_t)N + i], j5 = J[5 * (size_t)N + i];
                    const float r = R[i];
                    h11 += j0 * j0; h12 += j0 * j1; h13 += j0 * j2; h14 += j0 * j3; h15 += j0 * j4; h16 += j0 * j5;
                    h22 += j1 * j1; h23 += j1 * j2; h24 += j1 * j3; h25 += j1 * j4; h26 += j1 * j5;
                    h33 += j2 * j2; h34 += j2 * j3; h35 += j2 * j4; h36 += j2 * j5;
                    h44 += j3 * j3; h45 += j3 * j4; h46 += j3 * j5;
                    h55 += j4 * j4; h56 += j4 * j5; h66 += j5 * j5;
                    g1 += j0 * r; g2 += j1 * r; g3 += j2 * r; g4 += j3 * r; g5 += j4 * r; g6 += j5 * r;
                    ++cnt;
                }
            if (pass == 0) nPhoto = cnt; else nDepth = cnt;
        }
        const float hv[21] = { h11, h12, h13, h14, h15, h16, h22, h23, h24, h25, h26, h33, h34, h35,
                               h36, h44, h45, h46, h55, h56, h66 };
        const float gv[6] = { g1, g2, g3, g4, g5, g6 };
        for (int q = 0; q < 21; ++q) acc[q] = hv[q];
        for (int q = 0; q < 6; ++q) acc[21 + q] = gv[q];
        free(JP); free(JD); free(RP); free(RD); free(VP); free(VD); free(ZB);
    } else {
        // STABLE: same per-pixel float rows and float products, double accumulation.
        double a0 = 0, a1 = 0, a2 = 0, a3 = 0, a4 = 0, a5 = 0, a6 = 0, a7 = 0, a8 = 0, a9 = 0, a10 = 0,
               a11 = 0, a12 = 0, a13 = 0, a14 = 0, a15 = 0, a16 = 0, a17 = 0, a18 = 0, a19 = 0, a20 = 0,
               b0 = 0, b1 = 0, b2 = 0, b3 = 0, b4 = 0, b5 = 0;
#pragma omp parallel for reduction(+ : numVisible, nPhoto, nDepth, a0, a1, a2, a3, a4, a5, a6, a7, a8, a9, a10, a11, a12, a13, a14, a15, a16, a17, a18, a19, a20, b0, b1, b2, b3, b4, b5)
        for (int i = 0; i < N; ++i) {
            if (lut[3 * (size_t)i] == ORC_INVALID_POINT) continue;
            Warped w;
            if (!warp_point<M>(T, &lut[3 * (size_t)i], k, w)) continue;
            ++numVisible;
            const size_t j = (size_t)w.r * k.cols + w.c;
            float Jp[6], Jd[6], rp = 0.f, rd = 0.f;
            int v = jacobian_rows(w, k, P, Is[i], It[j], Dt[j], Ix[j], Iy[j], Dx[j], Dy[j],
                                  stdDevPhoto_inv, Jp, &rp, Jd, &rd);
            for (int pass = 0; pass < 2; ++pass) {
                if (!(v & (1 << pass))) continue;
                const float* J = pass == 0 ? Jp : Jd;
                const float r = pass == 0 ? rp : rd;
                if (pass == 0) ++nPhoto; else ++nDepth;
                a0 += J[0] * J[0]; a1 += J[0] * J[1]; a2 += J[0] * J[2]; a3 += J[0] * J[3]; a4 += J[0] * J[4]; a5 += J[0] * J[5];
                a6 += J[1] * J[1]; a7 += J[1] * J[2]; a8 += J[1] * J[3]; a9 += J[1] * J[4]; a10 += J[1] * J[5];
                a11 += J[2] * J[2]; a12 += J[2] * J[3]; a13 += J[2] * J[4]; a14 += J[2] * J[5];
                a15 += J[3] * J[3]; a16 += J[3] * J[4]; a17 += J[3] * J[5];
                a18 += J[4] * J[4]; a19 += J[4] * J[5]; a20 += J[5] * J[5];
                b0 += J[0] * r; b1 += J[1] * r; b2 += J[2] * r; b3 += J[3] * r; b4 += J[4] * r; b5 += J[5] * r;
            }
        }
        const double av[27] = { a0, a1, a2, a3, a4, a5, a6, a7, a8, a9, a10, a11, a12, a13, a14, a15, a16,
                                a17, a18, a19, a20, b0, b1, b2, b3, b4, b5 };
        for (int q = 0; q < 27; ++q) acc[q] = av[q];
    }
    // RPI.h:3197-3224: symmetric fill
    int q = 0;
    for (int a = 0; a < 6; ++a)
        for (int b = a; b < 6; ++b, ++q) {
            out->Hd[q] = acc[q];
            out->H[a + 6 * b] = out->H[b + 6 * a] = (float)acc[q];
        }
    for (int a = 0; a < 6; ++a) { out->gd[a] = acc[21 + a]; out->g[a] = (float)acc[21 + a]; }
    out->n_visible = numVisible;
    out->n_photo = nPhoto;
    out->n_depth = nDepth;
}


// ------------------------------------------------------------------ f2: occlusion variants
// errorPhotoICP_sphereOcc1 / Occ2 (RPI.h:3232-3369, 3720-3858) and calcHessGrad_sphereOcc1 / Occ2
// (RPI.h:3373-3716, 3861-4249).  Upstream these loops carry `#pragma omp parallel for` over a
// z-buffer that every iteration reads and writes, so their result depends on the thread schedule;
// the semantics restated here (and implemented on the GPU) are those of ONE thread, i.e. source
// pixels visited in index order -- what the compiled reference (oracle/_ref) computes with
// omp_set_num_threads(1).  The loops below are therefore serial on purpose.
//   Occ1 error : z-buffer per TARGET texel; a pixel passes iff its 1/|p| is >= every earlier
//                candidate's of the same texel; squared residuals are stored per target texel (the
//                last passer wins), the counters count every passer.
//   Occ1 H/g   : the z-buffer is indexed with the SOURCE index (RPI.h:3473-3475), so it never
//                rejects anything; rows are assigned after both terms (occ_store above).
//   Occ2 error : |depth2 - dist| > thresDepthOutliers (0.3, RPI.h:4525) rejects first; then the
//                Occ1 z-test; residuals per SOURCE pixel; both RMS use nValidDepthPts = #passers.
//   Occ2 H/g   : outlier gate, no z-test; rows stored per TARGET texel (the last candidate in source
//                order wins); numVisiblePixels counts distinct target texels.
struct OccErr {
    double photo, depth;      // PhotoResidual, DepthResidual
    int n_photo, n_depth;     // nValidPhotoPts, nValidDepthPts
};
inline double occ_error_value(const OccErr& e, int occ) {
    // RPI.h:3360-3367 (Occ1: each RMS over its own count) / 3849-3856 (Occ2: both over nValidDepthPts)
    const double avP = sqrt(e.photo / (occ == 1 ? e.n_photo : e.n_depth));
    const double avD = sqrt(e.depth / e.n_depth);
    return avP + avD;
}

template <class M>
void error_sphere_occ(const Frame* src, const Frame* trg, int level, const float* T, const r360_params* P,
                      const std::vector<float>& lut, int occ, OccErr* out) {
    const LevelK k = level_consts(src, level);
    const int N = k.rows * k.cols;
    const double stdDevPhoto_inv = 1. / P->std_photo;          // RPI.h:3256 (double)
    const float* Is = src->gray[level].data();
    const float* It = trg->gray[level].data();
    const float* Dt = trg->depth[level].data();
    const float* Ix = trg->ggx[level].data();
    const float* Iy = trg->ggy[level].data();
    const float* Dx = trg->dgx[level].data();
    const float* Dy = trg->dgy[level].data();
    const int method = P->method;
    const float thresDepthOutliers = (float)0.3;               // RPI.h:4525
    std::vector<float> resP(N, 0.f), resD(N, 0.f), zbuf(N, 0.f);
    int nP = 0, nD = 0;
    for (int i = 0; i < N; ++i) {
        if (lut[3 * (size_t)i] == ORC_INVALID_POINT) continue;
        Warped w;
        if (!warp_point<M>(T, &lut[3 * (size_t)i], k, w)) continue;
        const size_t j = (size_t)w.r * k.cols + w.c;
        float depth2 = Dt[j];
        float depthDiff = depth2 - w.dist;
        if (occ == 2 && fabsf(depthDiff) > thresDepthOutliers) continue;       // RPI.h:3788-3791
        if (zbuf[j] > 0 && w.dinv < zbuf[j]) continue;                          // RPI.h:3299 / 3794
        zbuf[j] = w.dinv;
        if (occ == 2) ++nD;                                                     // RPI.h:3798
        const size_t slot = occ == 1 ? j : (size_t)i;                           // RPI.h:3318 (ii) / 3814 (i)
        if (method == R360_PHOTO_CONSISTENCY || method == R360_PHOTO_DEPTH) {
            if (fabsf(Ix[j]) < P->thres_sal_int && fabsf(Iy[j]) < P->thres_sal_int) continue;
            float photoDiff = It[j] - Is[i];
            double weight_photo = r360_huber(photoDiff, P->std_photo) * stdDevPhoto_inv;
            float werr = (float)(weight_photo * photoDiff);
            resP[slot] = werr * werr;
            if (occ == 1) ++nP;
        }
        if (method == R360_DEPTH_CONSISTENCY || method == R360_PHOTO_DEPTH) {
            if (std::isfinite(depth2)) {
                if (fabsf(Dx[j]) < P->thres_sal_depth && fabsf(Dy[j]) < P->thres_sal_depth) continue;
                float sd = P->std_depth * depth2;
                double weight_depth = r360_huber(depthDiff, sd) / sd;
                float werr = (float)(weight_depth * depthDiff);
                resD[slot] = werr * werr;
                if (occ == 1) ++nD;
            }
        }
    }
    double PhotoResidual = 0.0, DepthResidual = 0.0;
    for (int i = 0; i < N; ++i) { PhotoResidual += resP[i]; DepthResidual += resD[i]; }
    out->photo = PhotoResidual; out->depth = DepthResidual; out->n_photo = nP; out->n_depth = nD;
}

template <class M>
void hessgrad_sphere_occ(const Frame* src, const Frame* trg, int level, const float* T, const r360_params* P,
                         const std::vector<float>& lut, int accum_mode, int occ, HessOut* out) {
    const LevelK k = level_consts(src, level);
    const int N = k.rows * k.cols;
    const float stdDevPhoto_inv = (float)(1. / P->std_photo);      // RPI.h:3402 (float)
    const float thresDepthOutliers = (float)0.3;
    const float* Is = src->gray[level].data();
    const float* It = trg->gray[level].data();
    const float* Dt = trg->depth[level].data();
    const float* Ix = trg->ggx[level].data();
    const float* Iy = trg->ggy[level].data();
    const float* Dx = trg->dgx[level].data();
    const float* Dy = trg->dgy[level].data();
    std::vector<float> JP(6 * (size_t)N), JD(6 * (size_t)N), RP(N, 0.f), RD(N, 0.f), zbuf(N, 0.f);
    std::vector<int> VP(N, 0), VD(N, 0);
    int numVisible = 0;
    for (int i = 0; i < N; ++i) {
        if (lut[3 * (size_t)i] == ORC_INVALID_POINT) continue;
        Warped w;
        if (!warp_point<M>(T, &lut[3 * (size_t)i], k, w)) continue;
        const size_t j = (size_t)w.r * k.cols + w.c;
        size_t slot;
        if (occ == 1) {
            // invDepthBuffer(i): first and only visit of source index i, never occluded (RPI.h:3473-3475)
            ++numVisible;
            slot = (size_t)i;
        } else {
            if (fabsf(Dt[j] - w.dist) > thresDepthOutliers) continue;           // RPI.h:3968-3981
            if (zbuf[j] == 0) ++numVisible;                                     // RPI.h:3984-3985
            zbuf[j] = w.dinv;
            slot = j;
        }
        float Jp[6], Jd[6], rp = 0.f, rd = 0.f;
        const int v = jacobian_rows(w, k, P, Is[i], It[j], Dt[j], Ix[j], Iy[j], Dx[j], Dy[j], stdDevPhoto_inv,
                                    Jp, &rp, Jd, &rd, true);
        if (v & 1) { for (int q = 0; q < 6; ++q) JP[(size_t)q * N + slot] = Jp[q]; RP[slot] = rp; VP[slot] = 1; }
        if (v & 2) { for (int q = 0; q < 6; ++q) JD[(size_t)q * N + slot] = Jd[q]; RD[slot] = rd; VD[slot] = 1; }
    }
    double acc[27];
    int nPhoto = 0, nDepth = 0;
    float hf[27];
    double hd[27];
    for (int q = 0; q < 27; ++q) { hf[q] = 0.f; hd[q] = 0.0; }
    for (int pass = 0; pass < 2; ++pass) {
        const std::vector<float>& J = pass == 0 ? JP : JD;
        const std::vector<float>& R = pass == 0 ? RP : RD;
        const std::vector<int>& V = pass == 0 ? VP : VD;
        const bool on = pass == 0 ? (P->method != R360_DEPTH_CONSISTENCY) : (P->method != R360_PHOTO_CONSISTENCY);
        if (!on) continue;
        // an OpenMP `reduction(+: h11, ...)` sums into zero-initialised private copies and adds them to
        // the shared variables at the end of the loop -- also with one thread (RPI.h:3611, 3649)
        float pf[27];
        for (int q = 0; q < 27; ++q) pf[q] = 0.f;
        for (int i = 0; i < N; ++i)
            if (V[i]) {
                float jj[6];
                for (int q = 0; q < 6; ++q) jj[q] = J[(size_t)q * N + i];
                const float r = R[i];
                int q = 0;
                for (int a = 0; a < 6; ++a)
                    for (int b = a; b < 6; ++b, ++q) {
                        const float pr = jj[a] * jj[b];
                        pf[q] += pr; hd[q] += pr;
                    }
                for (int a = 0; a < 6; ++a) { const float pr = jj[a] * r; pf[21 + a] += pr; hd[21 + a] += pr; }
                if (pass == 0) ++nPhoto; else ++nDepth;
            }
        for (int q = 0; q < 27; ++q) hf[q] += pf[q];
    }
    for (int q = 0; q < 27; ++q) acc[q] = accum_mode == 0 ? (double)hf[q] : hd[q];
    int q = 0;
    for (int a = 0; a < 6; ++a)
        for (int b = a; b < 6; ++b, ++q) {
            out->Hd[q] = acc[q];
            out->H[a + 6 * b] = out->H[b + 6 * a] = (float)acc[q];
        }
    for (int a = 0; a < 6; ++a) { out->gd[a] = acc[21 + a]; out->g[a] = (float)acc[21 + a]; }
    out->n_visible = numVisible;
    out->n_photo = nPhoto;
    out->n_depth = nDepth;
}

// ------------------------------------------------------------------ f4: pinhole path
// errorPhotoICP (RPI.h:560-775), calcHessGrad (RPI.h:776-1104) and alignFrames (RPI.h:4254-4512) with
// occlusion = 0 and bUseSalientPixels = false (the default, RPI.h:205): the same pyramids as the spherical
// path (no sensor-joint mask: that lives in alignFrames360), a pinhole projection with the camera matrix
// of setCameraMatrix (RPI.h:254) scaled per level, and a Levenberg-Marquardt loop.  Differences from
// the spherical functions that the restatement keeps: the ERROR applies no saliency test while the
// HESSIAN does (and its depth-saliency `continue` drops the photo row too); both RMS terms divide by
// nValidDepthPts (photo-only => 0/0 = NaN => the loop never runs); the sum is returned through the
// FLOAT member avResidual; sigma_depth uses the transformed source z, not the target depth; H and g are
// accumulated pixel by pixel in float under `omp critical` (order-dependent: pinned at one thread);
// the full SE(3) exponential; lambda 0.01, step 10, tol_residual 1e-4, accept on diff_error > 0 and one
// damped retry otherwise.
struct PinK {
    int rows, cols;
    float fx, fy, ox, oy, inv_fx, inv_fy;
};
PinK pinhole_consts(const Frame* f, int level, const float* cam /* fx fy ox oy */) {
    PinK k;
    k.rows = f->rows >> level;
    k.cols = f->cols >> level;
    const float scaleFactor = 1.0 / pow(2, level);             // RPI.h:569
    k.fx = cam[0] * scaleFactor; k.fy = cam[1] * scaleFactor;
    k.ox = cam[2] * scaleFactor; k.oy = cam[3] * scaleFactor;
    k.inv_fx = 1. / k.fx; k.inv_fy = 1. / k.fy;                // RPI.h:574-575 (double division, narrowed)
    return k;
}
// LUT_xyz_sphere as alignFrames fills it (RPI.h:4275-4301)
void build_lut_pinhole(const Frame* src, int level, const r360_params* P, const PinK& k, std::vector<float>& lut) {
    lut.resize((size_t)3 * k.rows * k.cols);
    const std::vector<float>& D = src->depth[level];
    for (int r = 0; r < k.rows; ++r)
        for (int c = 0; c < k.cols; ++c) {
            const size_t i = (size_t)r * k.cols + c;
            const float z = D[i];
            lut[3 * i + 2] = z;
            if (P->min_depth < z && z < P->max_depth) {
                lut[3 * i + 0] = (c - k.ox) * z * k.inv_fx;
                lut[3 * i + 1] = (r - k.oy) * z * k.inv_fy;
            } else {
                lut[3 * i + 0] = ORC_INVALID_POINT;
            }
        }
}
struct WarpedPin { float px, py, pz, inv_z; int r, c; };
inline bool warp_point_pinhole(const float* T, const float* X, const PinK& k, WarpedPin& w) {
    w.px = ((T[0] * X[0] + T[4] * X[1]) + T[8] * X[2]) + T[12];
    w.py = ((T[1] * X[0] + T[5] * X[1]) + T[9] * X[2]) + T[13];
    w.pz = ((T[2] * X[0] + T[6] * X[1]) + T[10] * X[2]) + T[14];
    w.inv_z = 1.0 / w.pz;                                      // RPI.h:659: double division, narrowed to float
    const float tc = (w.px * k.fx) * w.inv_z + k.ox;           // RPI.h:662
    const float tr = (w.py * k.fy) * w.inv_z + k.oy;           // RPI.h:663
    w.r = r360_round_to_int(tr);
    w.c = r360_round_to_int(tc);
    return (w.r >= 0 && w.r < k.rows) && (w.c >= 0 && w.c < k.cols);      // RPI.h:667-668
}
// errorPhotoICP: out = {PhotoResidual, DepthResidual, nValidPhotoPts, nValidDepthPts}; returns the value
double error_pinhole(const Frame* src, const Frame* trg, int level, const float* T, const r360_params* P,
                     const PinK& k, const std::vector<float>& lut, OccErr* out) {
    const int N = k.rows * k.cols;
    const float stdDevPhoto_inv = 1. / P->std_photo;            // RPI.h:578 (float)
    const float* Is = src->gray[level].data();
    const float* It = trg->gray[level].data();
    const float* Dt = trg->depth[level].data();
    const int method = P->method;
    double PhotoResidual = 0.0, DepthResidual = 0.0;
    int nP = 0, nD = 0;
    for (int i = 0; i < N; ++i) {
        if (lut[3 * (size_t)i] == ORC_INVALID_POINT) continue;
        WarpedPin w;
        if (!warp_point_pinhole(T, &lut[3 * (size_t)i], k, w)) continue;
        const size_t j = (size_t)w.r * k.cols + w.c;
        if (method == R360_PHOTO_CONSISTENCY || method == R360_PHOTO_DEPTH) {
            const float photoDiff = It[j] - Is[i];
            const float weight_photo = r360_huber(photoDiff, P->std_photo) * stdDevPhoto_inv;
            const float werr = weight_photo * photoDiff;
            PhotoResidual += werr * werr;
            ++nP;
        }
        if (method == R360_DEPTH_CONSISTENCY || method == R360_PHOTO_DEPTH) {
            const float depth2 = Dt[j];
            if (std::isfinite(depth2)) {
                const float depth1 = w.pz;
                const float depthDiff = depth2 - depth1;
                const float sd = P->std_depth * depth1;          // RPI.h:694: the transformed SOURCE depth
                const float weight_depth = r360_huber(depthDiff, sd) / sd;
                const float werr = weight_depth * depthDiff;
                DepthResidual += werr * werr;
                ++nD;
            }
        }
    }
    out->photo = PhotoResidual; out->depth = DepthResidual; out->n_photo = nP; out->n_depth = nD;
    const double avP = sqrt(PhotoResidual / nD), avD = sqrt(DepthResidual / nD);      // RPI.h:768-769
    const float avResidual = avP + avD;                          // float member, RPI.h:183, 770
    return avResidual;
}
void hessgrad_pinhole(const Frame* src, const Frame* trg, int level, const float* T, const r360_params* P,
                      const PinK& k, const std::vector<float>& lut, int accum_mode, HessOut* out) {
    const int N = k.rows * k.cols;
    const float stdDevPhoto_inv = 1. / P->std_photo;
    const float* Is = src->gray[level].data();
    const float* It = trg->gray[level].data();
    const float* Dt = trg->depth[level].data();
    const float* Ix = trg->ggx[level].data();
    const float* Iy = trg->ggy[level].data();
    const float* Dx = trg->dgx[level].data();
    const float* Dy = trg->dgy[level].data();
    const int method = P->method;
    float Hf[36], gf[6];
    double Hd[36], gd[6];
    for (int q = 0; q < 36; ++q) { Hf[q] = 0.f; Hd[q] = 0.0; }
    for (int q = 0; q < 6; ++q) { gf[q] = 0.f; gd[q] = 0.0; }
    int nVisible = 0, nPhoto = 0, nDepth = 0;
    for (int i = 0; i < N; ++i) {
        if (lut[3 * (size_t)i] == ORC_INVALID_POINT) continue;
        WarpedPin w;
        if (!warp_point_pinhole(T, &lut[3 * (size_t)i], k, w)) continue;
        ++nVisible;
        const size_t j = (size_t)w.r * k.cols + w.c;
        const float x = w.px, y = w.py, iz = w.inv_z;
        float J0[6], J1[6];                                      // jacobianWarpRt rows, RPI.h:970-984
        J0[0] = k.fx * iz; J1[0] = 0;
        J0[1] = 0; J1[1] = k.fy * iz;
        const float iz2 = iz * iz;
        J0[2] = -k.fx * x * iz2;
        J1[2] = -k.fy * y * iz2;
        J0[3] = -k.fx * y * x * iz2;
        J1[3] = -k.fy * (1 + y * y * iz2);
        J0[4] = k.fx * (1 + x * x * iz2);
        J1[4] = k.fy * x * y * iz2;
        J0[5] = -k.fx * y * iz;
        J1[5] = k.fy * x * iz;
        float Jp[6], Jd[6], rp = 0.f, rd = 0.f;
        bool have_depth = false;
        if (method == R360_PHOTO_CONSISTENCY || method == R360_PHOTO_DEPTH) {
            if (fabsf(Ix[j]) < P->thres_sal_int && fabsf(Iy[j]) < P->thres_sal_int) continue;
            const float photoDiff = It[j] - Is[i];
            const float weight_photo = r360_huber(photoDiff, P->std_photo) * stdDevPhoto_inv;
            rp = weight_photo * photoDiff;
            const float a = weight_photo * Ix[j], b = weight_photo * Iy[j];
            for (int q = 0; q < 6; ++q) Jp[q] = a * J0[q] + b * J1[q];
        }
        if (method == R360_DEPTH_CONSISTENCY || method == R360_PHOTO_DEPTH) {
            if (fabsf(Dx[j]) < P->thres_sal_depth && fabsf(Dy[j]) < P->thres_sal_depth) continue;   // drops the photo row too
            const float depth2 = Dt[j];
            if (std::isfinite(depth2)) {
                const float depthDiff = depth2 - w.pz;
                const float sd = P->std_depth * w.pz;
                const float weight_depth = r360_huber(depthDiff, sd) / sd;
                rd = weight_depth * depthDiff;
                const float Rz[6] = { 0, 0, 1, y, -x, 0 };      // jacobianRt_z, RPI.h:1053
                for (int q = 0; q < 6; ++q) Jd[q] = weight_depth * ((Dx[j] * J0[q] + Dy[j] * J1[q]) - Rz[q]);
                have_depth = true;
            }
        }
        // hessian += J^T J; gradient += J^T r  (Eigen, float, every pixel in turn: RPI.h:1065-1083)
        for (int pass = 0; pass < 2; ++pass) {
            const bool on = pass == 0 ? (method != R360_DEPTH_CONSISTENCY) : (method != R360_PHOTO_CONSISTENCY && have_depth);
            if (!on) continue;
            const float* J = pass == 0 ? Jp : Jd;
            const float r = pass == 0 ? rp : rd;
            if (pass == 0) ++nPhoto; else ++nDepth;
            for (int a = 0; a < 6; ++a) {
                for (int b = 0; b < 6; ++b) { const float pr = J[a] * J[b]; Hf[a + 6 * b] += pr; Hd[a + 6 * b] += pr; }
                const float pr = J[a] * r; gf[a] += pr; gd[a] += pr;
            }
        }
    }
    int q = 0;
    for (int a = 0; a < 6; ++a)
        for (int b = a; b < 6; ++b, ++q) out->Hd[q] = accum_mode == 0 ? (double)Hf[a + 6 * b] : Hd[a + 6 * b];
    for (int a = 0; a < 36; ++a) out->H[a] = accum_mode == 0 ? Hf[a] : (float)Hd[a];
    for (int a = 0; a < 6; ++a) { out->gd[a] = accum_mode == 0 ? (double)gf[a] : gd[a]; out->g[a] = accum_mode == 0 ? gf[a] : (float)gd[a]; }
    out->n_visible = nVisible; out->n_photo = nPhoto; out->n_depth = nDepth;
}

template <class M>
int align_pinhole(const Frame* src, const Frame* trg, const float* guess, const r360_params* P, const float* cam,
                  int accum_mode, r360_result* out, r360_iter_record* trace, int trace_cap) {
    memset(out, 0, sizeof(*out));
    const int L = P->n_levels;
    const int per_level = 2 * P->max_iters + 2;
    float pose_estim[16], pose_tmp[16];
    memcpy(pose_estim, guess, sizeof(pose_estim));
    HessOut ho;
    memset(&ho, 0, sizeof(ho));
    bool have_hess = false, ill_posed = false;
    std::vector<float> lut;
    double error = 0.0;
    OccErr oe_acc; memset(&oe_acc, 0, sizeof(oe_acc));
    auto norm6 = [](const float* u) {
        float a = u[0] * u[0] + (u[1] * u[1] + u[2] * u[2]);
        float b = u[3] * u[3] + (u[4] * u[4] + u[5] * u[5]);
        return sqrtf(a + b);
    };
    auto exp_mul = [&](const float* upd, float* dst) {           // CPose3D::exp(update) * pose_estim, RPI.h:4375
        double ud[6], Td[16], A, B;
        for (int a = 0; a < 6; ++a) ud[a] = (double)upd[a];
        const double th2 = ud[3] * ud[3] + ud[4] * ud[4] + ud[5] * ud[5];
        if (r360_rodrigues_small(th2, &A, &B)) {
            const double th = sqrt(th2);
            double sn, cs;
            M::sincos_d(th, &sn, &cs);
            const double inv_th = 1.0 / th;
            A = sn * inv_th;
            B = (1 - cs) * (inv_th * inv_th);
        }
        r360_pseudo_exp_AB(ud, A, B, Td);
        r360_exp_translation(ud, A, B, th2, Td);
        float Tf[16];
        for (int a = 0; a < 16; ++a) Tf[a] = (float)Td[a];
        r360_mat4_mul(Tf, pose_estim, dst);
    };
    for (int level = L - 1; level >= 0 && !ill_posed; --level) {
        const PinK k = pinhole_consts(src, level, cam);
        build_lut_pinhole(src, level, P, k, lut);
        double lambda = 0.01;                                      // RPI.h:4304
        const double step = 10;
        int it = 0, ev = 0;
        float upd[6] = { 1, 1, 1, 1, 1, 1 };
        auto record = [&](const float* pose, const OccErr& oe, int accepted, int itv) {
            if (!trace) return;
            const int idx = level * per_level + ev;
            if (ev < per_level && idx < trace_cap) {
                r360_iter_record* r = &trace[idx];
                memset(r, 0, sizeof(*r));
                r->err2 = oe.photo; r->err2_depth = oe.depth; r->n_valid = oe.n_photo; r->n_valid_depth = oe.n_depth;
                r->level = level; r->it = itv; r->accepted = accepted; r->used = 1; memcpy(r->pose, pose, 64);
            }
            ++ev;
        };
        OccErr oe;
        error = error_pinhole(src, trg, level, pose_estim, P, k, lut, &oe);       // RPI.h:4312
        oe_acc = oe;
        out->passes[level] = 1;
        record(pose_estim, oe, 1, 0);
        double diff_error = error;
        while (it < P->max_iters && norm6(upd) > P->tol_update && diff_error > P->tol_residual) {
            hessgrad_pinhole(src, trg, level, pose_estim, P, k, lut, accum_mode, &ho);   // RPI.h:4340
            have_hess = true;
            ++out->passes[level];
            float Hl[36];
            const float lam = (float)lambda;
            for (int q = 0; q < 36; ++q) Hl[q] = ho.H[q];
            for (int a = 0; a < 6; ++a) Hl[a + 6 * a] = ho.H[a + 6 * a] + lam * ho.H[a + 6 * a];
            if (r360_rank6(Hl) != 6) {                             // RPI.h:4360-4368
                out->status = R360_PAIR_ILL_POSED;
                ill_posed = true;
                break;
            }
            float inv[36];
            r360_inverse6(ho.H, inv);
            r360_solve_update(inv, ho.g, upd);                      // RPI.h:4371
            exp_mul(upd, pose_tmp);
            OccErr noe;
            double new_error = error_pinhole(src, trg, level, pose_tmp, P, k, lut, &noe);   // RPI.h:4378
            ++out->passes[level];
            diff_error = error - new_error;
            if (diff_error > 0) {                                   // RPI.h:4390-4396
                lambda /= step;
                memcpy(pose_estim, pose_tmp, sizeof(pose_estim));
                error = new_error; oe_acc = noe;
                it = it + 1;
                record(pose_tmp, noe, 1, it);
            } else {
                record(pose_tmp, noe, 0, it);
                unsigned LM_it = 0;
                while (LM_it < 1 && diff_error < 0) {               // RPI.h:4399-4424, LM_maxIters = 1
                    lambda = lambda * step;
                    const float lm = (float)lambda;
                    float Hd2[36];
                    for (int q = 0; q < 36; ++q) Hd2[q] = ho.H[q];
                    for (int a = 0; a < 6; ++a) Hd2[a + 6 * a] = ho.H[a + 6 * a] + lm * ho.H[a + 6 * a];
                    r360_inverse6(Hd2, inv);
                    r360_solve_update(inv, ho.g, upd);
                    exp_mul(upd, pose_tmp);
                    new_error = error_pinhole(src, trg, level, pose_tmp, P, k, lut, &noe);
                    ++out->passes[level];
                    diff_error = error - new_error;
                    if (diff_error > 0) {
                        memcpy(pose_estim, pose_tmp, sizeof(pose_estim));
                        error = new_error; oe_acc = noe;
                        it = it + 1;
                        record(pose_tmp, noe, 1, it);
                    } else {
                        LM_it = LM_it + 1;
                        record(pose_tmp, noe, 0, it);
                    }
                }
            }
        }
        if (!ill_posed) out->iters[level] = it;
    }
    memcpy(out->pose, pose_estim, sizeof(pose_estim));
    if (have_hess) {
        memcpy(out->hessian, ho.H, sizeof(ho.H));
        memcpy(out->gradient, ho.g, sizeof(ho.g));
        out->n_visible = ho.n_visible;
    }
    out->final_error = error;
    out->final_err2 = oe_acc.photo + oe_acc.depth;
    out->final_n_valid = oe_acc.n_depth;
    return 0;
}

// ------------------------------------------------------------------ f4: the 8-sensor rig (RegisterRGBD360::RegisterDensePhotoICP)
// calcPhotoICPError_robot (RPI.h:4905-5092) and calcHessianGradient_robot (RPI.h:5100-5407), the branches taken with
// bUseSalientPixels = false, and the driver RegisterRGBD360::RegisterDensePhotoICP (RegisterRGBD360.h:344-520) that sums
// them over the rig's 8 sensors.  What the restatement keeps:
//   * the ERROR function warps with ONE matrix relPoseCam = Rt^-1 * pose * Rt formed in float, float intrinsics, a
//     double 1/z and double pixel coordinates, applies no saliency test and returns the SUM of squared weighted
//     residuals (no RMS); its depth term (methods 1, 2) uses the untransformed source depth;
//   * the HESSIAN function warps with three matrix-vector products (Rt, pose, Rt^-1), DOUBLE intrinsics (the
//     back-projection is formed in double and narrowed), applies the photo saliency `continue`, and accumulates H and g
//     in float pixel by pixel in row-major order (no OpenMP: the pragma is commented out upstream);
//   * the Hessian's depth row subtracts `jacobianRt_z`, a matrix that is DECLARED BUT NEVER ASSIGNED upstream
//     (RPI.h:5226-5228, 5372-5374: the statement `jacobianT36.block(2,0,1,6);` has no effect): the depth methods are
//     undefined behaviour there, so only PHOTO_CONSISTENCY (the driver's default and what its callers pass) is defined;
//   * the driver evaluates `new_error` at pose_estim instead of pose_estim_temp (RegisterRGBD360.h:462, 488): with a
//     deterministic summation diff_error is exactly 0, no step is ever accepted, every level runs ONE loop body and the
//     function returns its initial guess with the summed Hessian at that guess.  `faithful = 0` evaluates the candidate
//     instead (the evident intent); everything else is unchanged.
// Third-party arithmetic: Matrix4f::inverse() = r360_inverse4, 4x4 / 4x1 products summed k ascending, rank / 6x6 inverse /
// exponential as the pinhole alignFrames.
inline int round_d_to_int(double v) {
    const double r = round(v);
    return (r >= -2147483648.0 && r <= 2147483647.0) ? (int)r : INT_MIN;       // x86 cvttsd2si on overflow / NaN
}
inline void mat4_vec4(const float* M, const float* v, float* o) {               // Eigen 4x4 * 4x1, k ascending
    for (int i = 0; i < 4; ++i) o[i] = ((M[i] * v[0] + M[i + 4] * v[1]) + M[i + 8] * v[2]) + M[i + 12] * v[3];
}
double error_robot(const Frame* src, const Frame* trg, int level, const float* pose, const float* Rt,
                   const r360_params* P, const float* cam, int* n_terms) {
    const PinK k = pinhole_consts(src, level, cam);                                // RPI.h:4915-4921: float, as errorPhotoICP
    float Rt_inv[16], tmp[16], rel[16];
    r360_inverse4(Rt, Rt_inv);                                                     // RPI.h:4923
    r360_mat4_mul(Rt_inv, pose, tmp);
    r360_mat4_mul(tmp, Rt, rel);                                                   // RPI.h:4924
    const double stdDevPhoto_inv = 1. / P->std_photo;                              // RPI.h:4927
    const float* Is = src->gray[level].data();
    const float* Ds = src->depth[level].data();
    const float* It = trg->gray[level].data();
    const float* Dt = trg->depth[level].data();
    const int method = P->method;
    double error2 = 0.0;
    int n = 0;
    for (int r = 0; r < k.rows; ++r)
        for (int c = 0; c < k.cols; ++c) {
            const size_t i = (size_t)r * k.cols + c;
            float p[4];
            p[2] = Ds[i];
            if (!(P->min_depth < p[2] && p[2] < P->max_depth)) continue;           // RPI.h:5018
            p[0] = (c - k.ox) * p[2] * k.inv_fx;
            p[1] = (r - k.oy) * p[2] * k.inv_fy;
            p[3] = 1;
            float tp[4];
            mat4_vec4(rel, p, tp);                                                 // RPI.h:5025
            const double inv_z = 1.0 / tp[2];
            const double tc = (tp[0] * k.fx) * inv_z + k.ox;                       // RPI.h:5030-5031
            const double tr = (tp[1] * k.fy) * inv_z + k.oy;
            const int ri = round_d_to_int(tr), ci = round_d_to_int(tc);
            if (!((ri >= 0 && ri < k.rows) && (ci >= 0 && ci < k.cols))) continue;
            const size_t j = (size_t)ri * k.cols + ci;
            if (method == R360_PHOTO_CONSISTENCY || method == R360_PHOTO_DEPTH) {
                const float photoDiff = It[j] - Is[i];
                const double weight_photo = r360_huber(photoDiff, P->std_photo) * stdDevPhoto_inv;
                const float werr = weight_photo * photoDiff;
                error2 += werr * werr;
                ++n;
            }
            if (method == R360_DEPTH_CONSISTENCY || method == R360_PHOTO_DEPTH) {
                const float depth2 = Dt[j];
                if (std::isfinite(depth2)) {
                    const float depth1 = Ds[i];                                    // RPI.h:5066: the UNtransformed source depth
                    const float depthDiff = depth2 - depth1;
                    const float sd = P->std_depth * depth1;
                    const double weight_depth = r360_huber(depthDiff, sd) / sd;
                    const float werr = weight_depth * depthDiff;
                    error2 += werr * werr;
                    ++n;
                }
            }
        }
    if (n_terms) *n_terms = n;
    return error2;
}
// PHOTO_CONSISTENCY only (see above).  H column-major 6x6 float, accumulated as upstream; Hd / gd: the same rows in double.
void hessgrad_robot(const Frame* src, const Frame* trg, int level, const float* pose, const float* Rt,
                    const r360_params* P, const float* cam, HessOut* out) {
    const int rows = src->rows >> level, cols = src->cols >> level;
    const double scaleFactor = 1.0 / pow(2, level);                                // RPI.h:5108-5114: double
    const double fx = cam[0] * scaleFactor, fy = cam[1] * scaleFactor, ox = cam[2] * scaleFactor, oy = cam[3] * scaleFactor;
    const double inv_fx = 1. / fx, inv_fy = 1. / fy;
    const double stdDevPhoto_inv = 1. / P->std_photo;
    float Rt_inv[16];
    r360_inverse4(Rt, Rt_inv);
    const float* Is = src->gray[level].data();
    const float* Ds = src->depth[level].data();
    const float* It = trg->gray[level].data();
    const float* Ix = trg->ggx[level].data();
    const float* Iy = trg->ggy[level].data();
    float Hf[36], gf[6];
    double Hd[36], gd[6];
    for (int q = 0; q < 36; ++q) { Hf[q] = 0.f; Hd[q] = 0.0; }
    for (int q = 0; q < 6; ++q) { gf[q] = 0.f; gd[q] = 0.0; }
    int nVisible = 0, nPhoto = 0;
    for (int r = 0; r < rows; ++r)
        for (int c = 0; c < cols; ++c) {
            const size_t i = (size_t)r * cols + c;
            float p[4];
            p[2] = Ds[i];
            if (!(P->min_depth < p[2] && p[2] < P->max_depth)) continue;           // RPI.h:5297
            p[0] = (c - ox) * p[2] * inv_fx;                                       // double, narrowed (RPI.h:5299-5300)
            p[1] = (r - oy) * p[2] * inv_fy;
            p[3] = 1;
            float p1[4], p2[4], tp[4];
            mat4_vec4(Rt, p, p1);                                                  // RPI.h:5304-5306
            mat4_vec4(pose, p1, p2);
            mat4_vec4(Rt_inv, p2, tp);
            const double inv_z = 1.0 / tp[2];
            const double tc = (tp[0] * fx) * inv_z + ox;
            const double tr = (tp[1] * fy) * inv_z + oy;
            const int ri = round_d_to_int(tr), ci = round_d_to_int(tc);
            if (!((ri >= 0 && ri < rows) && (ci >= 0 && ci < cols))) continue;
            ++nVisible;
            const size_t j = (size_t)ri * cols + ci;
            // jacobianT36 = Rt_inv(3x3) * [I | -skew(p2)], products summed k ascending (RPI.h:5322-5326)
            const float x = p2[0], y = p2[1], z = p2[2];
            float T36[3][6];
            for (int a = 0; a < 3; ++a) {
                const float R0 = Rt_inv[a], R1 = Rt_inv[a + 4], R2 = Rt_inv[a + 8];
                T36[a][0] = (R0 * 1.f + R1 * 0.f) + R2 * 0.f;
                T36[a][1] = (R0 * 0.f + R1 * 1.f) + R2 * 0.f;
                T36[a][2] = (R0 * 0.f + R1 * 0.f) + R2 * 1.f;
                T36[a][3] = (R0 * 0.f + R1 * (-z)) + R2 * y;
                T36[a][4] = (R0 * z + R1 * 0.f) + R2 * (-x);
                T36[a][5] = (R0 * (-y) + R1 * x) + R2 * 0.f;
            }
            const float P00 = fx * inv_z, P11 = fy * inv_z;                       // RPI.h:5328-5337 (double, narrowed)
            const float P02 = -fx * tp[0] * inv_z * inv_z, P12 = -fy * tp[1] * inv_z * inv_z;
            float W0[6], W1[6];                                                   // jacobianWarpRt = jacobianProj23 * jacobianT36
            for (int q = 0; q < 6; ++q) {
                W0[q] = (P00 * T36[0][q] + 0.f * T36[1][q]) + P02 * T36[2][q];
                W1[q] = (0.f * T36[0][q] + P11 * T36[1][q]) + P12 * T36[2][q];
            }
            if (fabsf(Ix[j]) < P->thres_sal_int && fabsf(Iy[j]) < P->thres_sal_int) continue;     // RPI.h:5354-5355
            const float photoDiff = It[j] - Is[i];
            const double weight_photo = r360_huber(photoDiff, P->std_photo) * stdDevPhoto_inv;
            const double weightedErrorPhoto = weight_photo * photoDiff;
            const float wf = (float)weight_photo;                                 // double * Matrix<float>: the scalar is narrowed first
            const float a0 = wf * Ix[j], a1 = wf * Iy[j];
            float J[6];
            for (int q = 0; q < 6; ++q) J[q] = a0 * W0[q] + a1 * W1[q];
            const float rf = (float)weightedErrorPhoto;
            ++nPhoto;
            for (int a = 0; a < 6; ++a) {
                for (int b = 0; b < 6; ++b) { const float pr = J[a] * J[b]; Hf[a + 6 * b] += pr; Hd[a + 6 * b] += pr; }
                const float pr = J[a] * rf; gf[a] += pr; gd[a] += pr;
            }
        }
    int q = 0;
    for (int a = 0; a < 6; ++a)
        for (int b = a; b < 6; ++b, ++q) out->Hd[q] = Hd[a + 6 * b];
    for (int a = 0; a < 36; ++a) out->H[a] = Hf[a];
    for (int a = 0; a < 6; ++a) { out->gd[a] = gd[a]; out->g[a] = gf[a]; }
    out->n_visible = nVisible; out->n_photo = nPhoto; out->n_depth = 0;
}
// RegisterRGBD360::RegisterDensePhotoICP.  src / trg: the 8 sensor frames of frame2 / frame1; Rt: 8 column-major 4x4
// (calib->Rt_).  out->pose = rigidTransf, out->hessian = informationM (zero when no loop body ever ran: upstream
// returns an uninitialised matrix then), out->status = ILL-POSED when the function returns false.
// accum_mode 0: H, g of each sensor in float, summed over the sensors in float (upstream); 1: double sums.
template <class M>
int align_rig(Frame* const* src, Frame* const* trg, const float* Rt, const float* guess, const r360_params* P,
              const float* cam, int faithful, int accum_mode, r360_result* out) {
    memset(out, 0, sizeof(*out));
    const int L = P->n_levels;
    float pose_estim[16], pose_tmp[16];
    memcpy(pose_estim, guess, sizeof(pose_estim));
    float H[36], g[6];
    memset(H, 0, sizeof(H)); memset(g, 0, sizeof(g));
    double error = 0.0;
    auto norm6 = [](const float* u) {
        float a = u[0] * u[0] + (u[1] * u[1] + u[2] * u[2]);
        float b = u[3] * u[3] + (u[4] * u[4] + u[5] * u[5]);
        return sqrtf(a + b);
    };
    auto exp_mul = [&](const float* upd, float* dst) {                             // CPose3D::exp(update) * pose_estim
        double ud[6], Td[16], A, B;
        for (int a = 0; a < 6; ++a) ud[a] = (double)upd[a];
        const double th2 = ud[3] * ud[3] + ud[4] * ud[4] + ud[5] * ud[5];
        if (r360_rodrigues_small(th2, &A, &B)) {
            const double th = sqrt(th2);
            double sn, cs;
            M::sincos_d(th, &sn, &cs);
            const double inv_th = 1.0 / th;
            A = sn * inv_th;
            B = (1 - cs) * (inv_th * inv_th);
        }
        r360_pseudo_exp_AB(ud, A, B, Td);
        r360_exp_translation(ud, A, B, th2, Td);
        float Tf[16];
        for (int a = 0; a < 16; ++a) Tf[a] = (float)Td[a];
        r360_mat4_mul(Tf, pose_estim, dst);
    };
    auto total_error = [&](int level, const float* pose) {
        double e = 0.0;
        for (int s = 0; s < 8; ++s) e += error_robot(src[s], trg[s], level, pose, Rt + 16 * s, P, cam, nullptr);
        return e;
    };
    auto damped_update = [&](double lambda, float* upd) {                          // -(H + lambda diag H)^-1 g
        float Hl[36], inv[36];
        const float lam = (float)lambda;
        for (int q = 0; q < 36; ++q) Hl[q] = H[q];
        for (int a = 0; a < 6; ++a) Hl[a + 6 * a] = H[a + 6 * a] + lam * H[a + 6 * a];
        r360_inverse6(Hl, inv);
        r360_solve_update(inv, g, upd);
    };
    for (int level = L - 1; level >= 0; --level) {
        double lambda = 0.001;                                                     // RegisterRGBD360.h:391
        const double step = 10;
        int it = 0;
        const double tol_residual = pow(10, -1), tol_update = pow(10, -6);
        float upd[6] = { 1, 1, 1, 1, 1, 1 };
        error = total_error(level, pose_estim);                                    // :403-410
        out->passes[level] = 1;
        double diff_error = error;
        while (it < 10 && norm6(upd) > tol_update && diff_error > tol_residual) {
            memset(H, 0, sizeof(H)); memset(g, 0, sizeof(g));
            double Hd[21], gd[6];
            memset(Hd, 0, sizeof(Hd)); memset(gd, 0, sizeof(gd));
            int nvis = 0;
            for (int s = 0; s < 8; ++s) {                                          // :427-440
                HessOut ho;
                hessgrad_robot(src[s], trg[s], level, pose_estim, Rt + 16 * s, P, cam, &ho);
                for (int q = 0; q < 36; ++q) H[q] += ho.H[q];
                for (int q = 0; q < 6; ++q) g[q] += ho.g[q];
                for (int q = 0; q < 21; ++q) Hd[q] += ho.Hd[q];
                for (int q = 0; q < 6; ++q) gd[q] += ho.gd[q];
                nvis += ho.n_visible;
            }
            if (accum_mode != 0) {
                int q = 0;
                for (int a = 0; a < 6; ++a)
                    for (int b = a; b < 6; ++b, ++q) H[a + 6 * b] = H[b + 6 * a] = (float)Hd[q];
                for (int a = 0; a < 6; ++a) g[a] = (float)gd[a];
            }
            out->n_visible = nvis;
            ++out->passes[level];
            float Hl[36];
            const float lam = (float)lambda;
            for (int q = 0; q < 36; ++q) Hl[q] = H[q];
            for (int a = 0; a < 6; ++a) Hl[a + 6 * a] = H[a + 6 * a] + lam * H[a + 6 * a];
            if (r360_rank6(Hl) != 6) {                                             // :443-450
                out->status = R360_PAIR_ILL_POSED;
                memcpy(out->pose, pose_estim, sizeof(pose_estim));
                memcpy(out->hessian, H, sizeof(H)); memcpy(out->gradient, g, sizeof(g));
                out->final_error = error;
                return 0;
            }
            damped_update(lambda, upd);                                            // :453
            exp_mul(upd, pose_tmp);
            double new_error = total_error(level, faithful ? pose_estim : pose_tmp);    // :459-462
            ++out->passes[level];
            diff_error = error - new_error;
            if (diff_error > 0) {
                lambda /= step;
                memcpy(pose_estim, pose_tmp, sizeof(pose_estim));
                error = new_error;
                it = it + 1;
            } else {
                unsigned LM_it = 0;
                while (LM_it < 1 && diff_error < 0) {                              // :474-500
                    lambda = lambda * step;
                    damped_update(lambda, upd);
                    exp_mul(upd, pose_tmp);
                    new_error = total_error(level, faithful ? pose_estim : pose_tmp);
                    ++out->passes[level];
                    diff_error = error - new_error;
                    if (diff_error > 0) {
                        memcpy(pose_estim, pose_tmp, sizeof(pose_estim));
                        error = new_error;
                        it = it + 1;
                    }
                    LM_it = LM_it + 1;
                }
            }
        }
        out->iters[level] = it;
    }
    memcpy(out->pose, pose_estim, sizeof(pose_estim));
    memcpy(out->hessian, H, sizeof(H));
    memcpy(out->gradient, g, sizeof(g));
    out->final_error = error;
    out->final_err2 = error;
    return 0;
}

// ------------------------------------------------------------------ a10 alignFrames360
template <class M>
int align360(const Frame* src, const Frame* trg, const float* guess, const r360_params* P,
             int accum_mode, r360_result* out, r360_iter_record* trace, int trace_cap) {
    memset(out, 0, sizeof(*out));
    const int L = P->n_levels;
    const int per_level = P->max_iters + 2;
    float pose_estim[16], pose_tmp[16];
    memcpy(pose_estim, guess, sizeof(pose_estim));
    double error = 0.0, err2 = 0.0;
    int nvalid = 0;
    HessOut ho;
    memset(&ho, 0, sizeof(ho));
    bool have_hess = false;
    std::vector<float> lut;
    auto rec_at = [&](int level, int idx) -> r360_iter_record* {
        if (!trace) return nullptr;
        int k = level * per_level + idx;
        if (idx >= per_level || k >= trace_cap) return nullptr;
        return &trace[k];
    };
    bool ill_posed = false;
    for (int level = L - 1; level >= 0 && !ill_posed; --level) {
        build_lut<M>(src, level, P, lut);                                  // RPI.h:4553-4587
        double lambda = 1e0;                                               // RPI.h:4589
        const double step = 5;
        int it = 0;
        float upd[6] = { 1, 1, 1, 1, 1, 1 };                               // RPI.h:4596
        const int occ = P->occlusion;
        // errorPhotoICP_sphere / _sphereOcc1 / _sphereOcc2 (RPI.h:4598-4603, 4704-4709)
        auto eval_error = [&](const float* pose, double* e2, int* nv, OccErr* oe) -> double {
            if (occ == 0) {
                error_sphere<M>(src, trg, level, pose, P, lut, e2, nv);
                oe->photo = *e2; oe->depth = 0.0; oe->n_photo = *nv; oe->n_depth = 0;
                return sqrt(*e2 / *nv);
            }
            error_sphere_occ<M>(src, trg, level, pose, P, lut, occ, oe);
            *e2 = oe->photo + oe->depth;
            *nv = occ == 1 ? oe->n_photo + oe->n_depth : oe->n_depth;
            return occ_error_value(*oe, occ);
        };
        auto fill_err = [&](r360_iter_record* r, const OccErr& oe, double e2, int nv) {
            if (occ == 0) { r->err2 = e2; r->n_valid = nv; return; }
            r->err2 = oe.photo; r->err2_depth = oe.depth;
            r->n_valid = occ == 1 ? oe.n_photo : oe.n_depth; r->n_valid_depth = oe.n_depth;
        };
        OccErr oe;
        error = eval_error(pose_estim, &err2, &nvalid, &oe);               // RPI.h:4599
        out->passes[level] = 1;
        int ev = 0;
        r360_iter_record* cur = rec_at(level, ev++);
        if (cur) {
            memset(cur, 0, sizeof(*cur));
            fill_err(cur, oe, err2, nvalid); cur->level = level; cur->it = 0;
            cur->accepted = 1; cur->used = 1; memcpy(cur->pose, pose_estim, 64);
        }
        double diff_error = error;                                         // RPI.h:4605
        auto norm6 = [](const float* u) {
            // Eigen redux of 6 floats: (u0^2 + u1^2 + u2^2) + (u3^2 + u4^2 + u5^2), halves split 1+2
            float a = u[0] * u[0] + (u[1] * u[1] + u[2] * u[2]);
            float b = u[3] * u[3] + (u[4] * u[4] + u[5] * u[5]);
            return sqrtf(a + b);
        };
        while (it < P->max_iters && norm6(upd) > P->tol_update && diff_error > P->tol_residual) {
            if (occ == 0) hessgrad_sphere<M>(src, trg, level, pose_estim, P, lut, accum_mode, &ho);   // RPI.h:4623
            else hessgrad_sphere_occ<M>(src, trg, level, pose_estim, P, lut, accum_mode, occ, &ho);    // RPI.h:4625-4627
            have_hess = true;
            ++out->passes[level];
            if (cur) {
                cur->used |= 2; cur->n_visible = ho.n_visible;
                int q = 0;
                for (int a = 0; a < 6; ++a) for (int b = a; b < 6; ++b, ++q) cur->hessian[q] = ho.H[a + 6 * b];
                for (int a = 0; a < 6; ++a) cur->gradient[a] = ho.g[a];
            }
            float Hl[36];
            const float lam = (float)lambda;
            for (int q = 0; q < 36; ++q) Hl[q] = ho.H[q];
            for (int a = 0; a < 6; ++a) Hl[a + 6 * a] = ho.H[a + 6 * a] + lam * ho.H[a + 6 * a];
            if (r360_rank6(Hl) != 6) {                                     // RPI.h:4682-4690
                out->status = R360_PAIR_ILL_POSED;
                ill_posed = true;
                break;
            }
            {
                float inv[36];
                r360_inverse6(ho.H, inv);
                r360_solve_update(inv, ho.g, upd);                          // RPI.h:4693
                double ud[6], Td[16], A, B;
                for (int a = 0; a < 6; ++a) ud[a] = (double)upd[a];
                const double th2 = ud[3] * ud[3] + ud[4] * ud[4] + ud[5] * ud[5];
                if (r360_rodrigues_small(th2, &A, &B)) {
                    const double th = sqrt(th2);
                    double s, c;
                    M::sincos_d(th, &s, &c);
                    const double inv_th = 1.0 / th;
                    A = s * inv_th;
                    B = (1 - c) * (inv_th * inv_th);
                }
                r360_pseudo_exp_AB(ud, A, B, Td);
                float Tf[16];
                for (int a = 0; a < 16; ++a) Tf[a] = (float)Td[a];
                r360_mat4_mul(Tf, pose_estim, pose_tmp);                    // RPI.h:4697
            }
            double new_err2; int new_nvalid;
            OccErr noe;
            double new_error = eval_error(pose_tmp, &new_err2, &new_nvalid, &noe);   // RPI.h:4705
            ++out->passes[level];
            diff_error = error - new_error;                                // RPI.h:4711
            r360_iter_record* nr = rec_at(level, ev++);
            if (nr) {
                memset(nr, 0, sizeof(*nr));
                fill_err(nr, noe, new_err2, new_nvalid); nr->level = level; nr->it = it;
                nr->used = 1; memcpy(nr->pose, pose_tmp, 64);
            }
            if (diff_error > P->tol_residual) {                            // RPI.h:4715-4722
                lambda /= step;
                memcpy(pose_estim, pose_tmp, sizeof(pose_estim));
                error = new_error; err2 = new_err2; nvalid = new_nvalid;
                it = it + 1;
                if (nr) { nr->accepted = 1; nr->it = it; }
                cur = nr;
            }
        }
        if (!ill_posed) out->iters[level] = it;                            // RPI.h:4772
    }
    memcpy(out->pose, pose_estim, sizeof(pose_estim));                     // RPI.h:4783 / 4687
    if (have_hess) {
        memcpy(out->hessian, ho.H, sizeof(ho.H));
        memcpy(out->gradient, ho.g, sizeof(ho.g));
        out->n_visible = ho.n_visible;
        // SSO of the last calcHessGrad_sphere call (its level's imgSize), RPI.h:3226
    }
    out->final_error = error;
    out->final_err2 = err2;
    out->final_n_valid = nvalid;
    return 0;
}

// Warp index maps + validity masks at `pose` (what r360_dump_warp returns):
// r_idx/c_idx = transformed_r_int / transformed_c_int for every source pixel with a valid LUT
// point (INT_MIN elsewhere); valid_photo / valid_depth = validPixelsPhoto / validPixelsDepth.
template <class M>
void warp_dump(const Frame* src, const Frame* trg, int level, const float* T,
                      const r360_params* P, int32_t* ri, int32_t* ci, uint8_t* vp, uint8_t* vd) {
    std::vector<float> lut;
    build_lut<M>(src, level, P, lut);
    const LevelK k = level_consts(src, level);
    const int N = k.rows * k.cols;
    const float inv = (float)(1. / P->std_photo);
#pragma omp parallel for
    for (int i = 0; i < N; ++i) {
        int r = INT_MIN, c = INT_MIN, v = 0;
        if (lut[3 * (size_t)i] != ORC_INVALID_POINT) {
            Warped w;
            bool inb = warp_point<M>(T, &lut[3 * (size_t)i], k, w);
            r = w.r; c = w.c;
            if (inb) {
                const size_t j = (size_t)w.r * k.cols + w.c;
                float Jp[6], Jd[6], rp, rd;
                v = jacobian_rows(w, k, P, src->gray[level][i], trg->gray[level][j], trg->depth[level][j],
                                  trg->ggx[level][j], trg->ggy[level][j], trg->dgx[level][j],
                                  trg->dgy[level][j], inv, Jp, &rp, Jd, &rd);
            }
        }
        if (ri) ri[i] = r;
        if (ci) ci[i] = c;
        if (vp) vp[i] = (uint8_t)(v & 1);
        if (vd) vd[i] = (uint8_t)((v >> 1) & 1);
    }
}

// Frame360::stitchSphericalImage (Frame360.h:386-405, 1099-1148): 8 sensor images -> one sphere
// image.  sensor_rgb: 8 x size_h x size_w x 3, sensor_depth: 8 x size_h x size_w (mm, z-depth);
// Rt_inv: 8 column-major 4x4.  Outputs rows x cols of r360_stitch_geom.
template <class M>
void stitch_sphere(const R360StitchGeom& g, const float* Rt_inv, const uint8_t* sensor_rgb,
                          const uint16_t* sensor_depth, uint8_t* rgb, uint16_t* depth) {
    memset(rgb, 0, (size_t)g.rows * g.cols * 3);
    memset(depth, 0, (size_t)g.rows * g.cols * 2);
    const size_t spx = (size_t)g.size_h * g.size_w;
#pragma omp parallel for
    for (int r = 0; r < g.rows; ++r) {
        const float phi = (g.offset_phi - r) * g.angle_pixel;
        const float sp = M::sin_(phi), cp = M::cos_(phi);
        for (int c = 0; c < g.cols; ++c) {
            const int s = 7 - c / g.size_h;
            const float theta = (c + g.offset_theta) * g.angle_pixel;
            int ui, vi;
            float sc;
            if (!r360_stitch_pixel(g, Rt_inv + 16 * s, sp, cp, M::sin_(theta), M::cos_(theta), &ui, &vi, &sc)) continue;
            const size_t j = (size_t)s * spx + (size_t)vi * g.size_w + ui, i = (size_t)r * g.cols + c;
            rgb[3 * i] = sensor_rgb[3 * j]; rgb[3 * i + 1] = sensor_rgb[3 * j + 1]; rgb[3 * i + 2] = sensor_rgb[3 * j + 2];
            depth[i] = r360_stitch_range(sensor_depth[j], sc);
        }
    }
}

}  // namespace

// ====================================================================== C interface (ctypes)
extern "C" {

void orc_set_math(int mode) { g_math_mode = mode ? 1 : 0; }
int orc_get_math(void) { return g_math_mode; }

void* orc_frame_build(const uint8_t* rgb, const uint16_t* depth_mm, const float* depth_m, int rows,
                      int cols, const r360_params* P, int with_grad) {
    return build_frame(rgb, depth_mm, depth_m, rows, cols, P, with_grad);
}
void orc_frame_free(void* f) { delete (Frame*)f; }

int orc_frame_level(void* fv, int level, float* gray, float* depth, float* ggx, float* ggy,
                    float* dgx, float* dgy) {
    Frame* f = (Frame*)fv;
    if (level < 0 || level >= f->levels) return -1;
    const size_t n = (size_t)(f->rows >> level) * (f->cols >> level) * sizeof(float);
    if (gray) memcpy(gray, f->gray[level].data(), n);
    if (depth) memcpy(depth, f->depth[level].data(), n);
    if (f->has_grad) {
        if (ggx) memcpy(ggx, f->ggx[level].data(), n);
        if (ggy) memcpy(ggy, f->ggy[level].data(), n);
        if (dgx) memcpy(dgx, f->dgx[level].data(), n);
        if (dgy) memcpy(dgy, f->dgy[level].data(), n);
    }
    return 0;
}

int orc_lut(void* srcv, int level, const r360_params* P, float* xyz) {
    std::vector<float> lut;
    if (g_math_mode) build_lut<MathLibm>((Frame*)srcv, level, P, lut);
    else build_lut<MathPinned>((Frame*)srcv, level, P, lut);
    memcpy(xyz, lut.data(), lut.size() * sizeof(float));
    return 0;
}

int orc_error(void* srcv, void* trgv, int level, const float* pose, const r360_params* P,
              double* err2, int* nvalid) {
    std::vector<float> lut;
    if (g_math_mode) {
        build_lut<MathLibm>((Frame*)srcv, level, P, lut);
        error_sphere<MathLibm>((Frame*)srcv, (Frame*)trgv, level, pose, P, lut, err2, nvalid);
    } else {
        build_lut<MathPinned>((Frame*)srcv, level, P, lut);
        error_sphere<MathPinned>((Frame*)srcv, (Frame*)trgv, level, pose, P, lut, err2, nvalid);
    }
    return 0;
}

// errorPhotoICP_sphereOcc1 / Occ2 (P->occlusion = 1 / 2): out4 = {PhotoResidual, DepthResidual},
// counts = {nValidPhotoPts, nValidDepthPts}; returns through *error the function's return value.
int orc_error_occ(void* srcv, void* trgv, int level, const float* pose, const r360_params* P,
                  double* res2, int* counts, double* error) {
    std::vector<float> lut;
    OccErr oe;
    if (g_math_mode) {
        build_lut<MathLibm>((Frame*)srcv, level, P, lut);
        error_sphere_occ<MathLibm>((Frame*)srcv, (Frame*)trgv, level, pose, P, lut, P->occlusion, &oe);
    } else {
        build_lut<MathPinned>((Frame*)srcv, level, P, lut);
        error_sphere_occ<MathPinned>((Frame*)srcv, (Frame*)trgv, level, pose, P, lut, P->occlusion, &oe);
    }
    if (res2) { res2[0] = oe.photo; res2[1] = oe.depth; }
    if (counts) { counts[0] = oe.n_photo; counts[1] = oe.n_depth; }
    if (error) *error = occ_error_value(oe, P->occlusion);
    return 0;
}

// H: 36 floats column-major, g: 6 floats, Hd/gd: optional double sums (21 upper-tri + 6),
// counts: optional {n_visible, n_photo_rows, n_depth_rows}.
int orc_hessgrad(void* srcv, void* trgv, int level, const float* pose, const r360_params* P,
                 int accum_mode, float* H, float* g, double* Hd, double* gd, int* counts) {
    std::vector<float> lut;
    HessOut ho;
    const int occ = P->occlusion;
    if (g_math_mode) {
        build_lut<MathLibm>((Frame*)srcv, level, P, lut);
        if (occ) hessgrad_sphere_occ<MathLibm>((Frame*)srcv, (Frame*)trgv, level, pose, P, lut, accum_mode, occ, &ho);
        else hessgrad_sphere<MathLibm>((Frame*)srcv, (Frame*)trgv, level, pose, P, lut, accum_mode, &ho);
    } else {
        build_lut<MathPinned>((Frame*)srcv, level, P, lut);
        if (occ) hessgrad_sphere_occ<MathPinned>((Frame*)srcv, (Frame*)trgv, level, pose, P, lut, accum_mode, occ, &ho);
        else hessgrad_sphere<MathPinned>((Frame*)srcv, (Frame*)trgv, level, pose, P, lut, accum_mode, &ho);
    }
    if (H) memcpy(H, ho.H, sizeof(ho.H));
    if (g) memcpy(g, ho.g, sizeof(ho.g));
    if (Hd) memcpy(Hd, ho.Hd, sizeof(ho.Hd));
    if (gd) memcpy(gd, ho.gd, sizeof(ho.gd));
    if (counts) { counts[0] = ho.n_visible; counts[1] = ho.n_photo; counts[2] = ho.n_depth; }
    return 0;
}

int orc_warp(void* srcv, void* trgv, int level, const float* pose, const r360_params* P, int32_t* ri,
             int32_t* ci, uint8_t* vp, uint8_t* vd) {
    if (g_math_mode) warp_dump<MathLibm>((Frame*)srcv, (Frame*)trgv, level, pose, P, ri, ci, vp, vd);
    else warp_dump<MathPinned>((Frame*)srcv, (Frame*)trgv, level, pose, P, ri, ci, vp, vd);
    return 0;
}

int orc_align(void* srcv, void* trgv, const float* guess, const r360_params* P, int accum_mode,
              r360_result* out, r360_iter_record* trace, int trace_cap) {
    float ident[16] = { 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1 };
    const float* g0 = guess ? guess : ident;
    Frame* s = (Frame*)srcv;
    int rc;
    if (g_math_mode) rc = align360<MathLibm>(s, (Frame*)trgv, g0, P, accum_mode, out, trace, trace_cap);
    else rc = align360<MathPinned>(s, (Frame*)trgv, g0, P, accum_mode, out, trace, trace_cap);
    // SSO = numVisiblePixels / imgSize of the level of the last calcHessGrad_sphere call: the
    // finest level whose loop body ran.
    int lvl = -1;
    for (int l = 0; l < P->n_levels; ++l)
        if (out->passes[l] > 1) { lvl = l; break; }
    if (lvl >= 0) out->sso = (float)out->n_visible / (float)((s->rows >> lvl) * (s->cols >> lvl));
    return rc;
}

// ---- pinhole path: cam = {fx, fy, ox, oy} of setCameraMatrix (level 0)
int orc_error_pinhole(void* srcv, void* trgv, int level, const float* pose, const r360_params* P, const float* cam,
                      double* res2, int* counts, double* error) {
    const PinK k = pinhole_consts((Frame*)srcv, level, cam);
    std::vector<float> lut;
    build_lut_pinhole((Frame*)srcv, level, P, k, lut);
    OccErr oe;
    const double e = error_pinhole((Frame*)srcv, (Frame*)trgv, level, pose, P, k, lut, &oe);
    if (res2) { res2[0] = oe.photo; res2[1] = oe.depth; }
    if (counts) { counts[0] = oe.n_photo; counts[1] = oe.n_depth; }
    if (error) *error = e;
    return 0;
}
int orc_hessgrad_pinhole(void* srcv, void* trgv, int level, const float* pose, const r360_params* P, const float* cam,
                         int accum_mode, float* H, float* g, double* Hd, double* gd, int* counts) {
    const PinK k = pinhole_consts((Frame*)srcv, level, cam);
    std::vector<float> lut;
    build_lut_pinhole((Frame*)srcv, level, P, k, lut);
    HessOut ho;
    hessgrad_pinhole((Frame*)srcv, (Frame*)trgv, level, pose, P, k, lut, accum_mode, &ho);
    if (H) memcpy(H, ho.H, sizeof(ho.H));
    if (g) memcpy(g, ho.g, sizeof(ho.g));
    if (Hd) memcpy(Hd, ho.Hd, sizeof(ho.Hd));
    if (gd) memcpy(gd, ho.gd, sizeof(ho.gd));
    if (counts) { counts[0] = ho.n_visible; counts[1] = ho.n_photo; counts[2] = ho.n_depth; }
    return 0;
}
int orc_align_pinhole(void* srcv, void* trgv, const float* guess, const r360_params* P, const float* cam, int accum_mode,
                      r360_result* out, r360_iter_record* trace, int trace_cap) {
    float ident[16] = { 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1 };
    const float* g0 = guess ? guess : ident;
    if (g_math_mode) return align_pinhole<MathLibm>((Frame*)srcv, (Frame*)trgv, g0, P, cam, accum_mode, out, trace, trace_cap);
    return align_pinhole<MathPinned>((Frame*)srcv, (Frame*)trgv, g0, P, cam, accum_mode, out, trace, trace_cap);
}

// ---- f4: the 8-sensor rig
int orc_error_robot(void* srcv, void* trgv, int level, const float* pose, const float* Rt, const r360_params* P, const float* cam,
                    double* error2, int* n_terms) {
    const double e = error_robot((Frame*)srcv, (Frame*)trgv, level, pose, Rt, P, cam, n_terms);
    if (error2) *error2 = e;
    return 0;
}
int orc_hessgrad_robot(void* srcv, void* trgv, int level, const float* pose, const float* Rt, const r360_params* P, const float* cam,
                       float* H, float* g, double* Hd, double* gd, int* counts) {
    HessOut ho;
    hessgrad_robot((Frame*)srcv, (Frame*)trgv, level, pose, Rt, P, cam, &ho);
    if (H) memcpy(H, ho.H, sizeof(ho.H));
    if (g) memcpy(g, ho.g, sizeof(ho.g));
    if (Hd) memcpy(Hd, ho.Hd, sizeof(ho.Hd));
    if (gd) memcpy(gd, ho.gd, sizeof(ho.gd));
    if (counts) { counts[0] = ho.n_visible; counts[1] = ho.n_photo; counts[2] = 0; }
    return 0;
}
int orc_align_rig(void* const* srcv, void* const* trgv, const float* Rt, const float* guess, const r360_params* P, const float* cam,
                  int faithful, int accum_mode, r360_result* out) {
    float ident[16] = { 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1 };
    const float* g0 = guess ? guess : ident;
    if (g_math_mode) return align_rig<MathLibm>((Frame* const*)srcv, (Frame* const*)trgv, Rt, g0, P, cam, faithful, accum_mode, out);
    return align_rig<MathPinned>((Frame* const*)srcv, (Frame* const*)trgv, Rt, g0, P, cam, faithful, accum_mode, out);
}
void orc_inverse4(const float* A, float* inv) { r360_inverse4(A, inv); }

// Synthetic frame (host render of rgbd360_b200/csrc/synth.h).
void orc_synth_frame(int kind, int id, int rows, int cols, uint8_t* rgb, uint16_t* depth_mm) {
    double Rd[9], td[3];
    r360_synth_pose(kind, id, Rd, td);
    float R[9], t[3];
    for (int i = 0; i < 9; ++i) R[i] = (float)Rd[i];
    for (int i = 0; i < 3; ++i) t[i] = (float)td[i];
    const float res = (float)(2 * R360_PI_D / cols);
    const float half_rows = (float)(0.5 * rows - 0.5);
    std::vector<float> st(cols), ct(cols);
    for (int c = 0; c < cols; ++c) r360_sincosf(c * res, &st[c], &ct[c]);
#pragma omp parallel for
    for (int r = 0; r < rows; ++r) {
        float sp, cp;
        r360_sincosf((half_rows - r) * res, &sp, &cp);
        for (int c = 0; c < cols; ++c) {
            uint8_t g; uint16_t d;
            r360_synth_pixel(R, t, sp, cp, st[c], ct[c], &g, &d);
            const size_t i = (size_t)r * cols + c;
            rgb[3 * i] = rgb[3 * i + 1] = rgb[3 * i + 2] = g;
            depth_mm[i] = d;
        }
    }
}
void orc_synth_gt_pose(int kind, int src_id, int trg_id, double* T) { r360_synth_relpose(kind, src_id, trg_id, T); }

// Pinhole view of the same synthetic room (test data of the pinhole path): pixel (r, c)
// looks along ((c - ox) / fx, (r - oy) / fy, 1) in the camera frame of frame `id`; depth_mm is the z-depth.
// Rt (optional, column-major 4x4): pose of the camera in the frame of synthetic frame `id` (a sensor of the rig).
void orc_synth_pinhole_frame_rt(int kind, int id, int rows, int cols, float fx, float fy, float ox, float oy,
                                const float* Rt, uint8_t* rgb, uint16_t* depth_mm) {
    double Rd[9], td[3];
    r360_synth_pose(kind, id, Rd, td);
    if (Rt) {                                                      // camera-to-world = frame pose o Rt
        double R2[9], t2[3];
        for (int i = 0; i < 3; ++i) {
            for (int j = 0; j < 3; ++j) R2[3 * i + j] = Rd[3 * i] * Rt[4 * j] + Rd[3 * i + 1] * Rt[4 * j + 1] + Rd[3 * i + 2] * Rt[4 * j + 2];
            t2[i] = Rd[3 * i] * Rt[12] + Rd[3 * i + 1] * Rt[13] + Rd[3 * i + 2] * Rt[14] + td[i];
        }
        memcpy(Rd, R2, sizeof(R2)); memcpy(td, t2, sizeof(t2));
    }
    float R[9], t[3];
    for (int i = 0; i < 9; ++i) R[i] = (float)Rd[i];
    for (int i = 0; i < 3; ++i) t[i] = (float)td[i];
#pragma omp parallel for
    for (int r = 0; r < rows; ++r)
        for (int c = 0; c < cols; ++c) {
            const double x = (c - ox) / fx, y = (r - oy) / fy;
            const double n = sqrt(x * x + y * y + 1.0);
            const double dx = x / n, dy = y / n, dz = 1.0 / n;
            // r360_synth_pixel's ray is (sphi, -cphi sth, -cphi cth)
            const double cphi = sqrt(dy * dy + dz * dz);
            uint8_t g; uint16_t range_mm;
            r360_synth_pixel(R, t, (float)dx, (float)cphi, (float)(-dy / cphi), (float)(-dz / cphi), &g, &range_mm);
            const size_t i = (size_t)r * cols + c;
            rgb[3 * i] = rgb[3 * i + 1] = rgb[3 * i + 2] = g;
            depth_mm[i] = (uint16_t)lround((double)range_mm * dz);
        }
}

void orc_synth_pinhole_frame(int kind, int id, int rows, int cols, float fx, float fy, float ox, float oy,
                             uint8_t* rgb, uint16_t* depth_mm) {
    orc_synth_pinhole_frame_rt(kind, id, rows, cols, fx, fy, ox, oy, nullptr, rgb, depth_mm);
}

void orc_stitch(int size_h, int size_w, float fx, float fy, float cx, float cy, const float* Rt_inv,
                const uint8_t* sensor_rgb, const uint16_t* sensor_depth, uint8_t* rgb, uint16_t* depth) {
    const R360StitchGeom g = r360_stitch_geom(size_h, size_w, fx, fy, cx, cy);
    if (g_math_mode) stitch_sphere<MathLibm>(g, Rt_inv, sensor_rgb, sensor_depth, rgb, depth);
    else stitch_sphere<MathPinned>(g, Rt_inv, sensor_rgb, sensor_depth, rgb, depth);
}

// math hooks for tests/test_sphere_math.py
float orc_pinned_asinf(float x) { return r360_asinf(x); }
float orc_pinned_atan2f(float y, float x) { return r360_atan2f(y, x); }
float orc_pinned_sinf(float x) { return r360_sinf(x); }
float orc_pinned_cosf(float x) { return r360_cosf(x); }
void orc_pinned_sincos(double x, double* s, double* c) { r360_sincos(x, s, c); }
void orc_pinned_vec(int fn, int n, const float* a, const float* b, float* out) {
    for (int i = 0; i < n; ++i) {
        switch (fn) {
            case 0: out[i] = r360_asinf(a[i]); break;
            case 1: out[i] = r360_atan2f(a[i], b[i]); break;
            case 2: out[i] = r360_sinf(a[i]); break;
            case 3: out[i] = r360_cosf(a[i]); break;
            case 4: out[i] = asinf(a[i]); break;
            case 5: out[i] = atan2f(a[i], b[i]); break;
            case 6: out[i] = sinf(a[i]); break;
            case 7: out[i] = cosf(a[i]); break;
            case 8: out[i] = (float)r360_round_to_int(a[i]); break;
            default: out[i] = 0.f;
        }
    }
}
int orc_rank6(const float* M) { return r360_rank6(M); }
int orc_inverse6(const float* M, float* inv) { return r360_inverse6(M, inv); }
void orc_solve_update(const float* inv, const float* g, float* upd) { r360_solve_update(inv, g, upd); }
void orc_pseudo_exp(const double* v, double* T) { r360_pseudo_exp(v, T); }
void orc_se3_exp(const double* v, double* T) { r360_se3_exp(v, T); }
void orc_mat4_mul(const float* A, const float* B, float* C) { r360_mat4_mul(A, B, C); }
int orc_omp_threads(void);
}

#include <omp.h>
extern "C" int orc_omp_threads(void) { return omp_get_max_threads(); }
// torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU baseline wants all host cores.
extern "C" void orc_set_threads(int n) { if (n > 0) omp_set_num_threads(n); }
