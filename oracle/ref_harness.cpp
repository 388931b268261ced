// ref_harness.cpp -- C entry points around the UNMODIFIED reference class, compiled from where it
// lies (/root/reference/include/RegisterPhotoICP.h) into oracle/_ref/librpi_ref.so.
//
// *** TEST INFRASTRUCTURE ONLY. ***  The reference header needs Eigen, OpenCV, MRPT and PCL, none
// of which exist in this container; oracle/refshim/ provides from-scratch stand-ins for the small
// subset of those libraries the header touches (see the headers there for what is restated and how
// it is pinned).  Everything inside RegisterPhotoICP.h itself -- pyramids, gradients, the joint
// mask, the LUT, errorPhotoICP_sphere, calcHessGrad_sphere, alignFrames360 -- is the reference's own
// code, so this library pins the oracle's restatement of those ~1500 lines (tests/test_reference.py).
//
// Build (oracle/Makefile, only when /root/reference exists):
//   g++ -std=gnu++98 -fno-access-control -O2 -fopenmp -ffp-contract=off -mfma
//       -Irefshim -I/root/reference/include -shared -fPIC ref_harness.cpp -o _ref/librpi_ref.so
// gnu++98 because Miscellaneous.h:120-124 returns an ifstream as bool; -fno-access-control to read
// the private LUT_xyz_sphere / num_iterations members.  The reference prints
// "error2 <e> numValidPts <n>" from every errorPhotoICP_sphere call (RPI.h:2737); the harness
// captures std::cout at 17 significant digits and returns those lines as the per-evaluation trace.
#include <stdint.h>
#include <cstdio>
#include <cstring>
#include <numeric>
#include <iostream>
#include <iomanip>
#include <sstream>
#include <string>
#include <vector>
#include <omp.h>
using namespace std;   // the reference header relies on a leaked `using namespace std` (RPI.h:4792)

#ifdef REF_PINNED_MATH
// Second build of the same reference code: its float asin / atan2 / sin / cos calls (and the
// double sin / cos of the MRPT stand-in) are routed to the pinned operation sequences of
// rgbd360_b200/csrc/sphere_math.h -- the ones the GPU kernels execute -- so that the CUDA path can
// be compared with the reference's own code bit for bit (index maps, masks, counts).
#include "../rgbd360_b200/csrc/sphere_math.h"
inline float ref_asin(float x) { return r360_asinf(x); }
inline double ref_asin(double x) { return asin(x); }
inline float ref_atan2(float y, float x) { return r360_atan2f(y, x); }
inline double ref_atan2(double y, double x) { return atan2(y, x); }
inline float ref_sin(float x) { return r360_sinf(x); }
inline float ref_cos(float x) { return r360_cosf(x); }
inline double ref_sin(double x) { double s, c; r360_sincos(x, &s, &c); return s; }
inline double ref_cos(double x) { double s, c; r360_sincos(x, &s, &c); return c; }
#define asin ref_asin
#define atan2 ref_atan2
#define sin ref_sin
#define cos ref_cos
#endif
#include "RegisterPhotoICP.h"

namespace {
struct Capture {
    std::ostringstream ss;
    std::streambuf* old;
    Capture() { old = std::cout.rdbuf(ss.rdbuf()); std::cout << std::setprecision(17); }
    ~Capture() { std::cout.rdbuf(old); }
};
cv::Mat own_rgb(const uint8_t* rgb, int rows, int cols) {
    return cv::Mat(rows, cols, CV_8UC3, (void*)rgb).clone();
}
cv::Mat own_depth(const uint16_t* d, int rows, int cols) {
    return cv::Mat(rows, cols, CV_16UC1, (void*)d).clone();
}
Eigen::Matrix4f to_mat4(const float* p) {   // column-major in
    Eigen::Matrix4f m;
    for (int i = 0; i < 16; ++i) m.data()[i] = p[i];
    return m;
}
void copy_plane(const cv::Mat& m, float* out) {
    if (!out || m.empty()) return;
    for (int r = 0; r < m.rows; ++r) memcpy(out + (size_t)r * m.cols, m.ptr<float>(r), sizeof(float) * m.cols);
}
// parse the captured "error2 <e> numValidPts <n>" lines
int parse_trace(const std::string& log, double* err2, int* nvalid, int cap) {
    std::istringstream is(log);
    std::string line;
    int n = 0;
    while (std::getline(is, line)) {
        double e; int k;
        if (sscanf(line.c_str(), "error2 %lf numValidPts %d", &e, &k) == 2) {
            if (n < cap) { if (err2) err2[n] = e; if (nvalid) nvalid[n] = k; }
            ++n;
        }
    }
    return n;
}
}  // namespace

extern "C" {

void* ref_create(int n_levels, float min_depth, float max_depth, float std_photo, float std_depth) {
    RegisterPhotoICP* r = new RegisterPhotoICP();
    r->setNumPyr(n_levels);          // RPI.h:224 (before set*Frame, as the call-order contract says)
    r->setMinDepth(min_depth);
    r->setMaxDepth(max_depth);
    r->setGrayVariance(std_photo);   // RPI.h:242 -- sets stdDevPhoto
    r->setDepthVariance(std_depth);
    return r;
}
void ref_destroy(void* h) { delete (RegisterPhotoICP*)h; }
void ref_set_threads(int n) { if (n > 0) omp_set_num_threads(n); }
int ref_pinned_math(void) {
#ifdef REF_PINNED_MATH
    return 1;
#else
    return 0;
#endif
}

void ref_set_source(void* h, const uint8_t* rgb, const uint16_t* depth_mm, int rows, int cols) {
    cv::Mat a = own_rgb(rgb, rows, cols), d = own_depth(depth_mm, rows, cols);
    ((RegisterPhotoICP*)h)->setSourceFrame(a, d);
}
void ref_set_target(void* h, const uint8_t* rgb, const uint16_t* depth_mm, int rows, int cols) {
    cv::Mat a = own_rgb(rgb, rows, cols), d = own_depth(depth_mm, rows, cols);
    ((RegisterPhotoICP*)h)->setTargetFrame(a, d);
}

// which: 0 = source pyramids (gray, depth), 1 = target pyramids (+ 4 gradient planes)
int ref_level(void* h, int which, int level, float* gray, float* depth, float* ggx, float* ggy, float* dgx, float* dgy) {
    RegisterPhotoICP* r = (RegisterPhotoICP*)h;
    if (level < 0 || level >= r->nPyrLevels) return -1;
    if (which == 0) {
        copy_plane(r->graySrcPyr[level], gray);
        copy_plane(r->depthSrcPyr[level], depth);
    } else {
        copy_plane(r->grayTrgPyr[level], gray);
        copy_plane(r->depthTrgPyr[level], depth);
        copy_plane(r->grayTrgGradXPyr[level], ggx);
        copy_plane(r->grayTrgGradYPyr[level], ggy);
        copy_plane(r->depthTrgGradXPyr[level], dgx);
        copy_plane(r->depthTrgGradYPyr[level], dgy);
    }
    return 0;
}

// alignFrames360 (RPI.h:4519).  Outputs: pose (column-major 4x4), getHessian() (36, column-major),
// getGradient(), SSO, num_iterations[level], and the (error2, numValidPts) of every
// errorPhotoICP_sphere call in call order.  Returns the number of trace entries, or -1 - n when the
// reference reported "ILL-POSED".
int ref_align(void* h, const float* guess, int method, int occlusion, float* pose, float* H, float* g, float* sso,
              int* iters, double* trace_err2, int* trace_nvalid, int cap) {
    RegisterPhotoICP* r = (RegisterPhotoICP*)h;
    std::string log;
    {
        Capture c;
        r->alignFrames360(to_mat4(guess), (RegisterPhotoICP::costFuncType)method, occlusion);
        log = c.ss.str();
    }
    Eigen::Matrix4f P = r->getOptimalPose();
    Eigen::Matrix<float, 6, 6> Hm = r->getHessian();
    Eigen::Matrix<float, 6, 1> gm = r->getGradient();
    for (int i = 0; i < 16; ++i) pose[i] = P.data()[i];
    for (int i = 0; i < 36; ++i) H[i] = Hm.data()[i];
    for (int i = 0; i < 6; ++i) g[i] = gm.data()[i];
    *sso = r->SSO;
    for (int l = 0; l < r->nPyrLevels; ++l) iters[l] = r->num_iterations[l];
    int n = parse_trace(log, trace_err2, trace_nvalid, cap);
    return log.find("ILL-POSED") != std::string::npos ? -1 - n : n;
}

// errorPhotoICP_sphere at `level` -- only meaningful for the level whose LUT_xyz_sphere is current
// (level 0 after alignFrames360 returns normally).
double ref_error(void* h, int level, const float* pose, int method, double* err2, int* nvalid) {
    RegisterPhotoICP* r = (RegisterPhotoICP*)h;
    std::string log;
    double e;
    {
        Capture c;
        e = r->errorPhotoICP_sphere(level, to_mat4(pose), (RegisterPhotoICP::costFuncType)method);
        log = c.ss.str();
    }
    parse_trace(log, err2, nvalid, 1);
    return e;
}
// errorPhotoICP_sphereOcc1 / Occ2 (RPI.h:3232, 3720) at `level` (LUT caveat as ref_error): returns the
// function's value avPhotoResidual + avDepthResidual; av[2] = the two public members it sets.
// Order-dependent upstream (OpenMP over a shared z-buffer): call ref_set_threads(1) first.
double ref_error_occ(void* h, int level, const float* pose, int method, int occlusion, double* av) {
    RegisterPhotoICP* r = (RegisterPhotoICP*)h;
    Capture c;
    double e = occlusion == 1 ? r->errorPhotoICP_sphereOcc1(level, to_mat4(pose), (RegisterPhotoICP::costFuncType)method)
                              : r->errorPhotoICP_sphereOcc2(level, to_mat4(pose), (RegisterPhotoICP::costFuncType)method);
    if (av) { av[0] = r->avPhotoResidual; av[1] = r->avDepthResidual; }
    return e;
}
// calcHessGrad_sphere / _sphereOcc1 / _sphereOcc2 by `occlusion` (RPI.h:2745, 3373, 3861)
void ref_hessgrad_occ(void* h, int level, const float* pose, int method, int occlusion, float* H, float* g, float* sso) {
    RegisterPhotoICP* r = (RegisterPhotoICP*)h;
    Capture c;
    if (occlusion == 1) r->calcHessGrad_sphereOcc1(level, to_mat4(pose), (RegisterPhotoICP::costFuncType)method);
    else if (occlusion == 2) r->calcHessGrad_sphereOcc2(level, to_mat4(pose), (RegisterPhotoICP::costFuncType)method);
    else r->calcHessGrad_sphere(level, to_mat4(pose), (RegisterPhotoICP::costFuncType)method);
    Eigen::Matrix<float, 6, 6> Hm = r->getHessian();
    Eigen::Matrix<float, 6, 1> gm = r->getGradient();
    for (int i = 0; i < 36; ++i) H[i] = Hm.data()[i];
    for (int i = 0; i < 6; ++i) g[i] = gm.data()[i];
    *sso = r->SSO;
}
void ref_hessgrad(void* h, int level, const float* pose, int method, float* H, float* g, float* sso) {
    RegisterPhotoICP* r = (RegisterPhotoICP*)h;
    Capture c;
    r->calcHessGrad_sphere(level, to_mat4(pose), (RegisterPhotoICP::costFuncType)method);
    Eigen::Matrix<float, 6, 6> Hm = r->getHessian();
    Eigen::Matrix<float, 6, 1> gm = r->getGradient();
    for (int i = 0; i < 36; ++i) H[i] = Hm.data()[i];
    for (int i = 0; i < 6; ++i) g[i] = gm.data()[i];
    *sso = r->SSO;
}
// ---- pinhole path: setCameraMatrix (RPI.h:254), alignFrames (RPI.h:4254),
//      errorPhotoICP (RPI.h:560), calcHessGrad (RPI.h:776)
void ref_set_camera(void* h, float fx, float fy, float ox, float oy) {
    Eigen::Matrix3f K;
    K << fx, 0, ox, 0, fy, oy, 0, 0, 1;
    ((RegisterPhotoICP*)h)->setCameraMatrix(K);
}
// returns 1 when the reference reported "ILL-POSED"
int ref_align_pinhole(void* h, const float* guess, int method, float* pose, float* H, float* g, int* iters) {
    RegisterPhotoICP* r = (RegisterPhotoICP*)h;
    std::string log;
    {
        Capture c;
        r->alignFrames(to_mat4(guess), (RegisterPhotoICP::costFuncType)method, 0);
        log = c.ss.str();
    }
    Eigen::Matrix4f P = r->getOptimalPose();
    Eigen::Matrix<float, 6, 6> Hm = r->getHessian();
    Eigen::Matrix<float, 6, 1> gm = r->getGradient();
    for (int i = 0; i < 16; ++i) pose[i] = P.data()[i];
    for (int i = 0; i < 36; ++i) H[i] = Hm.data()[i];
    for (int i = 0; i < 6; ++i) g[i] = gm.data()[i];
    for (int l = 0; l < r->nPyrLevels; ++l) iters[l] = r->num_iterations[l];
    return log.find("ILL-POSED") != std::string::npos ? 1 : 0;
}
// errorPhotoICP at `level` (LUT caveat as ref_error): returns avResidual; av = {avPhotoResidual, avDepthResidual}
double ref_error_pinhole(void* h, int level, const float* pose, int method, double* av) {
    RegisterPhotoICP* r = (RegisterPhotoICP*)h;
    Capture c;
    double e = r->errorPhotoICP(level, to_mat4(pose), (RegisterPhotoICP::costFuncType)method);
    if (av) { av[0] = r->avPhotoResidual; av[1] = r->avDepthResidual; }
    return e;
}
void ref_hessgrad_pinhole(void* h, int level, const float* pose, int method, float* H, float* g) {
    RegisterPhotoICP* r = (RegisterPhotoICP*)h;
    Capture c;
    r->calcHessGrad(level, to_mat4(pose), (RegisterPhotoICP::costFuncType)method);
    Eigen::Matrix<float, 6, 6> Hm = r->getHessian();
    Eigen::Matrix<float, 6, 1> gm = r->getGradient();
    for (int i = 0; i < 36; ++i) H[i] = Hm.data()[i];
    for (int i = 0; i < 6; ++i) g[i] = gm.data()[i];
}
// current LUT_xyz_sphere (private member, RPI.h:172): n points x 3 floats
int ref_lut(void* h, float* xyz, int cap_points) {
    RegisterPhotoICP* r = (RegisterPhotoICP*)h;
    const int n = (int)r->LUT_xyz_sphere.size();
    for (int i = 0; i < n && i < cap_points; ++i)
        for (int k = 0; k < 3; ++k) xyz[3 * i + k] = r->LUT_xyz_sphere[i](k);
    return n;
}

// ---- the 8-sensor rig: calcPhotoICPError_robot (RPI.h:4905) / calcHessianGradient_robot (RPI.h:5100)
//      of ONE sensor (source / target frames and camera matrix set through the handle); Rt: column-major 4x4
double ref_error_robot(void* h, int level, const float* pose, const float* Rt, int method) {
    RegisterPhotoICP* r = (RegisterPhotoICP*)h;
    Capture c;
    const Eigen::Matrix4f M = to_mat4(Rt);
    return r->calcPhotoICPError_robot(level, to_mat4(pose), M, (RegisterPhotoICP::costFuncType)method);
}
void ref_hessgrad_robot(void* h, int level, const float* pose, const float* Rt, int method, float* H, float* g) {
    RegisterPhotoICP* r = (RegisterPhotoICP*)h;
    Capture c;
    const Eigen::Matrix4f M = to_mat4(Rt);
    r->calcHessianGradient_robot(level, to_mat4(pose), M, (RegisterPhotoICP::costFuncType)method);
    Eigen::Matrix<float, 6, 6> Hm = r->getHessian();
    Eigen::Matrix<float, 6, 1> gm = r->getGradient();
    for (int i = 0; i < 36; ++i) H[i] = Hm.data()[i];
    for (int i = 0; i < 6; ++i) g[i] = gm.data()[i];
}
}  // extern "C"

// ---- RegisterRGBD360::RegisterDensePhotoICP (RegisterRGBD360.h:344-520), VERBATIM: the Makefile cuts the member function
//      out of /root/reference/include/RegisterRGBD360.h at build time into oracle/_ref/ (a build product, never committed);
//      the scaffold below supplies what it touches -- Frame360::frameRGBD_[8] (getRGBImage / getDepthImage),
//      Frame360::calib->Rt_[8], the members rigidTransf / informationM / bRegistrationDone and the registrationType enum.
//      RegisterRGBD360.h as a whole needs Frame360.h (PCL, PbMap, boost).  NOTE: the function calls
//      omp_set_num_threads(8) and sums the sensors' errors with an OpenMP reduction whose combination order is the threads'
//      arrival order, so `error - new_error` (both evaluated at pose_estim upstream) is 0 or +-1 ulp from run to run.
#define NUM_ASUS_SENSORS 8
struct RefSensorFrame {
    cv::Mat rgb, depth;
    cv::Mat& getRGBImage() { return rgb; }
    cv::Mat& getDepthImage() { return depth; }
};
struct RefRigCalib { Eigen::Matrix4f Rt_[NUM_ASUS_SENSORS]; };
struct Frame360 {
    RefSensorFrame frameRGBD_[NUM_ASUS_SENSORS];
    RefRigCalib* calib;
};
class RegisterRGBD360 {
public:
    Eigen::Matrix4f rigidTransf;
    Eigen::Matrix<float, 6, 6> informationM;
    bool bRegistrationDone;
    enum registrationType { DEFAULT_6DoF, PLANAR_3DoF, PLANAR_ODOMETRY_3DoF };
    RegisterRGBD360() : bRegistrationDone(false) {}
#include "_ref/register_rgbd360_dense_member.inc"
};

extern "C" {
// frame1 (target) / frame2 (source): 8 x h x w x 3 u8 and 8 x h x w u16; Rt: 8 column-major 4x4; guess column-major.
// Returns the function's bool; pose = rigidTransf, info = informationM (column-major 6x6; uninitialised upstream when
// no loop body ran).
int ref_rig_align(const uint8_t* rgb1, const uint16_t* d1, const uint8_t* rgb2, const uint16_t* d2, int h, int w,
                  const float* Rt, const float* guess, int method, float* pose, float* info) {
    RefRigCalib calib;
    Frame360 f1, f2;
    f1.calib = &calib; f2.calib = &calib;
    for (int s = 0; s < NUM_ASUS_SENSORS; ++s) {
        calib.Rt_[s] = to_mat4(Rt + 16 * s);
        f1.frameRGBD_[s].rgb = own_rgb(rgb1 + (size_t)s * h * w * 3, h, w);
        f1.frameRGBD_[s].depth = own_depth(d1 + (size_t)s * h * w, h, w);
        f2.frameRGBD_[s].rgb = own_rgb(rgb2 + (size_t)s * h * w * 3, h, w);
        f2.frameRGBD_[s].depth = own_depth(d2 + (size_t)s * h * w, h, w);
    }
    RegisterRGBD360 reg;
    reg.informationM = Eigen::Matrix<float, 6, 6>::Zero();
    bool ok;
    {
        // the function prints from inside its OpenMP regions: a stateless sink (a std::stringbuf is not thread-safe)
        struct NullBuf : std::streambuf { int overflow(int ch) { return ch; } std::streamsize xsputn(const char*, std::streamsize n) { return n; } } sink;
        struct Redirect { std::streambuf* old; Redirect(std::streambuf* b) { old = std::cout.rdbuf(b); } ~Redirect() { std::cout.rdbuf(old); } } redirect(&sink);
        omp_set_dynamic(0);
        ok = reg.RegisterDensePhotoICP(&f1, &f2, to_mat4(guess), (RegisterPhotoICP::costFuncType)method, RegisterRGBD360::DEFAULT_6DoF);
    }
    for (int i = 0; i < 16; ++i) pose[i] = reg.rigidTransf.data()[i];
    for (int i = 0; i < 36; ++i) info[i] = reg.informationM.data()[i];
    return ok ? 1 : 0;
}
}
