// RegisterPhotoICP_b200.hpp -- header-only C++ mirror of the reference class surface for the
// spherical path, on top of the C ABI (r360.h).
//
// Same method names, argument meaning and call-order contract as
// EduFdez/rgbd360 include/RegisterPhotoICP.h (cited as RPI.h:line):
//   setNumPyr before set*Frame; setTargetFrame / setSourceFrame in either order;
//   alignFrames360; then getOptimalPose / getHessian / getGradient / SSO.
// A call site such as Registration/OdometryRGBD360.cpp:189-193 compiles unchanged when OpenCV and
// Eigen are present (cv::Mat / Eigen overloads are enabled with __has_include); without them the
// same methods take r360::Image views and return std::array (column-major, Eigen's layout).
//
// One instance registers one pair at a time (two frame slots on the GPU).  For throughput, batch
// many pairs through r360_register_pairs directly.  Not thread-safe per instance (as upstream).
// Link with -lrgbd360_b200.  There is no CPU fallback: construction of the GPU context throws.
#ifndef REGISTER_PHOTO_ICP_B200_HPP
#define REGISTER_PHOTO_ICP_B200_HPP

#include <array>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>
#include "r360.h"

#if defined(__has_include)
#if __has_include(<opencv2/core/core.hpp>)
#include <opencv2/core/core.hpp>
#define R360_HAVE_OPENCV 1
#endif
#if __has_include(<Eigen/Core>)
#include <Eigen/Core>
#define R360_HAVE_EIGEN 1
#endif
#endif

namespace r360 {

// Minimal image view: rows x cols, `channels` interleaved, row stride in bytes.
struct Image {
    const void* data = nullptr;
    int rows = 0, cols = 0;
    int channels = 1;
    int elem_bytes = 1;        // 1: u8, 2: u16 (millimetres), 4: f32 (metres)
    size_t step = 0;           // bytes per row (0 = tightly packed)
    size_t row_bytes() const { return (size_t)cols * channels * elem_bytes; }
    size_t stride() const { return step ? step : row_bytes(); }
};

inline void pack(const Image& im, std::vector<uint8_t>& out) {
    out.resize(im.row_bytes() * im.rows);
    for (int r = 0; r < im.rows; ++r)
        std::memcpy(out.data() + (size_t)r * im.row_bytes(), (const uint8_t*)im.data + (size_t)r * im.stride(), im.row_bytes());
}

}  // namespace r360

class RegisterPhotoICP {
  public:
    /*! costFuncType, RPI.h:195 */
    enum costFuncType { PHOTO_CONSISTENCY = R360_PHOTO_CONSISTENCY, DEPTH_CONSISTENCY = R360_DEPTH_CONSISTENCY, PHOTO_DEPTH = R360_PHOTO_DEPTH } method;

    /*! Public result fields of the reference (RPI.h:179-192). */
    float SSO = 0.f;
    float avResidual = 0.f;          // never written by the occlusion-0 spherical path upstream either
    double avPhotoResidual = 0.0;
    double avDepthResidual = 0.0;
    int nPyrLevels;

    using Pose = std::array<float, 16>;      // column-major 4x4
    using Mat6 = std::array<float, 36>;      // column-major 6x6
    using Vec6 = std::array<float, 6>;

    explicit RegisterPhotoICP(int device = 0) : method(PHOTO_CONSISTENCY), device_(device) {   // RPI.h:201-221
        r360_default_params(&p_);
        nPyrLevels = p_.n_levels;
        res_ = r360_result();
        res_.pose[0] = res_.pose[5] = res_.pose[10] = res_.pose[15] = 1.f;
    }
    ~RegisterPhotoICP() { r360_destroy(ctx_); }
    RegisterPhotoICP(const RegisterPhotoICP&) = delete;
    RegisterPhotoICP& operator=(const RegisterPhotoICP&) = delete;

    // ---- setters, RPI.h:224-269
    void setNumPyr(int Npyr) { nPyrLevels = Npyr; p_.n_levels = Npyr; drop(); }
    void setMinDepth(float minD) { p_.min_depth = minD; drop(); }
    void setMaxDepth(float maxD) { p_.max_depth = maxD; drop(); }
    void setGrayVariance(float stdDev) { p_.std_photo = stdDev; drop(); }    // sets stdDevPhoto, RPI.h:242-245
    void setDepthVariance(float stdDev) { p_.std_depth = stdDev; drop(); }
    void setVisualization(bool) {}                                           // debug windows: not offered
    void useSaliency(bool) {}                                                // the salient-pixel branches are commented out upstream

    // ---- frames, RPI.h:480-516.  rgb: 8UC3 (channel 0 taken as R), depth: 16UC1 mm or 32FC1 m.
    void setSourceFrame(const r360::Image& rgb, const r360::Image& depth) { set_frame(0, R360_ROLE_SOURCE, rgb, depth); }
    void setTargetFrame(const r360::Image& rgb, const r360::Image& depth) { set_frame(1, R360_ROLE_TARGET, rgb, depth); }

    // ---- alignFrames360, RPI.h:4519-4784
    void alignFrames360(const Pose& pose_guess = identity(), costFuncType method_ = PHOTO_CONSISTENCY, const int occlusion = 0) {
        if (occlusion < 0 || occlusion > 2) throw std::invalid_argument("RegisterPhotoICP: occlusion must be 0, 1 or 2 (RPI.h:4517)");
        if (!ctx_ || !have_[0] || !have_[1]) throw std::logic_error("RegisterPhotoICP: setSourceFrame and setTargetFrame first");
        ensure_method(method_, occlusion);
        const int32_t s = 0, t = 1;
        check(r360_register_pairs(ctx_, 1, &s, &t, pose_guess.data(), &res_, nullptr));
        SSO = res_.sso;
        if (res_.status == R360_PAIR_ILL_POSED) avResidual = 0.f;           // RPI.h:4688
    }

    // ---- pinhole registration: setCameraMatrix (RPI.h:254), alignFrames (RPI.h:4254-4512), errorPhotoICP
    //      (RPI.h:560-775), calcHessGrad (RPI.h:776-1104).  occlusion must be 0 (the pinhole Occ variants are not built).
    void setCameraMatrix(float fx, float fy, float ox, float oy) { cam_[0] = fx; cam_[1] = fy; cam_[2] = ox; cam_[3] = oy; have_cam_ = true; if (ctx_ && p_.projection == R360_PINHOLE) check(r360_set_camera(ctx_, fx, fy, ox, oy)); }
    void alignFrames(const Pose& pose_guess = identity(), costFuncType method_ = PHOTO_CONSISTENCY, const int occlusion = 0) {
        if (occlusion != 0) throw std::invalid_argument("RegisterPhotoICP: the pinhole occlusion variants (RPI.h:1107-2023) are not built");
        if (!have_[0] || !have_[1]) throw std::logic_error("RegisterPhotoICP: setSourceFrame and setTargetFrame first");
        ensure_method(method_, 0, R360_PINHOLE);
        const int32_t s = 0, t = 1;
        check(r360_register_pairs(ctx_, 1, &s, &t, pose_guess.data(), &res_, nullptr));
    }
    double errorPhotoICP(const int& pyramidLevel, const Pose& poseGuess, costFuncType method_ = PHOTO_CONSISTENCY) {
        ensure_method(method_, 0, R360_PINHOLE);
        double pr = 0, dr = 0, e = 0; int32_t np = 0, nd = 0;
        check(r360_eval_error_pinhole(ctx_, 0, 1, pyramidLevel, poseGuess.data(), &pr, &dr, &np, &nd, &e));
        avPhotoResidual = std::sqrt(pr / (double)nd);                       // RPI.h:768
        avDepthResidual = std::sqrt(dr / (double)nd);
        avResidual = (float)(avPhotoResidual + avDepthResidual);
        return avResidual;
    }
    void calcHessGrad(const int& pyramidLevel, const Pose& poseGuess, costFuncType method_ = PHOTO_CONSISTENCY) {
        ensure_method(method_, 0, R360_PINHOLE);
        int32_t nvis = 0;
        check(r360_eval_hessgrad(ctx_, 0, 1, pyramidLevel, poseGuess.data(), res_.hessian, res_.gradient, &nvis));
    }

    /*! errorPhotoICP_sphere, RPI.h:2545-2739: sqrt(error2 / numValidPts). */
    double errorPhotoICP_sphere(const int& pyramidLevel, const Pose& poseGuess, costFuncType method_ = PHOTO_CONSISTENCY) {
        ensure_method(method_, 0);
        double e2 = 0; int32_t n = 0;
        check(r360_eval_error(ctx_, 0, 1, pyramidLevel, poseGuess.data(), &e2, &n));
        return std::sqrt(e2 / n);
    }
    /*! errorPhotoICP_sphereOcc1 / Occ2, RPI.h:3232-3369 / 3720-3858 (single-thread, source-order semantics):
        sets avPhotoResidual / avDepthResidual and returns their sum. */
    double errorPhotoICP_sphereOcc1(const int& pyramidLevel, const Pose& poseGuess, costFuncType method_ = PHOTO_CONSISTENCY) {
        return error_occ(1, pyramidLevel, poseGuess, method_);
    }
    double errorPhotoICP_sphereOcc2(const int& pyramidLevel, const Pose& poseGuess, costFuncType method_ = PHOTO_CONSISTENCY) {
        return error_occ(2, pyramidLevel, poseGuess, method_);
    }
    /*! calcHessGrad_sphereOcc1 / Occ2, RPI.h:3373-3716 / 3861-4249. */
    void calcHessGrad_sphereOcc1(const int& pyramidLevel, const Pose& poseGuess, costFuncType method_ = PHOTO_CONSISTENCY) {
        hessgrad_occ(1, pyramidLevel, poseGuess, method_);
    }
    void calcHessGrad_sphereOcc2(const int& pyramidLevel, const Pose& poseGuess, costFuncType method_ = PHOTO_CONSISTENCY) {
        hessgrad_occ(2, pyramidLevel, poseGuess, method_);
    }
    /*! calcHessGrad_sphere, RPI.h:2745-3228: fills hessian / gradient / SSO. */
    void calcHessGrad_sphere(const int& pyramidLevel, const Pose& poseGuess, costFuncType method_ = PHOTO_CONSISTENCY) {
        ensure_method(method_, 0);
        int32_t nvis = 0;
        check(r360_eval_hessgrad(ctx_, 0, 1, pyramidLevel, poseGuess.data(), res_.hessian, res_.gradient, &nvis));
        SSO = (float)nvis / (float)((rows_ >> pyramidLevel) * (cols_ >> pyramidLevel));
    }

    // ---- getters, RPI.h:273-288
    Pose getOptimalPoseArray() const { Pose p; std::memcpy(p.data(), res_.pose, sizeof(res_.pose)); return p; }
    Mat6 getHessianArray() const { Mat6 h; std::memcpy(h.data(), res_.hessian, sizeof(res_.hessian)); return h; }
    Vec6 getGradientArray() const { Vec6 g; std::memcpy(g.data(), res_.gradient, sizeof(res_.gradient)); return g; }
    const r360_result& result() const { return res_; }     // iterations per level, status, residual sums
    static Pose identity() { Pose p{}; p[0] = p[5] = p[10] = p[15] = 1.f; return p; }

#if defined(R360_HAVE_EIGEN)
    Eigen::Matrix4f getOptimalPose() const { return Eigen::Map<const Eigen::Matrix4f>(res_.pose); }
    Eigen::Matrix<float, 6, 6> getHessian() const { return Eigen::Map<const Eigen::Matrix<float, 6, 6>>(res_.hessian); }
    Eigen::Matrix<float, 6, 1> getGradient() const { return Eigen::Map<const Eigen::Matrix<float, 6, 1>>(res_.gradient); }
    void alignFrames360(const Eigen::Matrix4f pose_guess, costFuncType method_ = PHOTO_CONSISTENCY, const int occlusion = 0) {
        Pose p; std::memcpy(p.data(), pose_guess.data(), sizeof(float) * 16);
        alignFrames360(p, method_, occlusion);
    }
    void setCameraMatrix(Eigen::Matrix3f& K) { setCameraMatrix(K(0, 0), K(1, 1), K(0, 2), K(1, 2)); }   // RPI.h:254
    void alignFrames(const Eigen::Matrix4f pose_guess, costFuncType method_ = PHOTO_CONSISTENCY, const int occlusion = 0) {
        Pose p; std::memcpy(p.data(), pose_guess.data(), sizeof(float) * 16);
        alignFrames(p, method_, occlusion);
    }
#else
    Pose getOptimalPose() const { return getOptimalPoseArray(); }
    Mat6 getHessian() const { return getHessianArray(); }
    Vec6 getGradient() const { return getGradientArray(); }
#endif

#if defined(R360_HAVE_OPENCV)
    static r360::Image view(const cv::Mat& m) {
        r360::Image v; v.data = m.data; v.rows = m.rows; v.cols = m.cols; v.channels = m.channels();
        v.elem_bytes = (int)m.elemSize1(); v.step = m.step; return v;
    }
    void setSourceFrame(cv::Mat& imgRGB, cv::Mat& imgDepth) { setSourceFrame(view(imgRGB), view(imgDepth)); }
    void setTargetFrame(cv::Mat& imgRGB, cv::Mat& imgDepth) { setTargetFrame(view(imgRGB), view(imgDepth)); }
#endif

  private:
    double error_occ(int occ, int level, const Pose& pose, costFuncType m) {
        ensure_method(m, occ);
        double pr = 0, dr = 0, e = 0; int32_t np = 0, nd = 0;
        check(r360_eval_error_occ(ctx_, 0, 1, level, pose.data(), &pr, &dr, &np, &nd, &e));
        avPhotoResidual = std::sqrt(pr / (double)(occ == 1 ? np : nd));   // RPI.h:3360 / 3849
        avDepthResidual = std::sqrt(dr / (double)nd);                      // RPI.h:3362 / 3851
        return e;
    }
    void hessgrad_occ(int occ, int level, const Pose& pose, costFuncType m) {
        ensure_method(m, occ);
        int32_t nvis = 0;
        check(r360_eval_hessgrad(ctx_, 0, 1, level, pose.data(), res_.hessian, res_.gradient, &nvis));
        SSO = (float)nvis / (float)((rows_ >> level) * (cols_ >> level));
    }
    r360_params p_;
    r360_ctx* ctx_ = nullptr;
    r360_result res_;
    int device_, rows_ = 0, cols_ = 0;
    bool have_[2] = {false, false};
    float cam_[4] = {0.f, 0.f, 0.f, 0.f};
    bool have_cam_ = false;
    std::vector<uint8_t> rgb_[2], depth_[2];
    int depth_bytes_[2] = {0, 0};

    void drop() { r360_destroy(ctx_); ctx_ = nullptr; }
    void check(int rc) const {
        if (rc != R360_OK) throw std::runtime_error(std::string("r360: ") + r360_last_error(ctx_));
    }
    void upload(int slot, int role) {
        const uint8_t r = (uint8_t)role;
        if (depth_bytes_[slot] == 2) check(r360_set_frames(ctx_, slot, 1, rgb_[slot].data(), (const uint16_t*)depth_[slot].data(), &r));
        else check(r360_set_frames_f32(ctx_, slot, 1, rgb_[slot].data(), (const float*)depth_[slot].data(), &r));
    }
    void ensure_ctx(int rows, int cols) {
        if (ctx_ && rows == rows_ && cols == cols_) return;
        drop();
        rows_ = rows; cols_ = cols;
        p_.method = (int)method_or_default();
        if (r360_create(&ctx_, device_, rows, cols, 2, 1, &p_) != R360_OK)
            throw std::runtime_error(std::string("r360_create: ") + r360_last_error(nullptr));
        if (p_.projection == R360_PINHOLE) {
            if (!have_cam_) throw std::logic_error("RegisterPhotoICP: setCameraMatrix first (RPI.h:254)");
            check(r360_set_camera(ctx_, cam_[0], cam_[1], cam_[2], cam_[3]));
        }
        for (int s = 0; s < 2; ++s)
            if (have_[s] && rgb_[s].size() == (size_t)rows * cols * 3) upload(s, s == 0 ? R360_ROLE_SOURCE : R360_ROLE_TARGET);
    }
    int method_or_default() const { return p_.method; }
    void ensure_method(costFuncType m, int occlusion = 0, int projection = R360_SPHERE) {
        method = m;
        if (p_.method == (int)m && p_.occlusion == occlusion && p_.projection == projection && ctx_) return;
        p_.method = (int)m;
        p_.occlusion = occlusion;
        if (p_.projection != projection) {
            // the two registrations differ in the constants alignFrames / alignFrames360 hard-code
            p_.projection = projection;
            p_.tol_residual = projection == R360_PINHOLE ? 1e-4 : 1e-3;      // RPI.h:4308 / 4594
            p_.n_sensors_mask = projection == R360_PINHOLE ? 0 : 8;          // RPI.h:4537 (alignFrames360 only)
        }
        const int r = rows_, c = cols_;
        drop();
        ensure_ctx(r, c);
    }
    void set_frame(int slot, int role, const r360::Image& rgb, const r360::Image& depth) {
        if (rgb.channels != 3 || rgb.elem_bytes != 1) throw std::invalid_argument("RegisterPhotoICP: RGB image must be 8UC3");
        if (depth.channels != 1 || (depth.elem_bytes != 2 && depth.elem_bytes != 4)) throw std::invalid_argument("RegisterPhotoICP: depth must be 16UC1 (mm) or 32FC1 (m)");
        if (depth.rows != rgb.rows || depth.cols != rgb.cols) throw std::invalid_argument("RegisterPhotoICP: RGB / depth size mismatch");
        r360::pack(rgb, rgb_[slot]);
        r360::pack(depth, depth_[slot]);
        depth_bytes_[slot] = depth.elem_bytes;
        have_[slot] = true;
        if (!ctx_ || rgb.rows != rows_ || rgb.cols != cols_) ensure_ctx(rgb.rows, rgb.cols);
        else upload(slot, role);
    }
};

#endif  // REGISTER_PHOTO_ICP_B200_HPP
