// Drivers360_b200.hpp -- header-only host drivers that feed BATCHES of sphere pairs to the C ABI
// (r360.h), mirroring what the reference's callers do around alignFrames360:
//
//   Odometry360::run      Registration/OdometryRGBD360.cpp:137-139, 185-193
//       target = frame k, source = frame k+1, PHOTO_DEPTH; the dense pose is conjugated from the
//       sphere-image frame to the robot frame: rigidTransf = rotOffset^-1 * pose * rotOffset.
//   LoopClosure360::run   include/LoopClosure360.h:119-126, 291-321
//       setNumPyr(5), setGrayVariance(3/255); candidates = keyframes closer than 5 m to the new one;
//       guess = rotOffset * relativePose * rotOffset^-1; the edge added to the pose graph is
//       (rotOffset^-1 * getOptimalPose() * rotOffset, information = getHessian(), score = SSO).
//
// The reference registers one pair per call on the CPU; here all pairs of a sequence / all
// candidates go through ONE r360_register_pairs call (frames resident once, both roles).
// PbMap pre-registration and the pose-graph optimiser are not part of this path: guesses come from
// the caller (Identity for odometry, as OdometryRGBD360.cpp:191 offers).
#ifndef DRIVERS360_B200_HPP
#define DRIVERS360_B200_HPP

#include <array>
#include <cmath>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>
#include "r360.h"

namespace r360 {

using Pose = std::array<float, 16>;      // column-major 4x4 (Eigen::Matrix4f layout)
using Mat6 = std::array<float, 36>;

inline Pose identityPose() { Pose p{}; p[0] = p[5] = p[10] = p[15] = 1.f; return p; }
inline Pose mul(const Pose& A, const Pose& B) {
    Pose C{};
    for (int j = 0; j < 4; ++j)
        for (int i = 0; i < 4; ++i) {
            float acc = 0.f;
            for (int k = 0; k < 4; ++k) acc += A[i + 4 * k] * B[k + 4 * j];
            C[i + 4 * j] = acc;
        }
    return C;
}
inline Pose inverseRigid(const Pose& T) {     // [R t; 0 1]^-1 = [R^T  -R^T t; 0 1]
    Pose I = identityPose();
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) I[i + 4 * j] = T[j + 4 * i];
    for (int i = 0; i < 3; ++i) I[12 + i] = -(I[i] * T[12] + I[i + 4] * T[13] + I[i + 8] * T[14]);
    return I;
}
/*! rotOffset of OdometryRGBD360.cpp:137-139 / LoopClosure360.h:125-126: 157.5 deg about x. */
inline Pose rotOffset() {
    const float a = (float)(157.5 * 3.14159265359 / 180);
    Pose R = identityPose();
    R[5] = R[10] = std::cos(a);          // (1,1), (2,2)
    R[1 + 4 * 2] = std::sin(a);          // (1,2)
    R[2 + 4 * 1] = -R[1 + 4 * 2];        // (2,1)
    return R;
}
/*! robot-frame relative pose -> initial guess in the sphere-image frame (LoopClosure360.h:311). */
inline Pose toSphereFrame(const Pose& robot) { return mul(mul(rotOffset(), robot), inverseRigid(rotOffset())); }
/*! getOptimalPose() -> robot-frame relative pose (OdometryRGBD360.cpp:193, LoopClosure360.h:313). */
inline Pose toRobotFrame(const Pose& sphere) { return mul(mul(inverseRigid(rotOffset()), sphere), rotOffset()); }

/*! Keyframes closer than max_dist to keyframe `new_id` (LoopClosure360.h:289-293):
 *  |(pose_kf^-1 * pose_new).translation| < 5 m.  Returns (compare id, new id) pairs, compare id ascending. */
inline std::vector<std::pair<int, int>> loopCandidates(const std::vector<Pose>& kf_poses, int new_id, float max_dist = 5.f) {
    std::vector<std::pair<int, int>> out;
    for (int k = 0; k < (int)kf_poses.size(); ++k) {
        if (k == new_id) continue;
        const Pose rel = mul(inverseRigid(kf_poses[k]), kf_poses[new_id]);
        const float d = std::sqrt(rel[12] * rel[12] + rel[13] * rel[13] + rel[14] * rel[14]);
        if (d < max_dist) out.emplace_back(k, new_id);
    }
    return out;
}

/*! What the callers keep of one registration (LoopClosure360.h:313-321). */
struct Edge {
    int source = 0, target = 0;          // frame ids
    Pose relativePose{};                 // robot frame
    Mat6 informationMatrix{};            // getHessian()
    float SSO = 0.f;
    int status = 0;                      // R360_PAIR_*
    int iterations[R360_MAX_LEVELS]{};
};

/*! A batch of sphere frames resident on one GPU and the pair lists registered over them. */
class BatchRegistrar {
  public:
    BatchRegistrar(int rows, int cols, int max_frames, int max_pairs, int n_levels = 5, float std_photo = 3.f / 255,
                   int device = 0) {
        r360_default_params(&p_);
        p_.n_levels = n_levels;          // setNumPyr(5), LoopClosure360.h:121
        p_.std_photo = std_photo;        // setGrayVariance(3.f/255), LoopClosure360.h:124
        p_.method = R360_PHOTO_DEPTH;
        if (r360_create(&ctx_, device, rows, cols, max_frames, max_pairs, &p_))
            throw std::runtime_error(std::string("BatchRegistrar: ") + r360_last_error(nullptr));
    }
    ~BatchRegistrar() { r360_destroy(ctx_); }
    BatchRegistrar(const BatchRegistrar&) = delete;
    BatchRegistrar& operator=(const BatchRegistrar&) = delete;
    r360_ctx* ctx() { return ctx_; }

    /*! Stitched sphere frames (n x rows x cols [x 3]) into slots [first, first+n), both roles. */
    void setFrames(int first, int n, const uint8_t* rgb, const uint16_t* depth_mm) {
        check(r360_set_frames(ctx_, first, n, rgb, depth_mm, nullptr));
    }
    /*! Raw Frame360 sensor images -> stitched on the device -> slots (Frame360::stitchSphericalImage). */
    void setFramesFromSensors(const r360_rig& rig, int first, int n, const uint8_t* sensor_rgb, const uint16_t* sensor_depth_mm) {
        check(r360_stitch_frames(ctx_, &rig, first, n, sensor_rgb, sensor_depth_mm, nullptr, nullptr, nullptr));
    }

    /*! Sequence odometry over resident frames [first, first+n): pair k = (target first+k, source first+k+1),
     *  guess Identity; returns n-1 robot-frame edges. */
    std::vector<Edge> odometry(int first, int n) { return odometry(first, n, nullptr); }
    /*! Same with one robot-frame guess per pair (n-1 of them).  The reference seeds pair k with the previous pair's
     *  result (rigidTransf_dense, OdometryRGBD360.cpp:191) -- a serial dependency a batched call cannot have; callers
     *  with a motion prior pass it here, nullptr = Identity (MethodsRegisterRGBD360.cpp:448). */
    std::vector<Edge> odometry(int first, int n, const std::vector<Pose>* robot_guess) {
        std::vector<int32_t> s, t;
        std::vector<float> g;
        for (int k = 0; k + 1 < n; ++k) { t.push_back(first + k); s.push_back(first + k + 1); }
        if (robot_guess) {
            if (robot_guess->size() != s.size()) throw std::invalid_argument("odometry: one guess per pair (n - 1)");
            for (const Pose& P : *robot_guess) { const Pose gi = toSphereFrame(P); g.insert(g.end(), gi.begin(), gi.end()); }
        }
        return registerPairs(s, t, robot_guess ? g.data() : nullptr);
    }
    /*! Loop-closure candidates over resident keyframes: (compare id = SOURCE, new id = TARGET) as
     *  LoopClosure360.h:308-309; robot_guess[i] = relativePose of candidate i in the robot frame. */
    std::vector<Edge> loopClosures(const std::vector<std::pair<int, int>>& candidates, const std::vector<Pose>& robot_guess) {
        std::vector<int32_t> s, t;
        std::vector<float> g;
        for (size_t i = 0; i < candidates.size(); ++i) {
            s.push_back(candidates[i].first); t.push_back(candidates[i].second);
            const Pose gi = toSphereFrame(robot_guess[i]);
            g.insert(g.end(), gi.begin(), gi.end());
        }
        return registerPairs(s, t, g.data());
    }

  private:
    std::vector<Edge> registerPairs(const std::vector<int32_t>& s, const std::vector<int32_t>& t, const float* guess) {
        std::vector<r360_result> res(s.size());
        std::vector<Edge> out(s.size());
        if (s.empty()) return out;
        check(r360_register_pairs(ctx_, (int)s.size(), s.data(), t.data(), guess, res.data(), nullptr));
        for (size_t i = 0; i < s.size(); ++i) {
            Edge& e = out[i];
            e.source = s[i]; e.target = t[i];
            Pose p; for (int k = 0; k < 16; ++k) p[k] = res[i].pose[k];
            e.relativePose = toRobotFrame(p);
            for (int k = 0; k < 36; ++k) e.informationMatrix[k] = res[i].hessian[k];
            e.SSO = res[i].sso;
            e.status = res[i].status;
            for (int l = 0; l < R360_MAX_LEVELS; ++l) e.iterations[l] = res[i].iters[l];
        }
        return out;
    }
    void check(int rc) { if (rc) throw std::runtime_error(std::string("BatchRegistrar: ") + r360_last_error(ctx_)); }
    r360_params p_{};
    r360_ctx* ctx_ = nullptr;
};

}  // namespace r360
#endif
