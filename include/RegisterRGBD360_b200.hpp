// RegisterRGBD360_b200.hpp -- the dense 8-sensor registration of the reference's RegisterRGBD360 class
// (include/RegisterRGBD360.h:344-520, RegisterDensePhotoICP) on top of the C ABI (r360.h): same method name, same
// argument meaning, same result members (rigidTransf, informationM).  The PbMap registration of that class and
// Frame360 itself are outside the path (PCL, MRPT); frames are passed as their 8 sensor images.
//
//   RegisterRGBD360 reg(240, 320);                               // sensor image size (QVGA rig)
//   reg.setExtrinsics(Rt);                                       // calib->Rt_[8], column-major 4x4 each (Calib360.h:122-131)
//   bool ok = reg.RegisterDensePhotoICP(frame1, frame2, guess);  // frame1 = target, frame2 = source (RegisterRGBD360.h:344)
//   reg.getPose(); reg.getInfoMat();
//
// Upstream evaluates `new_error` at pose_estim instead of pose_estim_temp (RegisterRGBD360.h:462, 488), so no step is
// ever accepted and the function returns its guess with the rig's summed Hessian at that guess; that behaviour is the
// default here (faithful = true).  setFaithfulNewError(false) evaluates the candidate pose (the evident intent).
#ifndef REGISTER_RGBD360_B200_HPP
#define REGISTER_RGBD360_B200_HPP

#include <array>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>
#include "r360.h"

namespace r360 {

/*! The 8 sensor images of one Frame360 (frameRGBD_[s].getRGBImage() / getDepthImage()): rows x cols x 3 u8 and
 *  rows x cols u16 (z-depth, millimetres), sensor after sensor. */
struct RigFrame {
    const uint8_t* rgb;
    const uint16_t* depth_mm;
};

class RegisterRGBD360 {
  public:
    enum registrationType { DEFAULT_6DoF, PLANAR_3DoF, PLANAR_ODOMETRY_3DoF };            // RegisterRGBD360.h:76

    RegisterRGBD360(int sensor_rows, int sensor_cols, int device = 0, int n_levels = 4) : rows_(sensor_rows), cols_(sensor_cols) {
        r360_params P;
        r360_default_params_pinhole(&P);
        P.n_levels = n_levels;                                                           // RegisterPhotoICP's nPyrLevels (RPI.h:204)
        P.method = R360_PHOTO_CONSISTENCY;                                               // the default of RegisterDensePhotoICP
        if (r360_create(&ctx_, device, rows_, cols_, 16, 1, &P) != R360_OK) throw std::runtime_error(r360_last_error(nullptr));
        const float focal = 525 * ((float)cols_ / 640.0f);                              // RegisterRGBD360.h:361-369
        check(r360_set_camera(ctx_, focal, focal, (float)cols_ / 2 - 0.5f, (float)rows_ / 2 - 0.5f));
        for (int s = 0; s < 8; ++s)
            for (int k = 0; k < 16; ++k) Rt_[16 * s + k] = (k % 5 == 0) ? 1.f : 0.f;
        std::memset(info_, 0, sizeof(info_));
        for (int k = 0; k < 16; ++k) pose_[k] = (k % 5 == 0) ? 1.f : 0.f;
    }
    ~RegisterRGBD360() { r360_destroy(ctx_); }
    RegisterRGBD360(const RegisterRGBD360&) = delete;
    RegisterRGBD360& operator=(const RegisterRGBD360&) = delete;

    /*! calib->Rt_[0..7]: pose of every sensor in the robot frame, column-major 4x4 (Calib360.h:122-131). */
    void setExtrinsics(const float Rt[8 * 16]) { std::memcpy(Rt_, Rt, sizeof(Rt_)); }
    void setFaithfulNewError(bool f) { faithful_ = f; }

    /*! RegisterRGBD360.h:344: frame1 = target, frame2 = source; pose_estim column-major 4x4 (Eigen layout).
     *  Returns false where upstream does (ILL-POSED, RegisterRGBD360.h:443-450). */
    bool RegisterDensePhotoICP(const RigFrame& frame1, const RigFrame& frame2, const float* pose_estim = nullptr,
                               int method = R360_PHOTO_CONSISTENCY, registrationType = DEFAULT_6DoF) {
        if (method != R360_PHOTO_CONSISTENCY)
            throw std::invalid_argument("RegisterDensePhotoICP: only PHOTO_CONSISTENCY is defined (RPI.h:5372-5374 reads an unassigned matrix)");
        std::vector<uint8_t> roles(8, R360_ROLE_TARGET);
        check(r360_set_frames(ctx_, 0, 8, frame1.rgb, frame1.depth_mm, roles.data()));       // setTargetFrame x 8   (:380)
        roles.assign(8, R360_ROLE_SOURCE);
        check(r360_set_frames(ctx_, 8, 8, frame2.rgb, frame2.depth_mm, roles.data()));       // setSourceFrame x 8   (:379)
        const int32_t src = 8, trg = 0;
        r360_result res;
        check(r360_register_rig_pairs(ctx_, 1, &src, &trg, Rt_, pose_estim, faithful_ ? 1 : 0, &res));
        std::memcpy(pose_, res.pose, sizeof(pose_));
        std::memcpy(info_, res.hessian, sizeof(info_));
        bRegistrationDone = res.status == R360_PAIR_OK;
        return res.status == R360_PAIR_OK;
    }
    /*! rigidTransf / informationM (RegisterRGBD360.h:508-510), column-major. */
    const float* getPose() const { return pose_; }
    const float* getInfoMat() const { return info_; }
    bool bRegistrationDone = false;

  private:
    void check(int rc) const { if (rc != R360_OK) throw std::runtime_error(std::string("r360: ") + r360_last_error(ctx_)); }
    r360_ctx* ctx_ = nullptr;
    int rows_, cols_;
    float Rt_[8 * 16];
    float pose_[16], info_[36];
    bool faithful_ = true;
};

}  // namespace r360
#endif
