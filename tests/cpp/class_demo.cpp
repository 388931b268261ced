// Drop-in check of include/RegisterPhotoICP_b200.hpp: the call sequence of
// Registration/OdometryRGBD360.cpp:189-193 (setNumPyr, setTargetFrame, setSourceFrame,
// alignFrames360(guess, PHOTO_DEPTH), getOptimalPose) against r360_register_pairs on the same frames.
#include <cstdio>
#include <cmath>
#include <vector>
#include "RegisterPhotoICP_b200.hpp"

int main() {
    const int rows = 128, cols = 256, L = 3;
    r360_params P; r360_default_params(&P); P.n_levels = L;
    r360_ctx* ctx = nullptr;
    if (r360_create(&ctx, 0, rows, cols, 2, 1, &P)) { std::printf("create failed: %s\n", r360_last_error(nullptr)); return 2; }
    std::vector<uint8_t> rgb((size_t)2 * rows * cols * 3);
    std::vector<uint16_t> depth((size_t)2 * rows * cols);
    if (r360_synth_frames(ctx, 0, 0, 2, rgb.data(), depth.data())) { std::printf("synth failed\n"); return 2; }
    const uint8_t roles[2] = {R360_ROLE_TARGET, R360_ROLE_SOURCE};      // frame 0 target, frame 1 source
    r360_set_frames(ctx, 0, 2, rgb.data(), depth.data(), roles);
    const int32_t s = 1, t = 0;
    r360_result ref;
    if (r360_register_pairs(ctx, 1, &s, &t, nullptr, &ref, nullptr)) { std::printf("register failed: %s\n", r360_last_error(ctx)); return 2; }

    RegisterPhotoICP reg;
    reg.setNumPyr(L);
    r360::Image rgb0{rgb.data(), rows, cols, 3, 1, 0}, rgb1{rgb.data() + (size_t)rows * cols * 3, rows, cols, 3, 1, 0};
    r360::Image d0{depth.data(), rows, cols, 1, 2, 0}, d1{depth.data() + (size_t)rows * cols, rows, cols, 1, 2, 0};
    reg.setTargetFrame(rgb0, d0);
    reg.setSourceFrame(rgb1, d1);
    reg.alignFrames360(RegisterPhotoICP::identity(), RegisterPhotoICP::PHOTO_DEPTH);
    auto pose = reg.getOptimalPoseArray();
    double dmax = 0;
    for (int k = 0; k < 16; ++k) dmax = std::fmax(dmax, std::fabs(pose[k] - ref.pose[k]));
    double T[16]; r360_synth_gt_pose(0, 1, 0, T);
    double terr = 0; for (int k = 12; k < 15; ++k) terr = std::fmax(terr, std::fabs(T[k] - pose[k]));
    std::printf("class vs C ABI max |dPose| = %g, |t - t_gt| = %g, SSO = %g, iters = %d %d %d\n", dmax, terr, reg.SSO,
                reg.result().iters[0], reg.result().iters[1], reg.result().iters[2]);
    double e = reg.errorPhotoICP_sphere(0, pose, RegisterPhotoICP::PHOTO_DEPTH);
    std::printf("rms at optimum = %g (final_error %g)\n", e, reg.result().final_error);
    r360_destroy(ctx);
    // occlusion variants through the class surface (RPI.h:4519 third argument; RPI.h:3232, 3373, 3720, 3861)
    bool occ_ok = true;
    for (int occ = 1; occ <= 2; ++occ) {
        reg.alignFrames360(RegisterPhotoICP::identity(), RegisterPhotoICP::PHOTO_DEPTH, occ);
        auto po = reg.getOptimalPoseArray();
        double to = 0; for (int k = 12; k < 15; ++k) to = std::fmax(to, std::fabs(T[k] - po[k]));
        const double eo = occ == 1 ? reg.errorPhotoICP_sphereOcc1(0, po, RegisterPhotoICP::PHOTO_DEPTH)
                                   : reg.errorPhotoICP_sphereOcc2(0, po, RegisterPhotoICP::PHOTO_DEPTH);
        if (occ == 1) reg.calcHessGrad_sphereOcc1(0, po, RegisterPhotoICP::PHOTO_DEPTH);
        else reg.calcHessGrad_sphereOcc2(0, po, RegisterPhotoICP::PHOTO_DEPTH);
        std::printf("occlusion %d: |t - t_gt| = %g, error = %g (final_error %g), avPhoto %g avDepth %g, SSO %g\n", occ, to, eo,
                    reg.result().final_error, reg.avPhotoResidual, reg.avDepthResidual, reg.SSO);
        occ_ok = occ_ok && to < 2e-2 && std::fabs(eo - (reg.avPhotoResidual + reg.avDepthResidual)) < 1e-12 &&
                 reg.SSO > 0.5f && reg.SSO <= 1.0f;
    }
    // pinhole registration through the class surface (RPI.h:254, 4254, 560, 776); the sphere images serve as
    // generic RGB-D input here -- the numbers only have to be self-consistent
    reg.setCameraMatrix(200.f, 200.f, 127.5f, 63.5f);
    reg.alignFrames(RegisterPhotoICP::identity(), RegisterPhotoICP::PHOTO_DEPTH);
    auto pp = reg.getOptimalPoseArray();
    const double fe = reg.result().final_error;
    const double ep = reg.errorPhotoICP(0, pp, RegisterPhotoICP::PHOTO_DEPTH);
    reg.calcHessGrad(0, pp, RegisterPhotoICP::PHOTO_DEPTH);
    auto Hh = reg.getHessianArray();
    std::printf("pinhole: iters %d %d %d, final_error %g, errorPhotoICP at the optimum %g, H00 %g\n", reg.result().iters[0],
                reg.result().iters[1], reg.result().iters[2], fe, ep, Hh[0]);
    occ_ok = occ_ok && std::isfinite(ep) && Hh[0] > 0.f && pp[15] == 1.f;
    reg.alignFrames360(RegisterPhotoICP::identity(), RegisterPhotoICP::PHOTO_DEPTH);          // and back to the sphere
    auto back = reg.getOptimalPoseArray();
    for (int k = 0; k < 16; ++k) occ_ok = occ_ok && back[k] == ref.pose[k];
    const bool ok = occ_ok && dmax == 0.0 && terr < 2e-2 && std::fabs(e - reg.result().final_error) < 1e-6 * e;
    std::printf(ok ? "OK\n" : "FAIL\n");
    return ok ? 0 : 1;
}
