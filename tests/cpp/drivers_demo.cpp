// include/Drivers360_b200.hpp: odometry over a resident sequence and loop-closure candidates, checked
// against single-pair calls through the class mirror (RegisterPhotoICP_b200.hpp).
#include <cstdio>
#include <cmath>
#include <vector>
#include "Drivers360_b200.hpp"
#include "RegisterPhotoICP_b200.hpp"

static double maxdiff(const r360::Pose& a, const r360::Pose& b) {
    double d = 0;
    for (int k = 0; k < 16; ++k) d = std::fmax(d, std::fabs(a[k] - b[k]));
    return d;
}

int main() {
    const int rows = 128, cols = 256, L = 3, n = 5;
    r360::BatchRegistrar reg(rows, cols, n, n * n, L);
    std::vector<uint8_t> rgb((size_t)n * rows * cols * 3);
    std::vector<uint16_t> depth((size_t)n * rows * cols);
    if (r360_synth_frames(reg.ctx(), 0, 60, n, rgb.data(), depth.data())) { std::printf("synth failed\n"); return 2; }
    reg.setFrames(0, n, rgb.data(), depth.data());
    std::vector<r360::Edge> odo = reg.odometry(0, n);
    bool ok = (int)odo.size() == n - 1;
    // the same pairs one at a time through the class mirror, conjugated by hand
    for (int k = 0; k + 1 < n && ok; ++k) {
        RegisterPhotoICP one;
        one.setNumPyr(L);
        one.setGrayVariance(3.f / 255);
        r360::Image t_rgb{rgb.data() + (size_t)k * rows * cols * 3, rows, cols, 3, 1, 0}, t_d{depth.data() + (size_t)k * rows * cols, rows, cols, 1, 2, 0};
        r360::Image s_rgb{rgb.data() + (size_t)(k + 1) * rows * cols * 3, rows, cols, 3, 1, 0}, s_d{depth.data() + (size_t)(k + 1) * rows * cols, rows, cols, 1, 2, 0};
        one.setTargetFrame(t_rgb, t_d);
        one.setSourceFrame(s_rgb, s_d);
        one.alignFrames360(RegisterPhotoICP::identity(), RegisterPhotoICP::PHOTO_DEPTH);
        const r360::Pose want = r360::toRobotFrame(one.getOptimalPoseArray());
        const double d = maxdiff(want, odo[k].relativePose);
        std::printf("odometry pair %d: batch vs single max |dPose| = %g, SSO %g\n", k, d, odo[k].SSO);
        ok = ok && d < 1e-6 && odo[k].status == R360_PAIR_OK && odo[k].source == k + 1 && odo[k].target == k;
    }
    // conjugation round trip and candidate enumeration
    r360::Pose G = r360::identityPose(); G[12] = 0.3f; G[13] = -0.1f; G[14] = 0.2f;
    ok = ok && maxdiff(r360::toRobotFrame(r360::toSphereFrame(G)), G) < 1e-6;
    std::vector<r360::Pose> kf(4, r360::identityPose());
    kf[1][12] = 3.f; kf[2][12] = 7.f; kf[3][13] = 4.9f;
    std::vector<std::pair<int, int>> cand = r360::loopCandidates(kf, 0);
    ok = ok && cand.size() == 2 && cand[0].first == 1 && cand[1].first == 3;
    // loop closures over the resident frames with ground-truth guesses in the robot frame
    std::vector<std::pair<int, int>> pairs = {{2, 0}, {4, 1}};
    std::vector<r360::Pose> guesses;
    for (auto& pr : pairs) {
        double T[16]; r360_synth_gt_pose(0, 60 + pr.first, 60 + pr.second, T);
        r360::Pose p; for (int k = 0; k < 16; ++k) p[k] = (float)T[k];
        guesses.push_back(r360::toRobotFrame(p));
    }
    std::vector<r360::Edge> lc = reg.loopClosures(pairs, guesses);
    for (size_t i = 0; i < lc.size(); ++i) {
        double T[16]; r360_synth_gt_pose(0, 60 + pairs[i].first, 60 + pairs[i].second, T);
        r360::Pose gt; for (int k = 0; k < 16; ++k) gt[k] = (float)T[k];
        const double d = maxdiff(r360::toSphereFrame(lc[i].relativePose), gt);
        std::printf("loop closure %zu: |pose - gt| = %g, H(0,0) = %g\n", i, d, lc[i].informationMatrix[0]);
        ok = ok && d < 2e-2 && lc[i].informationMatrix[0] > 0;
    }
    std::printf(ok ? "OK\n" : "FAIL\n");
    return ok ? 0 : 1;
}
