"""Host drivers around the path: rotOffset conjugation, loop-closure candidate
enumeration, batched odometry / loop-closure edges.  Reference call sites:
Registration/OdometryRGBD360.cpp:137-139,185-193; include/LoopClosure360.h:119-126,291-321."""
import os
import subprocess
import numpy as np
import pytest
from util import pose_err

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_rot_offset_and_conjugation(r360):
    from rgbd360_b200 import drivers
    R = drivers.rot_offset()
    a = 157.5 * 3.14159265359 / 180
    assert np.allclose(R[1:3, 1:3], [[np.cos(a), np.sin(a)], [-np.sin(a), np.cos(a)]], atol=1e-7)   # OdometryRGBD360.cpp:139
    assert np.allclose(R[:3, :3] @ R[:3, :3].T, np.eye(3), atol=1e-6)
    G = np.eye(4, dtype=np.float32); G[:3, 3] = [0.3, -0.1, 0.2]
    assert np.allclose(drivers.to_robot_frame(drivers.to_sphere_frame(G)), G, atol=1e-6)
    # a translation along the robot's y axis shows up rotated by 157.5 deg about x in the sphere frame
    S = drivers.to_sphere_frame(G)
    assert np.allclose(S[:3, 3], R[:3, :3] @ G[:3, 3], atol=1e-6)


def test_loop_candidates(r360):
    from rgbd360_b200 import drivers
    kf = [np.eye(4, dtype=np.float32) for _ in range(5)]
    kf[1][0, 3] = 3.0; kf[2][0, 3] = 7.0; kf[3][1, 3] = 4.9; kf[4][2, 3] = 5.0        # < 5 m is strict (LoopClosure360.h:293)
    assert drivers.loop_candidates(kf, 0) == [(1, 0), (3, 0)]
    assert drivers.loop_candidates(kf, 2) == [(1, 2)]


@pytest.mark.gpu
def test_batch_registrar_matches_single_pairs(orc, r360):
    from rgbd360_b200 import drivers
    rows, cols, L, n = 128, 256, 3, 5
    reg = drivers.BatchRegistrar(rows, cols, n, n * n, n_levels=L)
    rgb, dep = reg.ctx.synth_frames(0, 60, n)
    reg.set_frames(0, rgb, dep)
    edges = reg.odometry(0, n)
    assert [(e["source"], e["target"]) for e in edges] == [(k + 1, k) for k in range(n - 1)]
    P = orc.default_params(n_levels=L, std_photo=3.0 / 255)
    for k, e in enumerate(edges):
        o = orc.align(orc.Frame(rgb[k + 1], dep[k + 1], P, False), orc.Frame(rgb[k], dep[k], P, True), None, P)
        ang, dist = pose_err(drivers.to_sphere_frame(e["relativePose"]), orc.pose_from(o.pose))
        assert ang <= 1e-4 and dist <= 1e-4, (k, ang, dist)
        assert e["iterations"] == list(o.iters)[:L] and e["status"] == 0
        Ho = np.array(o.hessian, np.float64).reshape(6, 6)
        sc = np.sqrt(np.outer(np.diag(Ho), np.diag(Ho)))
        assert np.all(np.abs(e["informationMatrix"] - Ho) <= 1e-3 * sc)               # getHessian() -> information matrix
    # loop closures with robot-frame ground-truth guesses
    cands = [(2, 0), (4, 1)]
    guesses = [drivers.to_robot_frame(orc.synth_gt_pose(0, 60 + s, 60 + t).astype(np.float32)) for s, t in cands]
    lc = reg.loop_closures(cands, guesses)
    for (s, t), e in zip(cands, lc):
        ang, dist = pose_err(drivers.to_sphere_frame(e["relativePose"]), orc.synth_gt_pose(0, 60 + s, 60 + t))
        assert ang < 5e-3 and dist < 2e-2 and 0.5 < e["SSO"] <= 1.0
    reg.close()


@pytest.mark.gpu
def test_cpp_drivers_header(tmp_path):
    exe = tmp_path / "drivers_demo"
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "drivers_demo.cpp"), "-o", str(exe),
                           "-L", os.path.join(ROOT, "rgbd360_b200"), "-lrgbd360_b200",
                           "-Wl,-rpath," + os.path.join(ROOT, "rgbd360_b200")])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0 and "OK" in out.stdout, out.stdout + out.stderr
