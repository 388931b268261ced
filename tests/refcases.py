"""Input cases shared by tests/golden/make_reference_golden.py (which runs the compiled reference,
oracle/_ref/) and tests/test_reference.py (which checks the oracle -- and, under -m gpu, the CUDA
path -- against the recorded reference outputs)."""
import os
import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _synth(orc, kind, a, b, rows, cols):
    rgb_t, d_t = orc.synth_frame(kind, a, rows, cols)
    rgb_s, d_s = orc.synth_frame(kind, b, rows, cols)
    return rgb_s, d_s, rgb_t, d_t


def _holes(rgb_s, d_s, rgb_t, d_t, seed):
    """Invalid depth (0), out-of-range depth (> maxDepth, < minDepth) and colour (non-gray RGB)
    patches: exercises the validity masks, buildPyramidRange's valid-mean and the luma weights."""
    rng = np.random.default_rng(seed)
    d_s = d_s.copy(); d_t = d_t.copy(); rgb_s = rgb_s.copy(); rgb_t = rgb_t.copy()
    H, W = d_s.shape
    for d in (d_s, d_t):
        for _ in range(6):
            r, c = rng.integers(0, H - 9), rng.integers(0, W - 17)
            d[r:r + rng.integers(2, 9), c:c + rng.integers(2, 17)] = rng.choice([0, 150, 6500, 65535])
        d[rng.random(d.shape) < 0.01] = 0
    for rgb in (rgb_s, rgb_t):
        rgb[..., 0] = np.clip(rgb[..., 0].astype(np.int32) + rng.integers(-20, 21, rgb.shape[:2]), 0, 255)
        rgb[..., 2] = np.clip(rgb[..., 2].astype(np.int32) + rng.integers(-20, 21, rgb.shape[:2]), 0, 255)
    return rgb_s, d_s, rgb_t, d_t


def small_guess(seed):
    """A perturbed initial guess (float32 4x4)."""
    rng = np.random.default_rng(seed)
    w = rng.uniform(-0.02, 0.02, 3); t = rng.uniform(-0.05, 0.05, 3)
    th = np.linalg.norm(w); k = w / th
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    T = np.eye(4)
    T[:3, :3] = np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * K @ K
    T[:3, 3] = t
    return T.astype(np.float32)


# name -> (builder kwargs).  method: 0 photo, 1 depth, 2 photo+depth (RPI.h:195)
CASES = {
    "synth_128x256_L3_pd": dict(rows=128, cols=256, levels=3, method=2, frames=(0, 1)),
    "synth_128x256_L3_photo": dict(rows=128, cols=256, levels=3, method=0, frames=(2, 3)),
    "synth_128x256_L3_depth": dict(rows=128, cols=256, levels=3, method=1, frames=(4, 5)),
    "synth_256x512_L4_pd_guess": dict(rows=256, cols=512, levels=4, method=2, frames=(6, 7), guess=11),
    "synth_256x512_L4_pd_far": dict(rows=256, cols=512, levels=4, method=2, frames=(0, 9)),
    "synth_128x256_L3_holes": dict(rows=128, cols=256, levels=3, method=2, frames=(8, 9), holes=5),
    "synth_256x512_L5_odo": dict(rows=256, cols=512, levels=5, method=2, frames=(20, 21), std_photo=3.0 / 255),
    "loop_128x256_L3": dict(rows=128, cols=256, levels=3, method=2, kind=1, frames=(3, 17), guess_gt=True),
    "sample_pair_1920x320_L4": dict(sample=True, levels=4, method=2),
    # BASELINE.json's own sizes: configs[1] (1024x512, 3 levels) and configs[2] (2048x1024, 4 levels), batch frames 2j / 2j+1
    "synth_512x1024_L3_pd": dict(rows=512, cols=1024, levels=3, method=2, frames=(10, 11)),
    "synth_1024x2048_L4_pd": dict(rows=1024, cols=2048, levels=4, method=2, frames=(12, 13)),
    # quirk 9: ILL-POSED (RPI.h:4682-4690) -- photo-only on a texture that varies along columns only over a
    # constant-range sphere: the first twist column of J is identically 0, rank(H + lambda diag H) = 5
    "illposed_cols_only_photo": dict(special="illposed", rows=64, cols=128, levels=2, method=0),
    # no salient pixel at all: numValidPts = 0, error = sqrt(0/0) = NaN, the loop never runs
    "no_valid_pixels_const": dict(special="const", rows=64, cols=128, levels=2, method=2),
}


def make_case(orc, name):
    """-> dict(rgb_s, d_s, rgb_t, d_t, levels, method, guess (4x4 f32 or None), std_photo)."""
    c = CASES[name]
    if c.get("special"):
        rows, cols = c["rows"], c["cols"]
        d = np.full((rows, cols), 2000, np.uint16)
        if c["special"] == "illposed":
            x = (np.sin(np.arange(cols) * 2 * np.pi / 32) * 60 + 128).astype(np.uint8)
            gt = np.ascontiguousarray(np.repeat(np.repeat(x[None, :, None], rows, 0), 3, 2))
            gs = np.ascontiguousarray(np.roll(gt, 1, axis=1))
        else:
            gt = gs = np.full((rows, cols, 3), 128, np.uint8)
        return dict(rgb_s=gs, d_s=d, rgb_t=gt, d_t=d, levels=c["levels"], method=c["method"], guess=None,
                    std_photo=6.0 / 255)
    if c.get("sample"):
        z = np.load(os.path.join(GOLD, "sample_pair.npz"))
        rgb_s, d_s, rgb_t, d_t = z["src_rgb"], z["src_depth"], z["trg_rgb"], z["trg_depth"]
    else:
        kind = c.get("kind", 0)
        a, b = c["frames"]
        rgb_s, d_s, rgb_t, d_t = _synth(orc, kind, a, b, c["rows"], c["cols"])   # target a, source b
        if "holes" in c:
            rgb_s, d_s, rgb_t, d_t = _holes(rgb_s, d_s, rgb_t, d_t, c["holes"])
    guess = None
    if "guess" in c:
        guess = small_guess(c["guess"])
    if c.get("guess_gt"):
        a, b = c["frames"]
        G = orc.synth_gt_pose(c.get("kind", 0), b, a)
        guess = (small_guess(99).astype(np.float64) @ G).astype(np.float32)
    return dict(rgb_s=rgb_s, d_s=d_s, rgb_t=rgb_t, d_t=d_t, levels=c["levels"], method=c["method"],
                guess=guess, std_photo=c.get("std_photo", 6.0 / 255))


def probe_poses(n=3):
    """Level-0 probe poses for errorPhotoICP_sphere / calcHessGrad_sphere called directly."""
    return [np.eye(4, dtype=np.float32)] + [small_guess(100 + i) for i in range(n - 1)]


# ---- pinhole path: QVGA sensor of the rig (Calib360.h:75-77), views of the synthetic room
PINHOLE_CAM = (262.5, 262.5, 159.5, 119.5)          # fx, fy, ox, oy
# name -> (scene kind, target frame id, source frame id, levels, method, guess)
PINHOLE_CASES = {
    "pin_odo_L4_pd": (0, 0, 1, 4, 2, None),
    "pin_odo_L3_pd": (0, 4, 5, 3, 2, None),
    "pin_loop_L4_pd_gt": (1, 3, 17, 4, 2, "gt"),
    "pin_far_L4_pd_rejected": (0, 0, 9, 4, 2, None),          # every first step is rejected: the damped retry runs
    "pin_odo_L3_depth_illposed": (0, 2, 3, 3, 1, None),        # depth-only on a flat wall: ILL-POSED at the coarsest level
    "pin_odo_L3_photo_nan": (0, 2, 3, 3, 0, None),             # photo-only: 0/0 = NaN error, the loop never runs
    "pin_loop_L4_pd_gt2": (1, 5, 9, 4, 2, "gt"),
    "pin_odo_L3_pd_holes": (0, 6, 7, 3, 2, None, 7),           # invalid / out-of-range depth and colour patches
}


# ---- the 8-sensor rig: RegisterRGBD360::RegisterDensePhotoICP over QVGA views of the synthetic room
# name -> (scene kind, frame1 = target id, frame2 = source id, levels, guess)
RIG_CASES = {
    "rig_odo_L4": (0, 0, 1, 4, "small"),
    "rig_odo_L3_identity": (0, 4, 5, 3, None),
    "rig_loop_L4_gt": (1, 3, 17, 4, "gt"),
}


def rig_extrinsics():
    """calib->Rt_: the reference's own Calibration/Extrinsics/Rt_0N.txt (kept in the ingest fixture), 8 x 4 x 4 float32."""
    return np.load(os.path.join(GOLD, "frame360_raw_1.npz"))["Rt"].astype(np.float32)


def make_rig_case(orc, name, rows=240, cols=320):
    kind, a, b, levels, gm = RIG_CASES[name]
    Rt = rig_extrinsics()
    rgb1, d1 = orc.synth_rig_frame(kind, a, rows, cols, Rt)          # frame1: target
    rgb2, d2 = orc.synth_rig_frame(kind, b, rows, cols, Rt)          # frame2: source
    guess = None
    if gm == "small":
        guess = small_guess(7)
    elif gm == "gt":
        guess = (small_guess(99).astype(np.float64) @ orc.synth_gt_pose(kind, b, a)).astype(np.float32)
    return dict(rgb1=rgb1, d1=d1, rgb2=rgb2, d2=d2, Rt=Rt, levels=levels, guess=guess, cam=orc.rig_camera(rows, cols),
                kind=kind, frames=(a, b))


def make_pinhole_case(orc, name, rows=240, cols=320):
    kind, a, b, levels, method, gm = PINHOLE_CASES[name][:6]
    rgb_t, d_t = orc.synth_pinhole_frame(kind, a, rows, cols, *PINHOLE_CAM)
    rgb_s, d_s = orc.synth_pinhole_frame(kind, b, rows, cols, *PINHOLE_CAM)
    if len(PINHOLE_CASES[name]) > 6:
        rgb_s, d_s, rgb_t, d_t = _holes(rgb_s, d_s, rgb_t, d_t, PINHOLE_CASES[name][6])
    guess = None
    if gm == "gt":
        guess = (small_guess(99).astype(np.float64) @ orc.synth_gt_pose(kind, b, a)).astype(np.float32)
    return dict(rgb_s=rgb_s, d_s=d_s, rgb_t=rgb_t, d_t=d_t, levels=levels, method=method, guess=guess, cam=PINHOLE_CAM)
