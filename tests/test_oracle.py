"""CPU tests of the oracle (oracle/rpi_oracle.cpp): pinned against python-cv2 golden vectors for
the OpenCV pieces, against finite differences / analytic ground truth for the rest (the
reference ships no test or recorded output for this path -- SURVEY.md 8c)."""
import json
import os
import numpy as np
import pytest
from util import pose_err, small_pose

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def cv2v():
    return np.load(os.path.join(GOLD, "cv2_vectors.npz"))


def test_gray_conversion_matches_cv2(orc, cv2v):
    """cvtColor(RGB2GRAY) 8U + convertTo(CV_32F, 1/255): bit-exact vs cv2 golden vector."""
    rgb = cv2v["rgb"]
    P = orc.default_params(n_levels=1)
    f = orc.Frame(rgb, np.full(rgb.shape[:2], 1000, np.uint16), P, False)
    g = f.level(0)["gray"]
    assert np.array_equal(g, cv2v["gray"].astype(np.float32) * np.float32(1.0 / 255))
    # cv2's own convertTo result (scale 1/255 applied in float)
    assert np.array_equal((cv2v["gray_u8"].astype(np.float32) * np.float32(1.0 / 255)), cv2v["gray_f"])


def test_depth_scaling(orc):
    rgb = np.zeros((8, 16, 3), np.uint8)
    d = np.arange(128, dtype=np.uint16).reshape(8, 16) * 47
    f = orc.Frame(rgb, d, orc.default_params(n_levels=1), False)
    assert np.array_equal(f.level(0)["depth"], d.astype(np.float32) * np.float32(0.001))


@pytest.mark.parametrize("k", ["", "2"])
def test_pyrdown_matches_cv2_within_4ulp(orc, cv2v, k):
    """pyrDown f32: same taps/borders as cv2; op order is OpenCV-2.4's, cv2 4.x uses FMA -> few ulp."""
    src, ref = cv2v["pyr_in" + k], cv2v["pyr_out" + k]
    P = orc.default_params(n_levels=2)
    g8 = np.zeros(src.shape + (3,), np.uint8)
    f = orc.Frame(g8, src.astype(np.float32), P, False)      # float depth path aliases the input
    # feed the f32 plane through the gray pyramid: build a frame whose gray equals src exactly is not
    # possible from u8, so call the level builder through a gray-exact trick: 255*k/255 grid
    lv = _pyrdown_via_oracle(orc, src)
    d = np.abs(lv.view(np.int32).astype(np.int64) - ref.view(np.int32).astype(np.int64))
    assert d.max() <= 4
    assert np.allclose(lv, ref, rtol=0, atol=4e-7)


def _pyrdown_via_oracle(orc, src):
    """The oracle only exposes pyrDown through the gray pyramid; numpy restatement of the SAME
    formula (horizontal scalar, vertical (r0+r4)+(r2+r2) + 4*((r1+r3)+r2), *1/256) is checked
    against the oracle on a u8-representable image, then applied to src."""
    def pd(s):
        H, W = s.shape
        f = np.float32
        def refl(i, n):
            i = np.where(i < 0, -i, i); return np.where(i >= n, 2 * n - 2 - i, i)
        xs = np.arange(W // 2) * 2
        c = [s[:, refl(xs + k, W)] for k in (-2, -1, 0, 1, 2)]
        row = ((c[2] * f(6) + (c[1] + c[3]) * f(4)) + c[0]) + c[4]
        ys = np.arange(H // 2) * 2
        r = [row[refl(ys + k, H)] for k in (-2, -1, 0, 1, 2)]
        return (((r[0] + r[4]) + (r[2] + r[2])) + ((r[1] + r[3]) + r[2]) * f(4)) * f(1 / 256)
    rng = np.random.default_rng(5)
    g = rng.integers(0, 256, (32, 48), dtype=np.uint8)
    rgb = np.stack([g, g, g], -1)
    P = orc.default_params(n_levels=2)
    fr = orc.Frame(rgb, np.full(g.shape, 1000, np.uint16), P, False)
    assert np.array_equal(fr.level(1)["gray"], pd(fr.level(0)["gray"]))
    return pd(src.astype(np.float32))


def test_range_pyramid(orc):
    """buildPyramidRange: mean of valid 2x2 parents, else 0 (RPI.h:322-350)."""
    d = np.array([[1000, 2000, 0, 0], [3000, 7000, 0, 100], [500, 500, 5999, 6000], [500, 500, 300, 301]], np.uint16)
    rgb = np.zeros((4, 4, 3), np.uint8)
    f = orc.Frame(rgb, d, orc.default_params(n_levels=2), False)
    l1 = f.level(1)["depth"]
    f32 = np.float32
    m = lambda *v: (sum((f32(x) * f32(0.001) for x in v), f32(0))) / f32(len(v))
    assert l1[0, 0] == ((f32(1000) * f32(0.001) + f32(2000) * f32(0.001)) + f32(3000) * f32(0.001)) / f32(3)
    assert l1[0, 1] == 0.0                                     # all parents invalid (0 and 0.1 m)
    assert l1[1, 0] == m(500, 500, 500, 500)
    assert l1[1, 1] == (f32(5999) * f32(0.001) + f32(301) * f32(0.001)) / f32(2)   # 6.0 and 0.3 excluded (strict)


def test_gradient_definition_and_joint_mask(orc):
    """calcGradientXY: harmonic mean on strictly monotone triples, 0 elsewhere / on borders;
    sensor-joint columns k*W/8-1, k*W/8 zeroed."""
    rows, cols = 16, 64
    g = (np.arange(cols)[None, :] * 3 + np.arange(rows)[:, None] * 2).astype(np.uint8)
    g[5, 10] = g[5, 9]                      # plateau -> zero x-gradient at (5,10) and (5,9)
    rgb = np.stack([g, g, g], -1)
    P = orc.default_params(n_levels=1)
    f = orc.Frame(rgb, np.full((rows, cols), 1000, np.uint16), P, True)
    L = f.level(0)
    s = L["gray"]
    assert np.all(L["ggx"][0] == 0) and np.all(L["ggx"][-1] == 0) and np.all(L["ggx"][:, 0] == 0) and np.all(L["ggx"][:, -1] == 0)
    r, c = 7, 21
    exp = np.float32(2) / (np.float32(1) / (s[r, c + 1] - s[r, c]) + np.float32(1) / (s[r, c] - s[r, c - 1]))
    assert L["ggx"][r, c] == exp
    assert L["ggx"][5, 10] == 0 and L["ggx"][5, 9] == 0
    for k in range(1, 8):
        assert np.all(L["ggx"][:, k * 8 - 1] == 0) and np.all(L["ggx"][:, k * 8] == 0)
        assert np.all(L["ggy"][:, k * 8] == 0)
    assert np.all(L["dgx"] == 0) and np.all(L["dgy"] == 0)       # constant depth: never strictly monotone
    f2 = orc.Frame(rgb, np.full((rows, cols), 1000, np.uint16), orc.default_params(n_levels=1, n_sensors_mask=0), True)
    assert f2.level(0)["ggx"][7, 8] != 0


@pytest.fixture(scope="module")
def pair(orc):
    rows, cols, L = 128, 256, 3
    P = orc.default_params(n_levels=L)
    rgb_t, d_t = orc.synth_frame(0, 0, rows, cols)
    rgb_s, d_s = orc.synth_frame(0, 1, rows, cols)
    return dict(P=P, trg=orc.Frame(rgb_t, d_t, P, True), src=orc.Frame(rgb_s, d_s, P, False), L=L, rows=rows, cols=cols)


def test_lut_and_projection_roundtrip(orc, pair):
    """Identity pose: every valid source pixel re-projects onto itself (LUT convention RPI.h:4575-4582
    is the inverse of the projection RPI.h:2676-2680), except column 0 <-> theta = 0 / 2 PI."""
    for level in range(pair["L"]):
        ri, ci, vp, vd = orc.warp(pair["src"], pair["trg"], level, np.eye(4), pair["P"])
        r, c = pair["rows"] >> level, pair["cols"] >> level
        rr, cc = np.divmod(np.arange(r * c), c)
        assert np.array_equal(ri, rr)
        bad = ci != cc
        # column 0: y = +0, z < 0 -> atan2 = +pi -> theta = 2 PI -> c' = nCols (dropped, quirk 3)
        assert np.all(cc[bad] == 0) and np.all(ci[bad] == c)
        assert bad.sum() <= r


def test_c_equals_ncols_is_dropped(orc, pair):
    """Quirk 3: c' == nCols is dropped, not wrapped (RPI.h:2683)."""
    T = np.eye(4)                     # column 0: theta = atan2(+0, z<0) + PI = 2 PI -> c' = nCols
    ri, ci, vp, vd = orc.warp(pair["src"], pair["trg"], 0, T, pair["P"])
    cols = pair["cols"]
    assert (ci == cols).any()
    assert not vp[ci == cols].any() and not vd[ci == cols].any()


def test_photo_continue_suppresses_depth(orc, pair):
    """Quirk 2: a pixel skipped by the photometric saliency test contributes no depth term."""
    ri, ci, vp, vd = orc.warp(pair["src"], pair["trg"], 0, np.eye(4), pair["P"])
    assert not (vd & ~vp).any()
    assert vp.sum() > 0.3 * vp.size and vd.sum() > 0


def test_error_counts_match_masks(orc, pair):
    for T in (np.eye(4), small_pose(0.01, 0.02, -0.01, 0.02, 0.01, -0.03)):
        e2, nv = orc.error(pair["src"], pair["trg"], 0, T, pair["P"])
        ri, ci, vp, vd = orc.warp(pair["src"], pair["trg"], 0, T, pair["P"])
        assert nv == int(vp.sum()) + int(vd.sum())
        h = orc.hessgrad(pair["src"], pair["trg"], 0, T, pair["P"])
        assert h["n_photo"] == int(vp.sum()) and h["n_depth"] == int(vd.sum())
        assert h["n_visible"] == int(((ri >= 0) & (ri < pair["rows"]) & (ci < pair["cols"]) & (ri > -2**31)).sum())


def test_faithful_and_stable_accumulation_agree(orc, pair):
    T = small_pose(0.01, 0.0, -0.01, 0.02, 0.0, -0.02)
    a = orc.hessgrad(pair["src"], pair["trg"], 0, T, pair["P"], accum=orc.ACC_FAITHFUL)
    b = orc.hessgrad(pair["src"], pair["trg"], 0, T, pair["P"], accum=orc.ACC_STABLE)
    sc = np.sqrt(np.outer(np.diag(b["H"]), np.diag(b["H"])))
    assert np.all(np.abs(a["H"] - b["H"]) <= 2e-4 * sc)
    assert a["n_visible"] == b["n_visible"]


def test_gradient_is_descent_direction(orc, pair):
    """g = J^T r: stepping along -H^-1 g reduces the error (first-order consistency of the analytic
    Jacobians, the twist order [t | w] and the left-multiplied pseudo-exponential)."""
    import ctypes as C
    T0 = small_pose(0.004, -0.003, 0.005, 0.01, -0.01, 0.015)
    gt = orc.synth_gt_pose(0, 1, 0).astype(np.float32)
    for level in (2, 1):
        h = orc.hessgrad(pair["src"], pair["trg"], level, T0, pair["P"])
        e0, n0 = orc.error(pair["src"], pair["trg"], level, T0, pair["P"])
        upd = -np.linalg.solve(h["H"].astype(np.float64), h["g"].astype(np.float64))
        Tm = np.zeros(16)
        orc.lib().orc_pseudo_exp(upd.ctypes.data_as(C.c_void_p), Tm.ctypes.data_as(C.c_void_p))
        T1 = Tm.reshape(4, 4).T @ T0.astype(np.float64)
        e1, n1 = orc.error(pair["src"], pair["trg"], level, T1, pair["P"])
        assert e1 / n1 < e0 / n0
        a0, d0 = pose_err(T0, gt); a1, d1 = pose_err(T1, gt)
        assert a1 + d1 < a0 + d0                      # and moves towards the ground truth


def test_align_converges_to_ground_truth(orc, pair):
    gt = orc.synth_gt_pose(0, 1, 0)
    for mode in (orc.MATH_PINNED, orc.MATH_LIBM):
        orc.set_math(mode)
        try:
            res = orc.align(pair["src"], pair["trg"], None, pair["P"])
        finally:
            orc.set_math(orc.MATH_PINNED)
        ang, dist = pose_err(orc.pose_from(res.pose), gt)
        assert res.status == 0 and ang < 5e-3 and dist < 1.5e-2, (ang, dist)   # 256x128: 1.4 deg/px
        assert sum(res.iters) >= 2


def test_pinned_and_libm_modes_agree(orc, pair):
    """The two arithmetic modes differ only in asin/atan2/sin/cos implementations (<= 2 ulp):
    index maps agree on > 99.9 % of pixels, sums within 1e-3, final poses within 1e-3."""
    T = small_pose(0.01, -0.02, 0.015, 0.03, -0.02, 0.05)
    out = {}
    for mode in (orc.MATH_PINNED, orc.MATH_LIBM):
        orc.set_math(mode)
        try:
            out[mode] = (orc.warp(pair["src"], pair["trg"], 0, T, pair["P"]), orc.error(pair["src"], pair["trg"], 0, T, pair["P"]),
                         orc.align(pair["src"], pair["trg"], None, pair["P"]))
        finally:
            orc.set_math(orc.MATH_PINNED)
    (wa, ea, ra), (wb, eb, rb) = out[0], out[1]
    assert np.mean((wa[0] == wb[0]) & (wa[1] == wb[1])) > 0.999
    assert abs(ea[0] - eb[0]) < 1e-3 * ea[0]
    ang, dist = pose_err(orc.pose_from(ra.pose), orc.pose_from(rb.pose))
    assert ang < 1e-3 and dist < 1e-3


def test_ill_posed_returns_guess(orc):
    """Textureless, gradient-free input: H = 0 -> rank != 6 -> status ILL_POSED, pose = guess
    (RPI.h:4682-4690) ... or no valid pixel at all -> NaN error -> loop never runs."""
    rows, cols = 32, 64
    rgb = np.full((rows, cols, 3), 128, np.uint8)
    d = np.full((rows, cols), 2000, np.uint16)
    P = orc.default_params(n_levels=2)
    t = orc.Frame(rgb, d, P, True); s = orc.Frame(rgb, d, P, False)
    g = small_pose(tx=0.01)
    res = orc.align(s, t, g, P)
    assert np.allclose(orc.pose_from(res.pose), g)
    assert list(res.iters)[:2] == [0, 0]
    assert res.final_n_valid == 0


def test_sample_pair_golden(orc):
    """Config #1: the reference's own sample pair; the committed oracle output is the regression
    vector (no reference-recorded pose exists upstream)."""
    d = np.load(os.path.join(GOLD, "sample_pair.npz"))
    gold = json.load(open(os.path.join(GOLD, "sample_pair_oracle.json")))["pinned"]
    P = orc.default_params(n_levels=4)
    trg = orc.Frame(d["trg_rgb"], d["trg_depth"], P, True)
    src = orc.Frame(d["src_rgb"], d["src_depth"], P, False)
    res = orc.align(src, trg, None, P)
    assert list(res.iters)[:4] == gold["iters"]
    assert res.final_n_valid == gold["final_n_valid"]
    assert abs(res.final_err2 - gold["final_err2"]) <= 1e-9 * gold["final_err2"]
    assert np.allclose(np.array(res.pose), np.array(gold["pose"]), atol=1e-6)
