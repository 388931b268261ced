"""Pinned math (rgbd360_b200/csrc/sphere_math.h, gn_math.h) against glibc / numpy: these are the
functions that stand in for the reference's libm calls on BOTH the CPU oracle and the GPU."""
import ctypes as C
import numpy as np


def ulp_err(a, ref):
    u = np.spacing(np.abs(ref).astype(np.float32)).astype(np.float64)
    return np.abs(a.astype(np.float64) - ref) / np.maximum(u, 1e-45)


def test_asinf_within_2ulp(orc):
    x = np.linspace(-1, 1, 400001).astype(np.float32)
    a = orc.pinned_vec(0, x)
    assert np.nanmax(ulp_err(a, np.arcsin(x.astype(np.float64)))) <= 2.5
    assert np.isnan(orc.pinned_vec(0, np.array([1.5, np.nan], np.float32))).all()
    assert orc.pinned_vec(0, np.array([1.0], np.float32))[0] == np.float32(np.pi / 2)


def test_atan2f_within_2ulp(orc):
    rng = np.random.default_rng(1)
    y = rng.standard_normal(400000).astype(np.float32)
    x = rng.standard_normal(400000).astype(np.float32)
    a = orc.pinned_vec(1, y, x)
    assert np.max(ulp_err(a, np.arctan2(y.astype(np.float64), x.astype(np.float64)))) <= 2.0
    # axes and signed zeros as libm
    yy = np.array([0.0, -0.0, 1.0, -1.0, 0.0, 0.0], np.float32)
    xx = np.array([-1.0, -1.0, 0.0, 0.0, 1.0, 0.0], np.float32)
    got = orc.pinned_vec(1, yy, xx)
    assert np.allclose(got, np.arctan2(yy, xx), atol=1e-7)
    assert np.signbit(got[1])


def test_sincosf_and_double(orc):
    t = np.linspace(-7, 7, 400001).astype(np.float32)
    assert np.max(np.abs(orc.pinned_vec(2, t) - np.sin(t.astype(np.float64)))) < 1.5e-7
    assert np.max(np.abs(orc.pinned_vec(3, t) - np.cos(t.astype(np.float64)))) < 1.5e-7
    s, c = C.c_double(), C.c_double()
    for v in np.linspace(-30, 30, 20001):
        orc.lib().orc_pinned_sincos(float(v), C.byref(s), C.byref(c))
        assert abs(s.value - np.sin(v)) < 4e-16 and abs(c.value - np.cos(v)) < 4e-16


def test_round_to_int_matches_c_round(orc):
    v = np.array([0.49999997, 0.5, 1.5, 2.5, -0.5, -1.5, -0.49999997, 1023.5, 8388607.5, np.nan, 3e9, -3e9], np.float32)
    got = orc.pinned_vec(8, v)
    exp = np.array([0, 1, 2, 3, -1, -2, 0, 1024, 8388608, -2147483648, -2147483648, -2147483648], np.float32)
    assert np.array_equal(got, exp)


def test_rank_and_inverse(orc):
    rng = np.random.default_rng(2)
    for _ in range(50):
        A = rng.standard_normal((40, 6)).astype(np.float32)
        H = (A.T @ A).astype(np.float32)
        Hc = np.ascontiguousarray(H.T)                       # column-major (symmetric anyway)
        assert orc.lib().orc_rank6(Hc.ctypes.data_as(C.c_void_p)) == np.linalg.matrix_rank(H.astype(np.float64)) == 6
        inv = np.zeros(36, np.float32)
        assert orc.lib().orc_inverse6(Hc.ctypes.data_as(C.c_void_p), inv.ctypes.data_as(C.c_void_p)) == 1
        assert np.allclose(inv.reshape(6, 6).T, np.linalg.inv(H.astype(np.float64)), rtol=2e-3, atol=1e-5)
    # rank deficient: two equal columns, and the zero matrix
    A = rng.standard_normal((40, 6)).astype(np.float32); A[:, 3] = A[:, 1]
    H = np.ascontiguousarray((A.T @ A).astype(np.float32))
    assert orc.lib().orc_rank6(H.ctypes.data_as(C.c_void_p)) == 5
    Z = np.zeros(36, np.float32)
    assert orc.lib().orc_rank6(Z.ctypes.data_as(C.c_void_p)) == 0


def test_pseudo_exp_is_mrpt_form(orc):
    """t copied verbatim, R = Rodrigues(w): first-order consistent with J_T = [I | -skew(p)]."""
    from scipy.spatial.transform import Rotation
    rng = np.random.default_rng(3)
    for scale in (1e-5, 5e-4, 1e-2, 0.3, 2.0):
        v = (rng.standard_normal(6) * scale).astype(np.float64)
        T = np.zeros(16)
        orc.lib().orc_pseudo_exp(v.ctypes.data_as(C.c_void_p), T.ctypes.data_as(C.c_void_p))
        T = T.reshape(4, 4).T
        assert np.array_equal(T[:3, 3], v[:3])
        assert np.allclose(T[:3, :3], Rotation.from_rotvec(v[3:]).as_matrix(), atol=1e-12)
        assert np.array_equal(T[3], [0, 0, 0, 1])


def test_inverse4_of_rigid_and_general_matrices(orc):
    """r360_inverse4 (the restatement of Eigen's Matrix4f::inverse() used for the sensor extrinsics, RPI.h:4923, 5125,
    Calib360.h:129 -- third-party arithmetic, shared by the reference stand-in, the oracle and the product) against numpy:
    float accuracy on the rig's own extrinsics, on random rigid poses and on general well-conditioned matrices."""
    import os
    rng = np.random.default_rng(4)
    Rt = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "frame360_raw_1.npz"))["Rt"].astype(np.float32)
    mats = [Rt[s] for s in range(8)]
    for _ in range(20):
        Q, _r = np.linalg.qr(rng.standard_normal((3, 3)))
        T = np.eye(4); T[:3, :3] = Q; T[:3, 3] = rng.standard_normal(3)
        mats.append(T.astype(np.float32))
        mats.append((np.eye(4) * 3 + rng.standard_normal((4, 4))).astype(np.float32))
    for M in mats:
        inv = orc.inverse4(M)
        want = np.linalg.inv(M.astype(np.float64))
        assert np.allclose(inv, want, rtol=2e-5, atol=2e-6 * np.abs(want).max()), np.abs(inv - want).max()
        assert np.allclose(inv.astype(np.float64) @ M.astype(np.float64), np.eye(4), atol=2e-5)
