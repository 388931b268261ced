"""the pinhole registration of RegisterPhotoICP -- setCameraMatrix (RPI.h:254),
errorPhotoICP (RPI.h:560-775), calcHessGrad (RPI.h:776-1104), alignFrames (RPI.h:4254-4512), occlusion 0.

tests/golden/reference_pinhole.json holds what the compiled reference (oracle/_ref, both arithmetic
variants, one OpenMP thread: H and g are accumulated pixel by pixel in float under `omp critical`)
produces on 8 pinhole views of the synthetic room (generator: tests/golden/make_reference_golden.py pinhole).

CPU tests: oracle == recording bit for bit.  GPU tests (-m gpu): CUDA path vs the oracle -- counters
exact (index maps are bit-exact: the projection is evaluated with the reference's own operation
sequence, including its double-precision 1/z), sums 1e-4, poses 1e-4.
"""
import json
import os
import numpy as np
import pytest
import refcases

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
REL = 1e-4
POSE_TOL = 1e-4


@pytest.fixture(scope="module")
def gold_pin():
    with open(os.path.join(GOLD, "reference_pinhole.json")) as f:
        return json.load(f)["cases"]


def _eq_nan(a, b):
    return (a != a) if b is None else (a == b)


def _frames(orc, case):
    P = orc.pinhole_params(n_levels=case["levels"], method=case["method"])
    trg = orc.Frame(case["rgb_t"], case["d_t"], P, True)
    src = orc.Frame(case["rgb_s"], case["d_s"], P, False)
    return P, src, trg


@pytest.mark.parametrize("pinned", [False, True], ids=["libm", "pinned"])
@pytest.mark.parametrize("name", list(refcases.PINHOLE_CASES))
def test_oracle_pinhole_equals_recorded_reference(orc, gold_pin, name, pinned):
    case = refcases.make_pinhole_case(orc, name)
    ref = gold_pin[name]["pinned" if pinned else "libm"]
    _check_oracle_pinhole_against_reference(orc, case, ref, pinned)


def _check_oracle_pinhole_against_reference(orc, case, ref, pinned):
    """Oracle vs one reference record (recorded or live) of a pinhole run, bit for bit at one thread."""
    orc.set_math(orc.MATH_PINNED if pinned else orc.MATH_LIBM)
    orc.lib().orc_set_threads(1)
    try:
        P, src, trg = _frames(orc, case)
        L = case["levels"]
        res = orc.align_pinhole(src, trg, case["guess"], P, case["cam"], accum=orc.ACC_FAITHFUL)
        assert list(res.iters)[:L] == ref["iters"]
        assert (res.status != 0) == ref["ill_posed"]
        assert np.array_equal(orc.pose_from(res.pose).astype(np.float64).ravel(), np.array(ref["pose"]))
        if ref["H"] is not None:
            assert np.array_equal(np.array(res.hessian, np.float64).reshape(6, 6).T.ravel(), np.array(ref["H"]))
            assert np.array_equal(np.array(res.gradient, np.float64), np.array(ref["g"]))
        for pr in ref["probes_level0"]:
            T = np.array(pr["pose"]).reshape(4, 4)
            eo = orc.error_pinhole(src, trg, 0, T, P, case["cam"])
            assert _eq_nan(eo["error"], pr["error"])
            with np.errstate(all="ignore"):
                assert _eq_nan(np.sqrt(np.float64(eo["photo"]) / np.float64(eo["n_depth"])), pr["av_photo"])
                assert _eq_nan(np.sqrt(np.float64(eo["depth"]) / np.float64(eo["n_depth"])), pr["av_depth"])
            hg = orc.hessgrad_pinhole(src, trg, 0, T, P, case["cam"], accum=orc.ACC_FAITHFUL)
            assert np.array_equal(hg["H"].astype(np.float64).ravel(), np.array(pr["H"]))
            assert np.array_equal(hg["g"].astype(np.float64), np.array(pr["g"]))
    finally:
        orc.set_math(orc.MATH_PINNED)


def test_live_reference_pinhole_matches_recording(orc, gold_pin):
    from oracle import refbind
    if not refbind.available():
        pytest.skip("oracle/_ref not built and /root/reference absent")
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_reference_golden", os.path.join(GOLD, "make_reference_golden.py"))
    m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
    for name in ["pin_odo_L3_pd", "pin_odo_L3_pd_holes"]:
        case = refcases.make_pinhole_case(orc, name)
        live = m.run_reference_pinhole(case, True)
        assert json.loads(json.dumps(live)) == gold_pin[name]["pinned"], name


def _random_pinhole_case(orc, seed):
    """A seeded random pinhole problem: resolution (QVGA or half of it, intrinsics scaled), pyramid depth, cost
    function, views of either scene, optional holes, optional perturbed / ground-truth-based guess."""
    rng = np.random.default_rng(7000 + seed)
    levels = int(rng.integers(1, 5))
    half = rng.random() < 0.5 and levels <= 3
    rows, cols = (120, 160) if half else (240, 320)
    cam = tuple(v * (0.5 if half else 1.0) for v in refcases.PINHOLE_CAM)
    kind = int(rng.integers(0, 2))
    a = int(rng.integers(0, 30)); b = a + int(rng.integers(1, 3))
    rgb_t, d_t = orc.synth_pinhole_frame(kind, a, rows, cols, *cam)
    rgb_s, d_s = orc.synth_pinhole_frame(kind, b, rows, cols, *cam)
    if rng.random() < 0.5:
        rgb_s, d_s, rgb_t, d_t = refcases._holes(rgb_s, d_s, rgb_t, d_t, int(rng.integers(0, 1 << 30)))
    u = rng.random()
    guess = None
    if u < 0.3:
        guess = refcases.small_guess(int(rng.integers(0, 1 << 30)))
    elif u < 0.65:
        guess = (refcases.small_guess(int(rng.integers(0, 1 << 30))).astype(np.float64) @ orc.synth_gt_pose(kind, b, a)).astype(np.float32)
    return dict(rgb_s=rgb_s, d_s=d_s, rgb_t=rgb_t, d_t=d_t, levels=levels, method=int(rng.integers(0, 3)), guess=guess, cam=cam)


@pytest.mark.parametrize("seed", range(24))
def test_oracle_pinhole_equals_live_reference_on_random_cases(orc, seed):
    """Seeded random problems through the compiled reference HERE (one thread) and through the oracle."""
    from oracle import refbind
    if not refbind.available():
        pytest.skip("oracle/_ref not built and /root/reference absent")
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_reference_golden", os.path.join(GOLD, "make_reference_golden.py"))
    m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
    case = _random_pinhole_case(orc, seed)
    pinned = bool(seed & 1)
    live = json.loads(json.dumps(m.run_reference_pinhole(case, pinned)))
    _check_oracle_pinhole_against_reference(orc, case, live, pinned)


def test_pinhole_exponential_is_the_full_se3_exp(orc):
    """alignFrames calls CPose3D::exp(v) WITHOUT the pseudo flag (RPI.h:4375): t = V u.  First-order
    consistency with the Jacobian's twist (translation first) and agreement with scipy's expm."""
    from scipy.linalg import expm
    import ctypes as C
    rng = np.random.default_rng(5)
    L = orc.lib()
    L.orc_se3_exp.argtypes = [C.c_void_p, C.c_void_p]
    for _ in range(20):
        v = rng.uniform(-0.3, 0.3, 6)
        T = np.zeros(16)
        L.orc_se3_exp(v.ctypes.data_as(C.c_void_p), T.ctypes.data_as(C.c_void_p))
        T = T.reshape(4, 4).T
        w = v[3:]
        X = np.zeros((4, 4)); X[:3, :3] = [[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]]; X[:3, 3] = v[:3]
        assert np.allclose(T, expm(X), atol=1e-12)


# ============================================================================ GPU (CUDA path vs the oracle, PINNED)
def _gpu_ctx(r360, case, max_pairs=1):
    rows, cols = case["d_s"].shape
    p = r360.pinhole_params(n_levels=case["levels"], method=case["method"])
    ctx = r360.Context(rows, cols, 2, max_pairs, p)
    ctx.set_camera(*case["cam"])
    ctx.set_frames(0, np.stack([case["rgb_s"], case["rgb_t"]]), np.stack([case["d_s"], case["d_t"]]),
                   [r360.ROLE_SOURCE, r360.ROLE_TARGET])
    return ctx


def _check_eval(orc, ctx, src, trg, P, cam, level, T):
    eo = orc.error_pinhole(src, trg, level, T, P, cam)
    eg = ctx.eval_error_pinhole(0, 1, level, T)
    assert (eg["n_photo"], eg["n_depth"]) == (eo["n_photo"], eo["n_depth"]), (level, eg, eo)     # integer work: exact
    for k in ("photo", "depth"):
        assert abs(eg[k] - eo[k]) <= REL * abs(eo[k]) + 1e-30, (level, k, eg[k], eo[k])
    if np.isfinite(eo["error"]):
        assert abs(eg["error"] - eo["error"]) <= REL * abs(eo["error"])
    else:
        assert not np.isfinite(eg["error"])
    ho = orc.hessgrad_pinhole(src, trg, level, T, P, cam, accum=orc.ACC_STABLE)
    Hg, gg, nv = ctx.eval_hessgrad(0, 1, level, T)
    assert nv == ho["n_visible"], (level, nv, ho["n_visible"])
    Ho = ho["H"].astype(np.float64)
    dg = np.abs(np.diag(Ho)) + 1e-30
    sc = np.sqrt(np.outer(dg, dg))
    assert np.all(np.abs(Hg.astype(np.float64) - Ho) <= REL * sc), np.max(np.abs(Hg - Ho) / sc)
    gs = np.sqrt(dg * max(eo["photo"] + eo["depth"], 1e-30))
    assert np.all(np.abs(gg.astype(np.float64) - ho["g"]) <= REL * gs), np.max(np.abs(gg - ho["g"]) / gs)


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(refcases.PINHOLE_CASES))
def test_gpu_pinhole_evaluations(orc, r360, gold_pin, name):
    """errorPhotoICP + calcHessGrad at every level, at probe poses and at the reference's recorded final pose."""
    from util import small_pose
    case = refcases.make_pinhole_case(orc, name)
    orc.set_math(orc.MATH_PINNED)
    P, src, trg = _frames(orc, case)
    ctx = _gpu_ctx(r360, case)
    try:
        poses = [np.eye(4, dtype=np.float32), small_pose(0.01, -0.02, 0.015, 0.03, -0.02, 0.05),
                 np.array(gold_pin[name]["pinned"]["pose"], np.float32).reshape(4, 4)]
        if case["guess"] is not None:
            poses.append(np.asarray(case["guess"], np.float32))
        for level in range(case["levels"]):
            for T in poses:
                _check_eval(orc, ctx, src, trg, P, case["cam"], level, T)
    finally:
        ctx.close()


def _lm_primitives(orc):
    """The update rule of alignFrames from the bit-exact primitives the oracle and the kernels share
    (rgbd360_b200/csrc/gn_math.h): -> candidate(H21, g6, lam_or_None, pose_estim) = (pose 4x4 f32, update 6 f32, rank)."""
    import ctypes as C
    L = orc.lib()
    L.orc_se3_exp.argtypes = [C.c_void_p, C.c_void_p]

    def ptr(a):
        return a.ctypes.data_as(C.c_void_p)

    def full(H21):
        H = np.zeros((6, 6), np.float32); q = 0
        for a in range(6):
            for b in range(a, 6):
                H[a, b] = H[b, a] = H21[q]; q += 1
        return H

    def damp(H, lam):
        Hl = H.copy()
        for a in range(6):
            Hl[a, a] = np.float32(H[a, a] + np.float32(lam) * H[a, a])
        return Hl

    def candidate(H21, g6, lam_solve, lam_rank, pose_estim):
        H = full(np.asarray(H21, np.float32)); g = np.ascontiguousarray(g6, np.float32)
        rank = L.orc_rank6(ptr(np.ascontiguousarray(damp(H, lam_rank).T))) if lam_rank is not None else 6
        Hs = np.ascontiguousarray((damp(H, lam_solve) if lam_solve is not None else H).T)
        inv = np.zeros(36, np.float32); upd = np.zeros(6, np.float32)
        L.orc_inverse6(ptr(Hs), ptr(inv)); L.orc_solve_update(ptr(inv), ptr(g), ptr(upd))
        ud = upd.astype(np.float64); Td = np.zeros(16, np.float64)
        L.orc_se3_exp(ptr(ud), ptr(Td))
        Tf = Td.astype(np.float32); Pe = np.ascontiguousarray(np.asarray(pose_estim, np.float32).T).reshape(16)
        out = np.zeros(16, np.float32)
        L.orc_mat4_mul(ptr(Tf), ptr(Pe), ptr(out))
        return out.reshape(4, 4).T.copy(), upd, rank
    return candidate


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(refcases.PINHOLE_CASES))
def test_gpu_pinhole_align(orc, r360, gold_pin, name):
    """alignFrames (RPI.h:4254-4512) step by step.  Its accept test is `diff_error > 0` on an RMS of float
    precision, so the reference's own control flow depends on the order in which it sums (STABLE vs FAITHFUL
    accumulation of the oracle already differ in iteration counts and by millimetres on these cases); parity
    of the loop is therefore established per step: (1) every pose the GPU evaluated is replayed through the
    oracle at the same bits -- counters exact, sums 1e-4; (2) given the sums the GPU recorded, every decision
    and every next pose is what the reference's rule produces, BIT FOR BIT (shared gn_math.h primitives):
    Gauss-Newton candidate, accept on diff_error > 0 with lambda / 10, one damped retry with lambda * 10,
    rank test, loop condition, full SE(3) exponential."""
    from util import pose_err
    case = refcases.make_pinhole_case(orc, name)
    orc.set_math(orc.MATH_PINNED)
    P, src, trg = _frames(orc, case)
    ctx = _gpu_ctx(r360, case)
    candidate = _lm_primitives(orc)
    try:
        L = case["levels"]
        guess = None if case["guess"] is None else r360.pose_to_colmajor(case["guess"])[None]
        res_g, tr_g = ctx.register_pairs([0], [1], guess, trace=True)
        res_g = res_g[0]
        per = 2 * P.max_iters + 2
        pose_estim = np.eye(4, dtype=np.float32) if case["guess"] is None else np.asarray(case["guess"], np.float32)
        ill = False
        for lvl in range(L - 1, -1, -1):
            recs = [tr_g[lvl * per + k] for k in range(per) if tr_g[lvl * per + k].used]
            if ill:
                assert not recs
                continue
            for g in recs:                                                        # (1) evaluations
                T = np.array(g.pose, np.float32).reshape(4, 4).T
                eo = orc.error_pinhole(src, trg, lvl, T, P, case["cam"])
                assert (g.n_valid, g.n_valid_depth) == (eo["n_photo"], eo["n_depth"]), (lvl, g.it)
                assert abs(g.err2 - eo["photo"]) <= REL * abs(eo["photo"]) + 1e-30
                assert abs(g.err2_depth - eo["depth"]) <= REL * abs(eo["depth"]) + 1e-30
                assert orc.hessgrad_pinhole(src, trg, lvl, T, P, case["cam"])["n_visible"] == g.n_visible
            # (2) the state machine, replayed from the recorded sums
            def errf(g):
                with np.errstate(all="ignore"):
                    nd = np.float64(g.n_valid_depth)
                    return np.float64(np.float32(np.sqrt(np.float64(g.err2) / nd) + np.sqrt(np.float64(g.err2_depth) / nd)))
            assert recs and recs[0].accepted == 1
            assert np.array_equal(np.array(recs[0].pose, np.float32).reshape(4, 4).T, pose_estim)
            error, lam, it, k = errf(recs[0]), 0.01, 0, 1
            H21, g6 = np.array(recs[0].hessian, np.float32), np.array(recs[0].gradient, np.float32)
            upd, diff = np.ones(6, np.float32), error
            while it < P.max_iters and np.float64(np.sqrt(np.float32(np.sum(upd[:3] ** 2, dtype=np.float32) + np.sum(upd[3:] ** 2, dtype=np.float32)))) > P.tol_update \
                    and diff > P.tol_residual:
                cand, upd, rank = candidate(H21, g6, None, lam, pose_estim)
                if rank != 6:
                    ill = True
                    break
                assert k < len(recs), (lvl, k, "the loop condition holds but the GPU stopped")
                assert np.array_equal(np.array(recs[k].pose, np.float32).reshape(4, 4).T, cand), (lvl, k)
                diff = error - errf(recs[k])
                assert recs[k].accepted == int(diff > 0), (lvl, k)
                if diff > 0:
                    lam /= 10; pose_estim = cand; error = errf(recs[k]); it += 1
                    H21, g6 = np.array(recs[k].hessian, np.float32), np.array(recs[k].gradient, np.float32)
                    k += 1
                else:
                    k += 1
                    if diff < 0:                                                  # one damped retry, RPI.h:4399-4424
                        lam *= 10
                        cand, upd, _ = candidate(H21, g6, lam, None, pose_estim)
                        assert k < len(recs)
                        assert np.array_equal(np.array(recs[k].pose, np.float32).reshape(4, 4).T, cand), (lvl, k)
                        diff = error - errf(recs[k])
                        assert recs[k].accepted == int(diff > 0), (lvl, k)
                        if diff > 0:
                            pose_estim = cand; error = errf(recs[k]); it += 1
                            H21, g6 = np.array(recs[k].hessian, np.float32), np.array(recs[k].gradient, np.float32)
                        k += 1
            assert k == len(recs), (lvl, k, len(recs))
            if not ill:
                assert res_g["iters"][lvl] == it
        ref = gold_pin[name]["pinned"]
        assert (res_g["status"] != 0) == ill == ref["ill_posed"]
        assert np.array_equal(np.array(res_g["pose"], np.float32).reshape(4, 4).T, pose_estim)
        # against the recorded reference run (float accumulation): same basin, not the same bits
        ang, dist = pose_err(pose_estim, np.array(ref["pose"]).reshape(4, 4))
        assert ang <= 1e-2 and dist <= 2e-2, (ang, dist)
    finally:
        ctx.close()


@pytest.mark.gpu
def test_gpu_pinhole_batch_and_errors(orc, r360):
    """A batch of pinhole pairs == one-by-one runs; the camera matrix is mandatory; sphere contexts refuse it."""
    rows, cols, L, n = 120, 160, 3, 5
    cam = (131.25, 131.25, 79.5, 59.5)
    p = r360.pinhole_params(n_levels=L)
    ctx = r360.Context(rows, cols, 2 * n, n, p)
    frames = [orc.synth_pinhole_frame(0, k, rows, cols, *cam) for k in range(2 * n)]
    rgb = np.stack([f[0] for f in frames]); dep = np.stack([f[1] for f in frames])
    ctx.set_frames(0, rgb, dep, np.array([r360.ROLE_TARGET, r360.ROLE_SOURCE] * n, np.uint8))
    trg_idx = np.arange(0, 2 * n, 2, dtype=np.int32); src_idx = trg_idx + 1
    with pytest.raises(r360.R360Error, match="set_camera"):
        ctx.register_pairs(src_idx, trg_idx)
    ctx.set_camera(*cam)
    res = ctx.register_pairs(src_idx, trg_idx)
    one = r360.Context(rows, cols, 2, 1, p); one.set_camera(*cam)
    for k in range(n):
        one.set_frames(0, np.stack([rgb[2 * k + 1], rgb[2 * k]]), np.stack([dep[2 * k + 1], dep[2 * k]]))
        r1 = one.register_pairs([0], [1])[0]
        assert np.array_equal(r1["pose"], res[k]["pose"]) and list(r1["iters"]) == list(res[k]["iters"])
    one.close(); ctx.close()
    sph = r360.Context(64, 128, 2, 1, r360.default_params(n_levels=2))
    with pytest.raises(r360.R360Error):
        sph.set_camera(*cam)
    sph.close()
    with pytest.raises(r360.R360Error):
        r360.Context(rows, cols, 2, 1, r360.pinhole_params(n_levels=L, occlusion=1))
