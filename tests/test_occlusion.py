"""alignFrames360(..., occlusion = 1 / 2) -- errorPhotoICP_sphereOcc1 / Occ2 and
calcHessGrad_sphereOcc1 / Occ2 (RPI.h:3232-4249).

Upstream these loops run `#pragma omp parallel for` over a z-buffer every iteration reads and writes, so
their result depends on the thread schedule; the semantics pinned here (and implemented by the CUDA
path) are those of ONE thread: source pixels in index order.  tests/golden/reference_occlusion.json
holds what the compiled reference (oracle/_ref, both arithmetic variants, omp threads = 1) produces
on 7 cases x {Occ1, Occ2} (generator: tests/golden/make_reference_golden.py occ).

CPU tests: oracle == recording, bit for bit (pose, Hessian, gradient, SSO, iteration counts, the error
values and RMS terms of direct calls).  GPU tests (-m gpu): CUDA path vs the oracle in PINNED mode --
integer work (counts, numVisiblePixels) exact, sums 1e-4 relative, poses 1e-4.
"""
import json
import os
import numpy as np
import pytest
import refcases

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def gold_occ():
    with open(os.path.join(GOLD, "reference_occlusion.json")) as f:
        return json.load(f)["cases"]


def _eq_nan(a, b):
    """b: recorded float or None (= NaN)."""
    return (a != a) if b is None else (a == b)


def _frames(orc, case, occ):
    P = orc.default_params(n_levels=case["levels"], method=case["method"], std_photo=case["std_photo"], occlusion=occ)
    trg = orc.Frame(case["rgb_t"], case["d_t"], P, True)
    src = orc.Frame(case["rgb_s"], case["d_s"], P, False)
    return P, src, trg


OCC_CASES = ["synth_128x256_L3_pd", "synth_256x512_L4_pd_far", "synth_128x256_L3_holes", "synth_128x256_L3_depth",
             "synth_128x256_L3_photo", "loop_128x256_L3", "sample_pair_1920x320_L4"]


@pytest.mark.parametrize("pinned", [False, True], ids=["libm", "pinned"])
@pytest.mark.parametrize("occ", [1, 2])
@pytest.mark.parametrize("name", OCC_CASES)
def test_oracle_occlusion_equals_recorded_reference(orc, gold_occ, name, occ, pinned):
    case = refcases.make_case(orc, name)
    ref = gold_occ[name][str(occ)]["pinned" if pinned else "libm"]
    _check_oracle_occ_against_reference(orc, case, ref, occ, pinned)


def _check_oracle_occ_against_reference(orc, case, ref, occ, pinned):
    """Oracle vs one reference record (recorded or live) of an occlusion run, bit for bit at one thread."""
    orc.set_math(orc.MATH_PINNED if pinned else orc.MATH_LIBM)
    orc.lib().orc_set_threads(1)
    try:
        P, src, trg = _frames(orc, case, occ)
        L = case["levels"]
        res = orc.align(src, trg, case["guess"], P, accum=orc.ACC_FAITHFUL)
        assert list(res.iters)[:L] == ref["iters"]
        assert np.array_equal(orc.pose_from(res.pose).astype(np.float64).ravel(), np.array(ref["pose"]))
        if ref["H"] is not None:
            assert np.array_equal(np.array(res.hessian, np.float64).reshape(6, 6).T.ravel(), np.array(ref["H"]))
            assert np.array_equal(np.array(res.gradient, np.float64), np.array(ref["g"]))
            assert np.float32(res.sso) == np.float32(ref["sso"])
        # the four functions called directly at level 0
        for pr in ref["probes_level0"]:
            T = np.array(pr["pose"]).reshape(4, 4)
            eo = orc.error_occ(src, trg, 0, T, P)
            with np.errstate(all="ignore"):
                avp = np.sqrt(np.float64(eo["photo"]) / np.float64(eo["n_photo"] if occ == 1 else eo["n_depth"]))
                avd = np.sqrt(np.float64(eo["depth"]) / np.float64(eo["n_depth"]))
            assert _eq_nan(eo["error"], pr["error"]) and _eq_nan(avp, pr["av_photo"]) and _eq_nan(avd, pr["av_depth"])
            hg = orc.hessgrad(src, trg, 0, T, P, accum=orc.ACC_FAITHFUL)
            assert np.array_equal(hg["H"].astype(np.float64).ravel(), np.array(pr["H"]))
            assert np.array_equal(hg["g"].astype(np.float64), np.array(pr["g"]))
            assert np.float32(hg["n_visible"]) / np.float32(src.rows * src.cols) == np.float32(pr["sso"])
    finally:
        orc.set_math(orc.MATH_PINNED)


def test_live_reference_occlusion_matches_recording(orc, gold_occ):
    from oracle import refbind
    if not refbind.available():
        pytest.skip("oracle/_ref not built and /root/reference absent")
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_reference_golden", os.path.join(GOLD, "make_reference_golden.py"))
    m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
    for name in ["synth_128x256_L3_holes", "loop_128x256_L3"]:
        case = refcases.make_case(orc, name)
        for occ in (1, 2):
            live = m.run_reference_occ(case, True, occ)
            assert json.loads(json.dumps(live)) == gold_occ[name][str(occ)]["pinned"], (name, occ)


@pytest.mark.parametrize("seed", range(24))
def test_oracle_occlusion_equals_live_reference_on_random_cases(orc, seed):
    """Seeded random problems through the compiled reference HERE (one thread) and through the oracle."""
    from oracle import refbind
    if not refbind.available():
        pytest.skip("oracle/_ref not built and /root/reference absent")
    import importlib.util
    from test_reference import _random_case
    spec = importlib.util.spec_from_file_location("make_reference_golden", os.path.join(GOLD, "make_reference_golden.py"))
    m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
    case = _random_case(orc, 500 + seed)
    occ, pinned = 1 + (seed & 1), bool(seed & 2)
    live = json.loads(json.dumps(m.run_reference_occ(case, pinned, occ)))
    _check_oracle_occ_against_reference(orc, case, live, occ, pinned)


def test_occlusion_semantics_properties(orc):
    """Properties of the single-thread semantics the GPU implementation relies on (a z-test passer is a
    non-strict prefix maximum of 1/|p| among the candidates of its target texel, in source order)."""
    case = refcases.make_case(orc, "synth_256x512_L4_pd_far")
    orc.set_math(orc.MATH_PINNED)
    P0, src, trg = _frames(orc, case, 0)
    T = np.eye(4, dtype=np.float32)
    ri, ci, vp, vd = orc.warp(src, trg, 0, T, P0)
    rows, cols = src.rows, src.cols
    inb = (ri >= 0) & (ri < rows) & (ci < cols) & (ri != np.iinfo(np.int32).min)
    n_inb = int(inb.sum())
    # Occ1 H/g: the z-buffer is indexed by the source pixel, nothing is ever occluded -> numVisible = all in-bounds
    P1, _, _ = _frames(orc, case, 1)
    h1 = orc.hessgrad(src, trg, 0, T, P1)
    assert h1["n_visible"] == n_inb
    # ... and the depth-saliency `continue` drops the photo row too: never more photo rows than the regular function
    h0 = orc.hessgrad(src, trg, 0, T, P0)
    assert h1["n_photo"] <= h0["n_photo"] and h1["n_depth"] == h0["n_depth"]
    # Occ2 H/g counts DISTINCT target texels among the pixels that pass the 0.3 m gate
    P2, _, _ = _frames(orc, case, 2)
    h2 = orc.hessgrad(src, trg, 0, T, P2)
    assert h2["n_visible"] <= n_inb
    # Occ1 error: counters count passers (>= distinct texels), residual sums hold one value per texel
    e1 = orc.error_occ(src, trg, 0, T, P1)
    assert e1["n_photo"] > 0 and e1["n_depth"] > 0 and np.isfinite(e1["error"])


# ============================================================================ GPU (CUDA path vs the oracle, PINNED)
REL = 1e-4            # north-star tolerance on sums
POSE_TOL = 1e-4       # rad / m


def _gpu_ctx(r360, case, occ, max_pairs=1):
    rows, cols = case["d_s"].shape
    p = r360.default_params(n_levels=case["levels"], method=case["method"], std_photo=np.float32(case["std_photo"]),
                            occlusion=occ)
    ctx = r360.Context(rows, cols, 2, max_pairs, p)
    ctx.set_frames(0, np.stack([case["rgb_s"], case["rgb_t"]]), np.stack([case["d_s"], case["d_t"]]),
                   [r360.ROLE_SOURCE, r360.ROLE_TARGET])
    return ctx


def _check_eval(orc, ctx, src, trg, P, level, T, occ):
    """One evaluation: counters and numVisiblePixels exact, sums / Hessian / gradient within 1e-4."""
    eo = orc.error_occ(src, trg, level, T, P)
    eg = ctx.eval_error_occ(0, 1, level, T)
    assert eg["n_depth"] == eo["n_depth"], (level, eg, eo)
    if occ == 1:
        assert eg["n_photo"] == eo["n_photo"], (level, eg, eo)
    for k in ("photo", "depth"):
        assert abs(eg[k] - eo[k]) <= REL * abs(eo[k]) + 1e-30, (level, k, eg[k], eo[k])
    if np.isfinite(eo["error"]):
        assert abs(eg["error"] - eo["error"]) <= REL * abs(eo["error"])
    else:
        assert not np.isfinite(eg["error"])
    ho = orc.hessgrad(src, trg, level, T, P, accum=orc.ACC_STABLE)
    Hg, gg, nv = ctx.eval_hessgrad(0, 1, level, T)
    assert nv == ho["n_visible"], (level, nv, ho["n_visible"])
    Ho = ho["H"].astype(np.float64)
    dg = np.abs(np.diag(Ho)) + 1e-30
    sc = np.sqrt(np.outer(dg, dg))
    assert np.all(np.abs(Hg.astype(np.float64) - Ho) <= REL * sc), np.max(np.abs(Hg - Ho) / sc)
    rs = max(eo["photo"] + eo["depth"], 1e-30)
    # |g_a| <= sqrt(H_aa * sum r^2) over the rows of the H-set; the error sums are over the E-set,
    # so scale with the larger of the two to stay a relative bound
    gs = np.sqrt(dg * max(rs, float(np.max(np.abs(ho["g"])) ** 2 / dg.max())))
    assert np.all(np.abs(gg.astype(np.float64) - ho["g"]) <= REL * gs), np.max(np.abs(gg - ho["g"]) / gs)


GPU_CASES = ["synth_128x256_L3_pd", "synth_256x512_L4_pd_far", "synth_128x256_L3_holes", "synth_128x256_L3_depth",
             "synth_128x256_L3_photo", "loop_128x256_L3", "sample_pair_1920x320_L4"]


@pytest.mark.gpu
@pytest.mark.parametrize("occ", [1, 2])
@pytest.mark.parametrize("name", GPU_CASES)
def test_gpu_occlusion_evaluations(orc, r360, gold_occ, name, occ):
    """errorPhotoICP_sphereOccN + calcHessGrad_sphereOccN at every level, at probe poses and at the
    reference's recorded final pose (many-to-one warps: several source pixels per target texel)."""
    from util import small_pose
    case = refcases.make_case(orc, name)
    orc.set_math(orc.MATH_PINNED)
    P, src, trg = _frames(orc, case, occ)
    ctx = _gpu_ctx(r360, case, occ)
    try:
        poses = [np.eye(4, dtype=np.float32), small_pose(0.01, -0.02, 0.015, 0.03, -0.02, 0.05),
                 np.array(gold_occ[name][str(occ)]["pinned"]["pose"], np.float32).reshape(4, 4)]
        if case["guess"] is not None:
            poses.append(np.asarray(case["guess"], np.float32))
        for level in range(case["levels"]):
            for T in poses:
                _check_eval(orc, ctx, src, trg, P, level, T, occ)
    finally:
        ctx.close()


@pytest.mark.gpu
@pytest.mark.parametrize("occ", [1, 2])
@pytest.mark.parametrize("name", GPU_CASES)
def test_gpu_occlusion_align(orc, r360, gold_occ, name, occ):
    """alignFrames360(guess, method, occlusion): every pose the GPU evaluated is replayed through the
    oracle at the same bits (counters exact, sums 1e-4); control flow and final pose against the oracle's
    own run and against the recorded reference run (1e-4 rad / 1e-4 m)."""
    from util import pose_err
    case = refcases.make_case(orc, name)
    orc.set_math(orc.MATH_PINNED)
    P, src, trg = _frames(orc, case, occ)
    ctx = _gpu_ctx(r360, case, occ)
    try:
        L = case["levels"]
        guess = None if case["guess"] is None else r360.pose_to_colmajor(case["guess"])[None]
        res_g, tr_g = ctx.register_pairs([0], [1], guess, trace=True)
        res_g = res_g[0]
        res_o, tr_o = orc.align(src, trg, case["guess"], P, accum=orc.ACC_STABLE, trace=True)
        per = P.max_iters + 2
        n_rec = 0
        for lvl in range(L):
            for k in range(per):
                g = tr_g[lvl * per + k]
                if not g.used:
                    continue
                n_rec += 1
                T = np.array(g.pose, np.float32).reshape(4, 4).T
                eo = orc.error_occ(src, trg, lvl, T, P)                       # replay, same bits
                assert g.n_valid_depth == eo["n_depth"], (lvl, k)
                assert g.n_valid == (eo["n_photo"] if occ == 1 else eo["n_depth"]), (lvl, k)
                assert abs(g.err2 - eo["photo"]) <= REL * abs(eo["photo"]) + 1e-30, (lvl, k)
                assert abs(g.err2_depth - eo["depth"]) <= REL * abs(eo["depth"]) + 1e-30, (lvl, k)
                ho = orc.hessgrad(src, trg, lvl, T, P)
                assert ho["n_visible"] == g.n_visible, (lvl, k)
        assert n_rec >= L
        # control flow + pose: the oracle's own run, and the recorded reference (single thread) run
        ref = gold_occ[name][str(occ)]["pinned"]
        assert list(res_g["iters"][:L]) == list(res_o.iters)[:L]
        Tg = np.array(res_g["pose"], np.float32).reshape(4, 4).T
        ang, dist = pose_err(Tg, orc.pose_from(res_o.pose))
        assert ang <= POSE_TOL and dist <= POSE_TOL, (ang, dist)
        if list(res_o.iters)[:L] == ref["iters"]:
            ang, dist = pose_err(Tg, np.array(ref["pose"]).reshape(4, 4))
            # The recording was taken with the reference's 27 FLOAT accumulators (RPI.h:3605-3606); the GPU
            # (and the oracle's STABLE mode it is held to 1e-4 against, above) sums in double.  On the real
            # sample pair (30 accepted steps, 28 % invalid depth) that alone moves the reference's result by
            # 1.5e-4 m -- its sensitivity to its own summation order, see test_sample_pair_summation_order_sensitivity.
            tol = 5e-4 if name.startswith("sample_pair") else POSE_TOL
            assert ang <= tol and dist <= tol, (ang, dist)
    finally:
        ctx.close()


@pytest.mark.gpu
def test_gpu_occlusion_batches_and_host_pairs(orc, r360):
    """More pairs than one scratch batch (64): results independent of the batching, equal to single runs;
    the pipelined host entry point goes through the same path."""
    rows, cols, L, n = 64, 128, 2, 70
    p = r360.default_params(n_levels=L, occlusion=2)
    ctx = r360.Context(rows, cols, 2 * n, n, p)
    frames = [orc.synth_frame(0, k, rows, cols) for k in range(2 * n)]
    rgb = np.stack([f[0] for f in frames]); dep = np.stack([f[1] for f in frames])
    roles = np.array([r360.ROLE_TARGET, r360.ROLE_SOURCE] * n, np.uint8)
    ctx.set_frames(0, rgb, dep, roles)
    trg_idx = np.arange(0, 2 * n, 2, dtype=np.int32); src_idx = trg_idx + 1
    res = ctx.register_pairs(src_idx, trg_idx)
    one = r360.Context(rows, cols, 2, 1, p)
    for k in (0, 63, 64, 69):
        one.set_frames(0, np.stack([rgb[2 * k + 1], rgb[2 * k]]), np.stack([dep[2 * k + 1], dep[2 * k]]))
        r1 = one.register_pairs([0], [1])[0]
        assert np.array_equal(r1["pose"], res[k]["pose"]) and list(r1["iters"]) == list(res[k]["iters"])
    one.close()
    res_h = ctx.register_host_pairs(rgb, dep, n)
    assert np.array_equal(res_h["pose"], res["pose"])
    ctx.close()
