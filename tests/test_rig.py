"""The 8-sensor rig: RegisterRGBD360::RegisterDensePhotoICP (RegisterRGBD360.h:344-520) and the per-sensor
calcPhotoICPError_robot / calcHessianGradient_robot (RPI.h:4905-5092, 5100-5407) it sums.

Pinned against the reference's own code: the two per-sensor functions belong to RegisterPhotoICP.h, which is compiled whole
(oracle/_ref/librpi_ref.so); the driver is cut out of RegisterRGBD360.h verbatim at build time and compiled in a scaffold
(oracle/ref_harness.cpp).  tests/golden/reference_rig.json records both (tests/golden/make_reference_golden.py rig).

What the reference does, and what is therefore asserted:
  * per sensor, the oracle equals the compiled reference BIT FOR BIT (error sum, float Hessian, float gradient), recorded and live;
  * the driver evaluates `new_error` at pose_estim (RegisterRGBD360.h:462, 488): with a reproducible summation no step is ever
    taken and it returns its guess with the rig's summed Hessian at the guess.  Upstream's own summation is an OpenMP reduction
    over the sensors whose order is the threads' arrival order, so `error - new_error` is 0 or +-1 ulp from run to run and some
    runs take an (unevaluated) step: the recording keeps how many of 12 runs returned the guess; the comparison is made on those;
  * `faithful = False` (the candidate is evaluated: the evident intent) is checked against the analytic ground truth;
  * only PHOTO_CONSISTENCY is defined upstream (the Hessian's depth row reads a matrix that is never assigned, RPI.h:5372-5374).
GPU: counters exact, sums 1e-4 (5e-4 on H against the reference's serial float accumulation over 8 x 76 800 pixels), the
faithful run returns the guess bit for bit, the fixed run's pose within 1e-4 of the oracle's.
"""
import json
import os
import numpy as np
import pytest
import refcases
from util import pose_err, upper21

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
REL = 1e-4


@pytest.fixture(scope="module")
def gold():
    with open(os.path.join(GOLD, "reference_rig.json")) as f:
        return json.load(f)["cases"]


def _frames(orc, case, P):
    src = [orc.Frame(case["rgb2"][s], case["d2"][s], P, False) for s in range(8)]
    trg = [orc.Frame(case["rgb1"][s], case["d1"][s], P, True) for s in range(8)]
    return src, trg


@pytest.mark.parametrize("name", list(refcases.RIG_CASES))
def test_oracle_rig_equals_recorded_reference(orc, gold, name):
    case = refcases.make_rig_case(orc, name)
    rec = gold[name]
    orc.lib().orc_set_threads(1)
    P = orc.rig_params(case["levels"])
    src, trg = _frames(orc, case, P)
    for s in range(8):
        for pr in rec["sensors"][s]:
            T = np.array(pr["pose"], np.float32).reshape(4, 4)
            e2, n = orc.error_robot(src[s], trg[s], pr["level"], T, case["Rt"][s], P, case["cam"])
            assert e2 == pr["error2"], (s, pr["level"])                                    # double sum at one thread: every bit
            hg = orc.hessgrad_robot(src[s], trg[s], pr["level"], T, case["Rt"][s], P, case["cam"])
            assert np.array_equal(hg["H"].T.astype(np.float64).ravel(), np.array(pr["H"])), (s, pr["level"])
            assert np.array_equal(hg["g"].astype(np.float64), np.array(pr["g"])), (s, pr["level"])
    # the driver: a recorded run of the verbatim function that returned its guess (see the module docstring)
    assert any(rec["driver_runs_returning_the_guess"]) and "driver" in rec
    res = orc.align_rig(src, trg, case["Rt"], case["guess"], P, case["cam"], faithful=True)
    guess = case["guess"] if case["guess"] is not None else np.eye(4, dtype=np.float32)
    assert res.status == 0 and rec["driver"]["ok"]
    assert np.array_equal(orc.pose_from(res.pose).astype(np.float64).ravel(), np.array(rec["driver"]["pose"]))
    assert np.array_equal(orc.pose_from(res.pose), guess)                                  # rigidTransf = the guess
    assert np.array_equal(np.array(res.hessian, np.float64).reshape(6, 6).T.ravel(), np.array(rec["driver"]["info"]))   # informationM
    assert list(res.iters)[:case["levels"]] == [0] * case["levels"]


def test_oracle_rig_equals_live_reference(orc):
    """Beyond the recording: other sensors / poses / levels through the compiled reference HERE (container only), and the
    verbatim driver run until it returns its guess."""
    from oracle import refbind
    if not refbind.available():
        pytest.skip("oracle/_ref not built and /root/reference absent")
    case = refcases.make_rig_case(orc, "rig_odo_L4")
    P = orc.rig_params(4)
    src, trg = _frames(orc, case, P)
    refbind.lib(False).ref_set_threads(1)
    orc.lib().orc_set_threads(1)
    rng = np.random.default_rng(3)
    for s in (1, 6):
        R = refbind.Reference(n_levels=4, pinned=False)
        R.set_camera(*case["cam"]); R.set_source(case["rgb2"][s], case["d2"][s]); R.set_target(case["rgb1"][s], case["d1"][s])
        for k in range(3):
            T = refcases.small_guess(int(rng.integers(0, 1 << 30)))
            lvl = int(rng.integers(0, 4))
            assert R.error_robot(lvl, T, case["Rt"][s], 0) == orc.error_robot(src[s], trg[s], lvl, T, case["Rt"][s], P, case["cam"])[0]
            Hr, gr = R.hessgrad_robot(lvl, T, case["Rt"][s], 0)
            ho = orc.hessgrad_robot(src[s], trg[s], lvl, T, case["Rt"][s], P, case["cam"])
            assert np.array_equal(Hr, ho["H"].T) and np.array_equal(gr, ho["g"])
        R.close()
    want = orc.align_rig(src, trg, case["Rt"], case["guess"], P, case["cam"], faithful=True)
    for attempt in range(12):
        a = refbind.rig_align(case["rgb1"], case["d1"], case["rgb2"], case["d2"], case["Rt"], case["guess"], 0)
        if np.array_equal(a["pose"], case["guess"]):
            assert a["ok"] and np.array_equal(a["info"], np.array(want.hessian).reshape(6, 6).T)
            return
    pytest.skip("12 runs of the reference driver all took an unevaluated step (its reduction order)")


def test_oracle_rig_fixed_mode_converges(orc):
    """faithful = False: the candidate pose is evaluated; the loop then converges to the analytic ground truth."""
    case = refcases.make_rig_case(orc, "rig_odo_L3_identity")
    P = orc.rig_params(3)
    src, trg = _frames(orc, case, P)
    res = orc.align_rig(src, trg, case["Rt"], None, P, case["cam"], faithful=False)
    assert res.status == 0 and sum(list(res.iters)[:3]) >= 2
    ang, dist = pose_err(orc.pose_from(res.pose), orc.synth_gt_pose(case["kind"], case["frames"][1], case["frames"][0]))
    assert ang < 2e-3 and dist < 5e-3, (ang, dist)


# ------------------------------------------------------------------------------------------------ GPU
def _ctx(r360, case, n_rig_frames=2, max_pairs=1):
    rows, cols = case["d1"].shape[1:]
    gp = r360.pinhole_params(n_levels=case["levels"], method=r360.PHOTO_CONSISTENCY)
    ctx = r360.Context(rows, cols, 8 * n_rig_frames, max_pairs, gp)
    ctx.set_camera(*case["cam"])
    ctx.set_frames(0, case["rgb1"], case["d1"], [r360.ROLE_TARGET] * 8)       # frame1: slots 0..7
    ctx.set_frames(8, case["rgb2"], case["d2"], [r360.ROLE_SOURCE] * 8)       # frame2: slots 8..15
    return ctx


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(refcases.RIG_CASES))
def test_gpu_rig_equals_reference(orc, r360, gold, name):
    case = refcases.make_rig_case(orc, name)
    rec = gold[name]
    P = orc.rig_params(case["levels"])
    src, trg = _frames(orc, case, P)
    ctx = _ctx(r360, case)
    try:
        guess = case["guess"] if case["guess"] is not None else np.eye(4, dtype=np.float32)
        # one evaluation of the rig == the sum over the sensors of the recorded reference functions
        for T in refcases.probe_poses(2) + [guess]:
            for lvl in (0, case["levels"] - 1):
                e_ref = 0.0; H_ref = np.zeros((6, 6)); g_ref = np.zeros(6); nv = 0; nt = 0
                for s in range(8):
                    pr = [q for q in rec["sensors"][s] if q["level"] == lvl and np.array_equal(np.array(q["pose"], np.float32).reshape(4, 4), np.asarray(T, np.float32))][0]
                    e_ref += pr["error2"]; H_ref += np.array(pr["H"]).reshape(6, 6); g_ref += np.array(pr["g"])
                    ho = orc.hessgrad_robot(src[s], trg[s], lvl, T, case["Rt"][s], P, case["cam"])
                    nv += ho["n_visible"]; nt += orc.error_robot(src[s], trg[s], lvl, T, case["Rt"][s], P, case["cam"])[1]
                ev = ctx.eval_rig(8, 0, lvl, T, case["Rt"])
                assert ev["n_visible"] == nv and ev["n_terms"] == nt                       # integer work: exact
                assert abs(ev["error2"] - e_ref) <= REL * e_ref, (lvl, ev["error2"], e_ref)
                sc = np.sqrt(np.outer(np.diag(H_ref), np.diag(H_ref)))
                assert np.all(np.abs(ev["H"] - H_ref) <= 5 * REL * sc), np.max(np.abs(ev["H"] - H_ref) / sc)
                gs = np.sqrt(np.diag(H_ref) * e_ref)
                assert np.all(np.abs(ev["g"] - g_ref) <= 5 * REL * gs)
        # the driver as upstream runs it: returns the guess bit for bit, informationM = the summed Hessian at the guess
        res = ctx.register_rig_pairs([8], [0], case["Rt"], None if case["guess"] is None else r360.pose_to_colmajor(case["guess"])[None],
                                     faithful=True)[0]
        assert res["status"] == 0
        assert np.array_equal(np.array(res["pose"], np.float32).reshape(4, 4).T, guess)
        assert list(res["iters"][:case["levels"]]) == [0] * case["levels"]
        Hr = np.array(rec["driver"]["info"]).reshape(6, 6)
        sc = np.sqrt(np.outer(np.diag(Hr), np.diag(Hr)))
        assert np.all(np.abs(np.array(res["hessian"], np.float64).reshape(6, 6) - Hr) <= 5 * REL * sc)
        again = ctx.register_rig_pairs([8], [0], case["Rt"], None if case["guess"] is None else r360.pose_to_colmajor(case["guess"])[None],
                                       faithful=True)[0]
        assert again.tobytes() == res.tobytes()                                            # reproducible sums: what makes diff_error 0
        # the intended loop (candidate evaluated): same iteration counts and pose as the oracle's run of the same rule
        fx = ctx.register_rig_pairs([8], [0], case["Rt"], None if case["guess"] is None else r360.pose_to_colmajor(case["guess"])[None],
                                    faithful=False)[0]
        o = orc.align_rig(src, trg, case["Rt"], case["guess"], P, case["cam"], faithful=False, accum=orc.ACC_STABLE)
        assert fx["status"] == o.status
        ang, dist = pose_err(np.array(fx["pose"], np.float32).reshape(4, 4).T, orc.synth_gt_pose(case["kind"], case["frames"][1], case["frames"][0]))
        assert ang < 2e-3 and dist < 5e-3, (ang, dist)
        if list(fx["iters"][:case["levels"]]) == list(o.iters)[:case["levels"]]:
            ang, dist = pose_err(np.array(fx["pose"], np.float32).reshape(4, 4).T, orc.pose_from(o.pose))
            assert ang <= 1e-4 and dist <= 1e-4, (ang, dist)
    finally:
        ctx.close()


@pytest.mark.gpu
def test_gpu_rig_errors_and_cpp_class(orc, r360, tmp_path):
    """Loud refusals (wrong projection / method, frames missing) and the C++ mirror of the class through the C ABI."""
    import subprocess
    case = refcases.make_rig_case(orc, "rig_odo_L3_identity")
    rows, cols = case["d1"].shape[1:]
    sph = r360.Context(rows, cols, 16, 1, r360.default_params(n_levels=3, n_sensors_mask=0))
    with pytest.raises(r360.R360Error):
        sph.register_rig_pairs([8], [0], case["Rt"])                                       # spherical context
    sph.close()
    pd = r360.Context(rows, cols, 16, 1, r360.pinhole_params(n_levels=3, method=r360.PHOTO_DEPTH))
    pd.set_camera(*case["cam"])
    with pytest.raises(r360.R360Error):
        pd.register_rig_pairs([8], [0], case["Rt"])                                        # undefined upstream
    pd.close()
    ctx = r360.Context(rows, cols, 16, 1, r360.pinhole_params(n_levels=3, method=r360.PHOTO_CONSISTENCY))
    ctx.set_camera(*case["cam"])
    with pytest.raises(r360.R360Error):
        ctx.register_rig_pairs([8], [0], case["Rt"])                                       # frames never set
    with pytest.raises(r360.R360Error):
        ctx.register_rig_pairs([9], [0], case["Rt"])                                       # 8 slots from 9 do not fit
    ctx.close()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = tmp_path / "rig_demo"
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(root, "include"),
                           os.path.join(root, "tests", "cpp", "rig_demo.cpp"), "-o", str(exe),
                           "-L", os.path.join(root, "rgbd360_b200"), "-lrgbd360_b200", "-Wl,-rpath," + os.path.join(root, "rgbd360_b200")])
    np.concatenate([case["rgb1"].ravel(), case["rgb2"].ravel()]).tofile(tmp_path / "rgb.bin")
    np.concatenate([case["d1"].ravel(), case["d2"].ravel()]).tofile(tmp_path / "depth.bin")
    np.ascontiguousarray(case["Rt"].transpose(0, 2, 1)).astype(np.float32).tofile(tmp_path / "rt.bin")
    out = subprocess.run([str(exe), str(tmp_path), str(rows), str(cols)], capture_output=True, text=True)
    assert out.returncode == 0 and "OK" in out.stdout, out.stdout + out.stderr
