"""Runs the REFERENCE ITSELF (oracle/_ref/: /root/reference/include/RegisterPhotoICP.h compiled
against oracle/refshim/, both arithmetic variants) on the cases of tests/refcases.py and records its
outputs as tests/golden/reference_outputs.json.  Only runnable where /root/reference exists.

    python tests/golden/make_reference_golden.py          # both recordings
    python tests/golden/make_reference_golden.py sphere   # reference_outputs.json only
    python tests/golden/make_reference_golden.py occ      # reference_occlusion.json only
    python tests/golden/make_reference_golden.py pinhole  # reference_pinhole.json only
"""
import hashlib
import json
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import orc, refbind   # noqa: E402
import refcases                   # noqa: E402


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def canonical_lut(lut):
    """The reference only writes x = INVALID_POINT for an invalid pixel (RPI.h:4585) and leaves y, z
    as whatever the vector held before; zero them before hashing."""
    lut = lut.copy()
    lut[lut[:, 0] == -10000.0, 1:] = 0
    return lut


def run_reference(case, pinned, threads=1):
    refbind.lib(pinned).ref_set_threads(threads)
    R = refbind.Reference(n_levels=case["levels"], std_photo=case["std_photo"], pinned=pinned)
    R.set_source(case["rgb_s"], case["d_s"])
    R.set_target(case["rgb_t"], case["d_t"])
    a = R.align(case["guess"], case["method"])
    out = dict(pose=a["pose"].astype(np.float64).ravel().tolist(), H=a["H"].astype(np.float64).ravel().tolist(),
               g=a["g"].astype(np.float64).tolist(), sso=a["sso"], iters=a["iters"].tolist(),
               trace_err2=a["err2"].tolist(), trace_n_valid=a["n_valid"].tolist(), ill_posed=bool(a["ill_posed"]))
    if len(a["err2"]) <= case["levels"] and not a["ill_posed"]:
        # no loop body ever ran: calcHessGrad_sphere was never called and the getters return the
        # class's UNINITIALISED hessian / gradient / SSO members (RPI.h:112-115, no ctor init)
        out["H"] = out["g"] = out["sso"] = None
    planes = {}
    for l in range(case["levels"]):
        for which, tag in ((0, "src"), (1, "trg")):
            for k, v in R.level(which, l).items():
                if k in ("ggx", "ggy", "dgx", "dgy"):
                    # The reference zeroes the sensor-joint columns inside alignFrames360, level by level
                    # (RPI.h:4537-4549); levels it never reaches (ILL-POSED return) stay unmasked.  The
                    # product applies the same (idempotent) mask when the target pyramid is built, so the
                    # recording is taken with the mask applied at every level.
                    ws = v.shape[1] // 8
                    for q in range(1, 8):
                        v[:, q * ws - 1:q * ws + 1] = 0
                planes[f"{tag}{l}_{k}"] = digest(v)
    out["planes_sha"] = planes
    # after an ILL-POSED return the LUT is the one of the level where the run stopped, not level 0
    out["lut0_sha"] = None if a["ill_posed"] else digest(canonical_lut(R.lut()))
    probes = []
    for T in ([] if a["ill_posed"] else refcases.probe_poses()):
        e, e2, n = R.error(0, T, case["method"])
        H, g, sso = R.hessgrad(0, T, case["method"])
        probes.append(dict(err2=e2, n_valid=n, H=H.astype(np.float64).ravel().tolist(), g=g.astype(np.float64).tolist(), sso=sso))
    out["probes_level0"] = probes
    R.close()
    return out


# ---- alignFrames360(..., occlusion = 1 / 2) and the *_sphereOcc1 / Occ2 functions.
# Upstream these are order-dependent under OpenMP; the recording is taken with ONE thread (source order).
OCC_CASES = ["synth_128x256_L3_pd", "synth_256x512_L4_pd_far", "synth_128x256_L3_holes", "synth_128x256_L3_depth",
             "synth_128x256_L3_photo", "loop_128x256_L3", "sample_pair_1920x320_L4"]


def _f(x):
    """JSON-safe float (NaN -> None)."""
    x = float(x)
    return None if x != x else x


def run_reference_occ(case, pinned, occ):
    refbind.lib(pinned).ref_set_threads(1)
    R = refbind.Reference(n_levels=case["levels"], std_photo=case["std_photo"], pinned=pinned)
    R.set_source(case["rgb_s"], case["d_s"])
    R.set_target(case["rgb_t"], case["d_t"])
    a = R.align(case["guess"], case["method"], occ)
    out = dict(pose=a["pose"].astype(np.float64).ravel().tolist(), H=a["H"].astype(np.float64).ravel().tolist(),
               g=a["g"].astype(np.float64).tolist(), sso=a["sso"], iters=a["iters"].tolist(), ill_posed=bool(a["ill_posed"]))
    if not any(a["iters"]):
        # the loop body may never have run (e.g. occlusion 1 with a single cost term: 0/0 = NaN error):
        # the getters then return uninitialised members
        out["H"] = out["g"] = out["sso"] = None
    probes = []
    for T in refcases.probe_poses() + [a["pose"]]:
        e, avp, avd = R.error_occ(0, T, case["method"], occ)
        H, g, sso = R.hessgrad_occ(0, T, case["method"], occ)
        probes.append(dict(pose=np.asarray(T, np.float64).ravel().tolist(), error=_f(e), av_photo=_f(avp), av_depth=_f(avd),
                           H=H.astype(np.float64).ravel().tolist(), g=g.astype(np.float64).tolist(), sso=sso))
    out["probes_level0"] = probes
    R.close()
    return out


def main_occ():
    gold = {"_how": "oracle/_ref (reference header + refshim), OMP threads = 1, alignFrames360 occlusion = 1 / 2 and "
                    "errorPhotoICP_sphereOcc{1,2} / calcHessGrad_sphereOcc{1,2} called directly; see this script",
            "cases": {}}
    for name in OCC_CASES:
        case = refcases.make_case(orc, name)
        gold["cases"][name] = {}
        for occ in (1, 2):
            gold["cases"][name][str(occ)] = {"libm": run_reference_occ(case, False, occ),
                                             "pinned": run_reference_occ(case, True, occ)}
            c = gold["cases"][name][str(occ)]
            print(name, "occ", occ, "iters", c["libm"]["iters"], c["pinned"]["iters"])
    with open(os.path.join(HERE, "reference_occlusion.json"), "w") as f:
        json.dump(gold, f, indent=0)


# ---- the pinhole alignFrames / errorPhotoICP / calcHessGrad (RPI.h:4254, 560, 776).
# H and g are accumulated pixel by pixel in float under `omp critical`: recorded with ONE thread.
def run_reference_pinhole(case, pinned):
    refbind.lib(pinned).ref_set_threads(1)
    R = refbind.Reference(n_levels=case["levels"], pinned=pinned)
    R.set_camera(*case["cam"])
    R.set_source(case["rgb_s"], case["d_s"])
    R.set_target(case["rgb_t"], case["d_t"])
    a = R.align_pinhole(case["guess"], case["method"])
    out = dict(pose=a["pose"].astype(np.float64).ravel().tolist(), H=a["H"].astype(np.float64).ravel().tolist(),
               g=a["g"].astype(np.float64).tolist(), iters=a["iters"].tolist(), ill_posed=bool(a["ill_posed"]))
    probes = []
    if not a["ill_posed"]:           # after an ILL-POSED return the LUT is the one of the level where the run stopped
        for T in refcases.probe_poses() + [a["pose"]]:
            e, avp, avd = R.error_pinhole(0, T, case["method"])
            H, g = R.hessgrad_pinhole(0, T, case["method"])
            probes.append(dict(pose=np.asarray(T, np.float64).ravel().tolist(), error=_f(e), av_photo=_f(avp), av_depth=_f(avd),
                               H=H.astype(np.float64).ravel().tolist(), g=g.astype(np.float64).tolist()))
        # The loop body runs at a level only if the error at the level's starting pose is finite (the `while`
        # compares diff_error = error).  When it is NaN (0/0: no valid pixel) at EVERY level the pose never leaves the
        # guess, calcHessGrad is never called and the getters return the class's uninitialised members.
        start = case["guess"] if case["guess"] is not None else np.eye(4, dtype=np.float32)
        if not any(a["iters"]) and not any(np.isfinite(R.error_pinhole(l, start, case["method"])[0]) for l in range(case["levels"])):
            out["H"] = out["g"] = None
    out["probes_level0"] = probes
    R.close()
    return out


def main_pinhole():
    gold = {"_how": "oracle/_ref (reference header + refshim), OMP threads = 1, setCameraMatrix + alignFrames (pinhole) and "
                    "errorPhotoICP / calcHessGrad called directly; see this script", "cases": {}}
    for name in refcases.PINHOLE_CASES:
        case = refcases.make_pinhole_case(orc, name)
        gold["cases"][name] = {"libm": run_reference_pinhole(case, False), "pinned": run_reference_pinhole(case, True)}
        c = gold["cases"][name]
        print(name, "iters", c["libm"]["iters"], c["pinned"]["iters"], "ill", c["pinned"]["ill_posed"])
    with open(os.path.join(HERE, "reference_pinhole.json"), "w") as f:
        json.dump(gold, f, indent=0)


def sample_pair_other_branch(case, one_thread_iters, pinned=True):
    """Config #1 sits on a knife edge of the accept test (RPI.h:4715) at level 0: with its 27 FLOAT accumulators summed
    in another order (another OpenMP thread count) the reference takes 1 accepted step there instead of 10 and ends
    ~1e-3 rad / 1 cm away from its own one-thread answer.  Records one run of the reference on that other branch (the
    one a well-conditioned accumulation -- the oracle's STABLE mode, the GPU's wide sums -- follows)."""
    for rep in range(4):
        for th in (3, 8, 4, 2, 5, 6, 7):
            refbind.lib(pinned).ref_set_threads(th)
            try:
                R = refbind.Reference(n_levels=case["levels"], std_photo=case["std_photo"], pinned=pinned)
                R.set_source(case["rgb_s"], case["d_s"]); R.set_target(case["rgb_t"], case["d_t"])
                a = R.align(case["guess"], case["method"])
                R.close()
            finally:
                refbind.lib(pinned).ref_set_threads(1)
            if a["iters"].tolist() == [1, 10, 10, 7]:
                return dict(threads=th, pose=a["pose"].astype(np.float64).ravel().tolist(),
                            H=a["H"].astype(np.float64).ravel().tolist(), g=a["g"].astype(np.float64).tolist(), sso=a["sso"],
                            iters=a["iters"].tolist(), trace_err2=a["err2"].tolist(), trace_n_valid=a["n_valid"].tolist())
    raise RuntimeError("no thread count reproduced the 1-step branch of the sample pair on this host")


def sample_pair_self_spread(case, pinned=True, reps=2):
    """The reference against ITSELF on config #1: runs at 2..8 OpenMP threads (its float accumulators are then summed
    in other orders).  Recorded as evidence for the tolerance of the GPU test on this pair: iteration counts and poses."""
    runs = []
    for rep in range(reps):
        for th in (2, 3, 4, 5, 6, 7, 8):
            refbind.lib(pinned).ref_set_threads(th)
            try:
                R = refbind.Reference(n_levels=case["levels"], std_photo=case["std_photo"], pinned=pinned)
                R.set_source(case["rgb_s"], case["d_s"]); R.set_target(case["rgb_t"], case["d_t"])
                a = R.align(case["guess"], case["method"])
                R.close()
            finally:
                refbind.lib(pinned).ref_set_threads(1)
            runs.append(dict(threads=th, iters=a["iters"].tolist(), pose=a["pose"].astype(np.float64).ravel().tolist()))
    return runs


# ---- SURVEY 8f row 4, the rig: calcPhotoICPError_robot / calcHessianGradient_robot per sensor (one thread) and the
# verbatim RegisterRGBD360::RegisterDensePhotoICP.  The driver's accept decisions depend on the arrival order of its OpenMP
# threads (error and new_error are both evaluated at pose_estim and summed by a reduction): recorded is a run whose
# returned pose IS the guess (no step taken at any level) -- the outcome of a reproducible summation.
def run_reference_rig(case):
    refbind.lib(False).ref_set_threads(1)
    out = dict(sensors=[])
    guess = case["guess"] if case["guess"] is not None else np.eye(4, dtype=np.float32)
    for s in range(8):
        R = refbind.Reference(n_levels=case["levels"], pinned=False)
        R.set_camera(*case["cam"])
        R.set_source(case["rgb2"][s], case["d2"][s]); R.set_target(case["rgb1"][s], case["d1"][s])
        probes = []
        for T in refcases.probe_poses(2) + [guess]:
            for lvl in (0, case["levels"] - 1):
                H, g = R.hessgrad_robot(lvl, T, case["Rt"][s], 0)
                probes.append(dict(level=lvl, pose=np.asarray(T, np.float64).ravel().tolist(), error2=R.error_robot(lvl, T, case["Rt"][s], 0),
                                   H=H.astype(np.float64).ravel().tolist(), g=g.astype(np.float64).tolist()))
        out["sensors"].append(probes)
        R.close()
    runs = []
    for k in range(12):
        a = refbind.rig_align(case["rgb1"], case["d1"], case["rgb2"], case["d2"], case["Rt"], case["guess"], 0)
        runs.append(bool(np.array_equal(a["pose"], guess)))
        if runs[-1] and "driver" not in out:
            out["driver"] = dict(ok=a["ok"], pose=a["pose"].astype(np.float64).ravel().tolist(), info=a["info"].astype(np.float64).ravel().tolist())
    out["driver_runs_returning_the_guess"] = runs
    return out


def main_rig():
    gold = {"_how": "oracle/_ref/librpi_ref.so: calcPhotoICPError_robot / calcHessianGradient_robot (RPI.h:4905, 5100) per sensor at one "
                    "thread, and RegisterRGBD360::RegisterDensePhotoICP cut out verbatim at build time; see this script", "cases": {}}
    for name in refcases.RIG_CASES:
        gold["cases"][name] = run_reference_rig(refcases.make_rig_case(orc, name))
        print(name, "driver runs returning the guess:", gold["cases"][name]["driver_runs_returning_the_guess"])
    with open(os.path.join(HERE, "reference_rig.json"), "w") as f:
        json.dump(gold, f, indent=0)


def main_stitch():
    """Frame360::stitchSphericalImage / stitchImage + Calib360 as the reference wrote them
    (oracle/ref_stitch_harness.cpp) on the raw sensor images of samples/sphere_images_1.bin (tests/golden/
    frame360_raw_1.npz): digests of the sphere images for both trig builds, the Rt_inv Calib360 computed."""
    z = np.load(os.path.join(HERE, "frame360_raw_1.npz"))
    gold = {"_how": "oracle/_ref/librpi_ref_stitch[_pinned].so: Calib360.h whole, the two Frame360 member functions verbatim "
                    "(cut out of Frame360.h at build time), refshim third-party stand-ins; see this script",
            "camera": list(refbind.stitch_camera())}
    for pinned in (False, True):
        rgb, d, Rt_inv = refbind.stitch(z["rgb"], z["depth"], pinned=pinned)
        gold["pinned" if pinned else "libm"] = dict(rgb_sha=digest(rgb), depth_sha=digest(d), rows=int(d.shape[0]), cols=int(d.shape[1]),
                                                    valid_depth=int((d > 0).sum()))
        gold["Rt_inv"] = Rt_inv.astype(np.float64).reshape(8, 16).tolist()          # row-major 4x4 per sensor
        print("stitch", "pinned" if pinned else "libm", gold["pinned" if pinned else "libm"])
    with open(os.path.join(HERE, "reference_stitch.json"), "w") as f:
        json.dump(gold, f, indent=0)


def main():
    gold = {"_how": "oracle/_ref (reference header + refshim), OMP threads = 1; see this script", "cases": {}}
    for name in refcases.CASES:
        case = refcases.make_case(orc, name)
        gold["cases"][name] = {"libm": run_reference(case, False), "pinned": run_reference(case, True)}
        c = gold["cases"][name]
        print(name, "iters", c["libm"]["iters"], c["pinned"]["iters"], "n_valid[0]", c["libm"]["trace_n_valid"][:1], flush=True)
        if name.startswith("sample_pair"):
            c["pinned_branch_1_10_10_7"] = sample_pair_other_branch(case, c["pinned"]["iters"])
            print("   other branch at", c["pinned_branch_1_10_10_7"]["threads"], "threads")
            c["pinned_multithread_runs"] = sample_pair_self_spread(case)
    with open(os.path.join(HERE, "reference_outputs.json"), "w") as f:
        json.dump(gold, f, indent=0)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "sphere":
        main()
    elif len(sys.argv) > 1 and sys.argv[1] == "stitch":
        main_stitch()
    elif len(sys.argv) > 1 and sys.argv[1] == "rig":
        main_rig()
    elif len(sys.argv) > 1 and sys.argv[1] == "occ":
        main_occ()
    elif len(sys.argv) > 1 and sys.argv[1] == "pinhole":
        main_pinhole()
    else:
        main()
        main_occ()
        main_pinhole()
        main_stitch()
        main_rig()
