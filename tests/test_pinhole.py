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
