"""GPU parity on seeded RANDOM problems (the generator of tests/test_reference.py, on which the oracle is bit-identical
to the compiled reference): image size and aspect ratio, 1-4 levels, the three cost functions, either scene, holes and
colour noise, identity / perturbed / ground-truth-based guesses.

Checks that do not depend on two independent runs taking the same knife-edge decisions: every plane bit for bit; every
pose the GPU evaluated replayed through the oracle at the same bits (counters exact, sums 1e-4); every decision of the
GPU's loop re-derived from the sums it recorded (RPI.h:4611, 4715); and, when the oracle's own run took the same number
of steps, the final pose within 1e-4.
"""
import numpy as np
import pytest
from util import pose_err, upper21
from test_reference import _random_case

pytestmark = pytest.mark.gpu
REL = 1e-4


@pytest.mark.parametrize("seed", range(24))
def test_random_problem_replayed_through_the_oracle(orc, r360, seed):
    case = _random_case(orc, 2000 + seed)
    L, method = case["levels"], case["method"]
    rows, cols = case["d_s"].shape
    orc.set_math(orc.MATH_PINNED)
    P = orc.default_params(n_levels=L, method=method, std_photo=case["std_photo"])
    trg = orc.Frame(case["rgb_t"], case["d_t"], P, True)
    src = orc.Frame(case["rgb_s"], case["d_s"], P, False)
    ctx = r360.Context(rows, cols, 2, 1, r360.default_params(n_levels=L, method=method, std_photo=case["std_photo"]))
    try:
        ctx.set_frames(0, np.stack([case["rgb_s"], case["rgb_t"]]), np.stack([case["d_s"], case["d_t"]]),
                       [r360.ROLE_SOURCE, r360.ROLE_TARGET])
        # a1-a5: planes bit for bit
        for l in range(L):
            g_t = ctx.dump_level(1, l)
            for k, v in trg.level(l).items():
                assert np.array_equal(np.asarray(g_t[k]).view(np.uint32), np.asarray(v, np.float32).view(np.uint32)), (l, k)
        guess = None if case["guess"] is None else r360.pose_to_colmajor(case["guess"])[None]
        res_g, tr_g = ctx.register_pairs([0], [1], guess, trace=True)
        res_g = res_g[0]
        per = P.max_iters + 2
        for lvl in range(L - 1, -1, -1):
            recs = [tr_g[lvl * per + k] for k in range(per) if tr_g[lvl * per + k].used]
            assert recs, lvl
            err_prev = None
            for g in recs:
                pose_g = np.array(g.pose, np.float32).reshape(4, 4).T
                e2r, nvr = orc.error(src, trg, lvl, pose_g, P)                    # replay at the same bits
                assert nvr == g.n_valid, (lvl, g.it)                              # integer work: exact
                if nvr:
                    assert abs(g.err2 - e2r) <= REL * abs(e2r), (lvl, g.it)
                if g.used & 2:
                    hr = orc.hessgrad(src, trg, lvl, pose_g, P)
                    assert hr["n_visible"] == g.n_visible
                    dg = np.diag(hr["H"]).astype(np.float64)
                    sc = upper21(np.sqrt(np.outer(dg, dg)))
                    assert np.all(np.abs(np.array(g.hessian, np.float64) - upper21(hr["H"].astype(np.float64))) <= REL * sc + 1e-30)
                # the decision the reference's rule takes on the GPU's own sums
                with np.errstate(all="ignore"):
                    err = np.sqrt(np.float64(g.err2) / np.float64(g.n_valid))
                if err_prev is None:
                    assert g.accepted == 1                                         # RPI.h:4599-4605
                else:
                    assert bool(g.accepted) == bool(err_prev - err > P.tol_residual)   # RPI.h:4711-4715
                if g.accepted:
                    err_prev = err
        res_o = orc.align(src, trg, case["guess"], P)
        assert res_g["status"] == res_o.status
        if list(res_g["iters"][:L]) == list(res_o.iters)[:L] and res_o.status == 0:
            ang, dist = pose_err(np.array(res_g["pose"], np.float32).reshape(4, 4).T, orc.pose_from(res_o.pose))
            assert ang <= 1e-4 and dist <= 1e-4, (ang, dist)
    finally:
        ctx.close()
