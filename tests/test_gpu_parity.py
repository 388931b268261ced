"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on the same inputs.

Bar (BASELINE.json north_star): integer work (index maps, masks, counts) bit-exact; float
pyramid planes bit-exact (same IEEE op sequence); per-iteration residual sums within 1e-4
relative; final poses within 1e-4 rad / 1e-4 m.
"""
import numpy as np
import pytest
from util import pose_err, small_pose, upper21

pytestmark = pytest.mark.gpu

REL = 1e-4          # tolerance on sums (north_star)
POSE_RAD = 1e-4
POSE_M = 1e-4


@pytest.fixture(scope="module")
def pair_small(orc, r360):
    """Synthetic 512x256 pair, 3 levels: oracle frames + GPU context."""
    rows, cols, L = 256, 512, 3
    P = orc.default_params(n_levels=L)
    rgb_t, d_t = orc.synth_frame(0, 0, rows, cols)
    rgb_s, d_s = orc.synth_frame(0, 1, rows, cols)
    trg = orc.Frame(rgb_t, d_t, P, True)
    src = orc.Frame(rgb_s, d_s, P, False)
    gp = r360.default_params(n_levels=L)
    ctx = r360.Context(rows, cols, 4, 2, gp)
    ctx.set_frames(0, np.stack([rgb_s, rgb_t]), np.stack([d_s, d_t]), [r360.ROLE_SOURCE, r360.ROLE_TARGET])
    yield dict(orc=orc, ctx=ctx, src=src, trg=trg, P=P, L=L, rows=rows, cols=cols,
               rgb=(rgb_s, rgb_t), depth=(d_s, d_t))
    ctx.close()


def test_synth_frames_bit_exact(pair_small):
    """GPU-rendered synthetic frames == host-rendered (same pinned math)."""
    ctx = pair_small["ctx"]
    rgb, d = ctx.synth_frames(0, 0, 2)
    assert np.array_equal(rgb[1], pair_small["rgb"][0])
    assert np.array_equal(rgb[0], pair_small["rgb"][1])
    assert np.array_equal(d[1], pair_small["depth"][0])
    assert np.array_equal(d[0], pair_small["depth"][1])


def test_pyramids_bit_exact(pair_small):
    """a1-a5: gray/depth pyramids, gradient planes and joint mask, every level."""
    ctx, trg, src = pair_small["ctx"], pair_small["trg"], pair_small["src"]
    for level in range(pair_small["L"]):
        g = ctx.dump_level(1, level)
        o = trg.level(level)
        for k in ("gray", "depth", "ggx", "ggy", "dgx", "dgy"):
            assert np.array_equal(g[k].view(np.int32), o[k].view(np.int32)), (level, k)
        gs = ctx.dump_source_level(0, level)
        os_ = src.level(level)
        for k in ("gray", "depth"):
            assert np.array_equal(gs[k].view(np.int32), os_[k].view(np.int32)), (level, k)


POSES = [small_pose(), small_pose(0.01, -0.02, 0.015, 0.03, -0.02, 0.05),
         small_pose(0.3, 0.1, -0.2, 0.2, 0.1, -0.3), small_pose(-3.0, 0.0, 0.0, 0.0, 0.0, 0.0)]


@pytest.mark.parametrize("pi", range(len(POSES)))
def test_warp_maps_bit_exact(pair_small, pi):
    """Index maps (r', c') and validPixelsPhoto/Depth masks, every level."""
    orc, ctx = pair_small["orc"], pair_small["ctx"]
    for level in range(pair_small["L"]):
        ro, co, vpo, vdo = orc.warp(pair_small["src"], pair_small["trg"], level, POSES[pi], pair_small["P"])
        rg, cg, vpg, vdg = ctx.dump_warp(0, 1, level, POSES[pi])
        assert np.array_equal(ro, rg), (level, int(np.count_nonzero(ro != rg)))
        assert np.array_equal(co, cg), (level, int(np.count_nonzero(co != cg)))
        assert np.array_equal(vpo, vpg)
        assert np.array_equal(vdo, vdg)


def test_packed_index_path(pair_small):
    """The packed (2 pixels per thread) pinned index sequence never disagrees with the scalar one,
    and sends only a small share of pixels (range guards, exact .5 ties) to the scalar code."""
    ctx = pair_small["ctx"]
    rng = np.random.default_rng(7)
    poses = list(POSES) + [small_pose(*(rng.uniform(-1, 1, 3) * 0.5), *(rng.uniform(-1, 1, 3) * 1.5)) for _ in range(6)]
    for T in poses:
        for level in range(pair_small["L"]):
            st = ctx.index_stats(0, 1, level, T)
            assert st["valid"] > 0
            assert st["mismatch"] == 0, st
            assert st["scalar"] < 0.01 * st["valid"] + 8, st


@pytest.mark.parametrize("pi", range(len(POSES)))
def test_error_and_hessgrad(pair_small, pi):
    """a8/a9: counts bit-exact, sums within 1e-4 relative."""
    orc, ctx, P = pair_small["orc"], pair_small["ctx"], pair_small["P"]
    for level in range(pair_small["L"]):
        e2o, nvo = orc.error(pair_small["src"], pair_small["trg"], level, POSES[pi], P)
        e2g, nvg = ctx.eval_error(0, 1, level, POSES[pi])
        assert nvo == nvg
        assert abs(e2g - e2o) <= REL * abs(e2o)
        ho = orc.hessgrad(pair_small["src"], pair_small["trg"], level, POSES[pi], P)
        Hg, gg, nvis = ctx.eval_hessgrad(0, 1, level, POSES[pi])
        assert nvis == ho["n_visible"]
        Ho = ho["H"].astype(np.float64)
        scale = np.sqrt(np.outer(np.diag(Ho), np.diag(Ho)))
        assert np.all(np.abs(Hg - Ho) <= REL * scale), np.max(np.abs(Hg - Ho) / scale)
        # g_a is a sum of J_a * r terms: bound by sqrt(H_aa * err2) (Cauchy-Schwarz)
        gscale = np.sqrt(np.diag(Ho) * e2o)
        assert np.all(np.abs(gg - ho["g"]) <= REL * gscale), np.max(np.abs(gg - ho["g"]) / gscale)


def _check_align(orc, res_g, tr_g, res_o, tr_o, P, src, trg):
    """Control flow and final pose against the oracle's own run; every pose the GPU evaluated is
    replayed through the oracle at the SAME bits, where counts must be exact and sums within 1e-4.
    (Along two independent runs the poses may differ in the last float bit after a solve, which
    legitimately moves a handful of pixels across a rounding boundary.)"""
    L = P.n_levels
    assert list(res_g["iters"][:L]) == list(res_o.iters)[:L]
    assert res_g["status"] == res_o.status
    per = P.max_iters + 2
    for lvl in range(L):
        for k in range(per):
            o, g = tr_o[lvl * per + k], tr_g[lvl * per + k]
            assert bool(o.used) == bool(g.used), (lvl, k)
            if not o.used:
                continue
            assert o.accepted == g.accepted and o.it == g.it
            pose_g = np.array(g.pose, np.float32).reshape(4, 4).T
            e2r, nvr = orc.error(src, trg, lvl, pose_g, P)                        # replay, same bits
            assert nvr == g.n_valid, (lvl, k)                                    # integer: bit-exact
            assert abs(g.err2 - e2r) <= REL * abs(e2r), (lvl, k)
            hr = orc.hessgrad(src, trg, lvl, pose_g, P)
            assert hr["n_visible"] == g.n_visible
            if not (g.used & 2):
                # error-only pass: this candidate was predicted (Gauss-Newton model) to end the level, so the
                # kernel did not form its normal equations -- exactly as the reference, which calls
                # calcHessGrad_sphere at a pose only when the loop goes on
                # (such a record is always the last of its level: had the loop gone on, the fused pass at the
                # same pose would have filled it in)
                assert k == max(kk for kk in range(per) if tr_g[lvl * per + kk].used), (lvl, k)
                continue
            Ho = upper21(hr["H"].astype(np.float64)); Hg = np.array(g.hessian, np.float64)
            dg = np.diag(hr["H"]).astype(np.float64)
            sc = upper21(np.sqrt(np.outer(dg, dg)))
            assert np.all(np.abs(Hg - Ho) <= REL * sc), (lvl, k)
            gs = np.sqrt(dg * e2r)
            assert np.all(np.abs(np.array(g.gradient) - hr["g"]) <= REL * gs), (lvl, k)
    To = orc.pose_from(res_o.pose)
    Tg = np.array(res_g["pose"], np.float32).reshape(4, 4).T
    ang, dist = pose_err(Tg, To)
    assert ang <= POSE_RAD and dist <= POSE_M, (ang, dist)
    # the final record replayed at the GPU's own final pose: count exact, sum 1e-4 (the getters' Hessian / SSO belong
    # to the last calcHessGrad_sphere call, replayed above with the trace)
    e2r, nvr = orc.error(src, trg, 0, Tg, P) if res_g["status"] == 0 else (res_g["final_err2"], res_g["final_n_valid"])
    assert nvr == res_g["final_n_valid"]
    assert abs(res_g["final_err2"] - e2r) <= REL * abs(e2r)


def test_align_identity_guess(pair_small):
    """alignFrames360 end to end with per-iteration trace: same control flow, sums, pose."""
    orc, ctx, P = pair_small["orc"], pair_small["ctx"], pair_small["P"]
    res_o, tr_o = orc.align(pair_small["src"], pair_small["trg"], None, P, trace=True)
    res_g, tr_g = ctx.register_pairs([0], [1], None, trace=True)
    _check_align(orc, res_g[0], tr_g, res_o, tr_o, P, pair_small["src"], pair_small["trg"])
    gt = orc.synth_gt_pose(0, 1, 0)
    ang, dist = pose_err(np.array(res_g[0]["pose"]).reshape(4, 4).T, gt)
    assert ang < 2e-3 and dist < 5e-3          # converged to the analytic ground truth


def test_align_with_guess_and_batch(pair_small, r360):
    """A batch of pairs with different initial guesses == one-by-one oracle runs."""
    orc, ctx, P = pair_small["orc"], pair_small["ctx"], pair_small["P"]
    guesses = [small_pose(0.01, 0.0, -0.01, 0.02, 0.0, -0.02), small_pose()]
    gp = np.stack([r360.pose_to_colmajor(g) for g in guesses])
    res_g, tr_g = ctx.register_pairs([0, 0], [1, 1], gp, trace=True)
    per = P.n_levels * (P.max_iters + 2)
    for k, g in enumerate(guesses):
        res_o, tr_o = orc.align(pair_small["src"], pair_small["trg"], g, P, trace=True)
        _check_align(orc, res_g[k], tr_g[k * per:(k + 1) * per], res_o, tr_o, P, pair_small["src"], pair_small["trg"])


@pytest.mark.parametrize("method", [0, 1])
def test_photo_only_and_depth_only(orc, r360, method):
    rows, cols, L = 128, 256, 2
    P = orc.default_params(n_levels=L, method=method)
    rgb_t, d_t = orc.synth_frame(0, 4, rows, cols)
    rgb_s, d_s = orc.synth_frame(0, 5, rows, cols)
    trg = orc.Frame(rgb_t, d_t, P, True); src = orc.Frame(rgb_s, d_s, P, False)
    ctx = r360.Context(rows, cols, 2, 1, r360.default_params(n_levels=L, method=method))
    ctx.set_frames(0, np.stack([rgb_s, rgb_t]), np.stack([d_s, d_t]))
    res_o, tr_o = orc.align(src, trg, None, P, trace=True)
    res_g, tr_g = ctx.register_pairs([0], [1], None, trace=True)
    _check_align(orc, res_g[0], tr_g, res_o, tr_o, P, src, trg)
    ctx.close()


def test_invalid_depth_and_float_depth(orc, r360):
    """Zero / out-of-range depth (INVALID_POINT path, zero-depth target texels) and CV_32F depth."""
    rows, cols, L = 128, 256, 3
    P = orc.default_params(n_levels=L)
    rng = np.random.default_rng(3)
    fr = []
    for fid in (8, 9):
        rgb, d = orc.synth_frame(0, fid, rows, cols)
        d = d.copy()
        d[rng.random(d.shape) < 0.15] = 0                     # holes
        d[10:20, 30:90] = 7000                                # beyond maxDepth
        d[40:44, :] = 200                                     # below minDepth
        fr.append((rgb, d))
    trg = orc.Frame(fr[0][0], fr[0][1], P, True); src = orc.Frame(fr[1][0], fr[1][1], P, False)
    ctx = r360.Context(rows, cols, 4, 1, r360.default_params(n_levels=L))
    ctx.set_frames(0, np.stack([fr[1][0], fr[0][0]]), np.stack([fr[1][1], fr[0][1]]))
    for level in range(L):
        g = ctx.dump_level(1, level); o = trg.level(level)
        for k in o:
            assert np.array_equal(g[k].view(np.int32), o[k].view(np.int32)), (level, k)
        ro, co, vpo, vdo = orc.warp(src, trg, level, POSES[1], P)
        rg, cg, vpg, vdg = ctx.dump_warp(0, 1, level, POSES[1])
        assert np.array_equal(ro, rg) and np.array_equal(co, cg)
        assert np.array_equal(vpo, vpg) and np.array_equal(vdo, vdg)
    res_o, tr_o = orc.align(src, trg, None, P, trace=True)
    res_g, tr_g = ctx.register_pairs([0], [1], None, trace=True)
    _check_align(orc, res_g[0], tr_g, res_o, tr_o, P, src, trg)
    # float depth in metres: same planes as the u16 path after * 0.001f
    dm = [(f[1].astype(np.float32) * np.float32(0.001)) for f in fr]
    ctx.set_frames(2, np.stack([fr[1][0], fr[0][0]]), np.stack([dm[1], dm[0]]))
    for level in range(L):
        a = ctx.dump_level(1, level); b = ctx.dump_level(3, level)
        for k in a:
            assert np.array_equal(a[k].view(np.int32), b[k].view(np.int32))
    ctx.close()


def test_non_power_of_two_width(orc, r360):
    """1920x320-like geometry (the sample pair's shape, scaled): generic row/col split."""
    rows, cols, L = 80, 480, 3
    P = orc.default_params(n_levels=L)
    rgb_t, d_t = orc.synth_frame(0, 2, rows, cols)
    rgb_s, d_s = orc.synth_frame(0, 3, rows, cols)
    trg = orc.Frame(rgb_t, d_t, P, True); src = orc.Frame(rgb_s, d_s, P, False)
    ctx = r360.Context(rows, cols, 2, 1, r360.default_params(n_levels=L))
    ctx.set_frames(0, np.stack([rgb_s, rgb_t]), np.stack([d_s, d_t]))
    for level in range(L):
        ro, co, vpo, vdo = orc.warp(src, trg, level, POSES[1], P)
        rg, cg, vpg, vdg = ctx.dump_warp(0, 1, level, POSES[1])
        assert np.array_equal(ro, rg) and np.array_equal(co, cg)
        assert np.array_equal(vpo, vpg) and np.array_equal(vdo, vdg)
    res_o, tr_o = orc.align(src, trg, None, P, trace=True)
    res_g, tr_g = ctx.register_pairs([0], [1], None, trace=True)
    _check_align(orc, res_g[0], tr_g, res_o, tr_o, P, src, trg)
    ctx.close()


@pytest.mark.parametrize("shape", [(80, 480, 3), (72, 200, 3), (256, 512, 4), (16, 32, 2)])
def test_pyramid_routes_agree_bit_for_bit(orc, r360, shape, monkeypatch):
    """The fused pyramid passes (k_pyr_head over the raw input, the same pass over the plane of every level >= 1) against the
    separate kernels (k_level0 / k_down / k_texel) and against the oracle: every plane of every level, both roles, partial
    tiles, widths that are not a multiple of the tile, 16-bit and CV_32F depth."""
    rows, cols, L = shape
    P = orc.default_params(n_levels=L)
    rng = np.random.default_rng(rows * 1000 + cols)
    rgb = rng.integers(0, 256, size=(2, rows, cols, 3), dtype=np.uint8)
    d = rng.integers(0, 12000, size=(2, rows, cols)).astype(np.uint16)
    d[rng.random(d.shape) < 0.1] = 0                                      # invalid depth
    ref = [orc.Frame(rgb[k], d[k], P, True) for k in range(2)]
    dumps = {}
    for mid, head in (("1", "1"), ("0", "1"), ("0", "0")):
        monkeypatch.setenv("R360_PYR_MID", mid)
        monkeypatch.setenv("R360_PYR_HEAD", head)
        for f32 in (False, True):
            ctx = r360.Context(rows, cols, 2, 1, r360.default_params(n_levels=L))
            ctx.set_frames(0, rgb, (d.astype(np.float32) * np.float32(0.001)) if f32 else d)
            out = []
            for k in range(2):
                for level in range(L):
                    g = ctx.dump_level(k, level)
                    gs = ctx.dump_source_level(k, level)
                    out.append(np.concatenate([g[n].view(np.int32).ravel() for n in ("gray", "depth", "ggx", "ggy", "dgx", "dgy")] +
                                              [gs[n].view(np.int32).ravel() for n in ("gray", "depth")]))
                    if not f32:
                        o = ref[k].level(level)
                        for n in ("gray", "depth", "ggx", "ggy", "dgx", "dgy"):
                            assert np.array_equal(g[n].view(np.int32), o[n].view(np.int32)), (mid, head, k, level, n)
            dumps[(mid, head, f32)] = np.concatenate(out)
            ctx.close()
    for f32 in (False, True):
        assert np.array_equal(dumps[("1", "1", f32)], dumps[("0", "1", f32)])
        assert np.array_equal(dumps[("1", "1", f32)], dumps[("0", "0", f32)])
    # the same depth values as CV_32F: the monotonicity test on 2^100-scaled differences == the one on the differences themselves
    assert np.array_equal(dumps[("1", "1", False)], dumps[("1", "1", True)])


def test_register_host_pairs_streaming(r360):
    """The pipelined host entry point (upload + pyramids + batched registration overlapped) returns
    the same records as set_frames + register_pairs; more pairs than one internal batch."""
    rows, cols, L, n = 64, 128, 3, 70
    ctx = r360.Context(rows, cols, 2 * n, n, r360.default_params(n_levels=L))
    rgb, dep = ctx.synth_frames(0, 0, 2 * n)
    roles = np.array([r360.ROLE_TARGET, r360.ROLE_SOURCE] * n, np.uint8)
    ctx.set_frames(0, rgb, dep, roles)
    trg_idx = np.arange(0, 2 * n, 2); src_idx = trg_idx + 1
    rng = np.random.default_rng(5)
    guesses = np.stack([r360.pose_to_colmajor(small_pose(*(rng.uniform(-1, 1, 6) * 0.01))) for _ in range(n)])
    ref = ctx.register_pairs(src_idx, trg_idx, guesses)
    ctx2 = r360.Context(rows, cols, 2 * n, n, r360.default_params(n_levels=L))
    got = ctx2.register_host_pairs(rgb, dep, n, guesses)
    for k in range(n):
        assert got[k]["pair_id"] == k and got[k]["status"] == ref[k]["status"]
        assert list(got[k]["iters"]) == list(ref[k]["iters"])
        # (two batch compositions -- 64 + 6 pairs here, 70 there -- cut the pixels of a pair into other per-thread
        # float partial sums: same decisions, last bits of the sums may differ; a repeated call is bit-identical,
        # test_deterministic_run_to_run)
        assert np.allclose(got[k]["pose"], ref[k]["pose"], atol=1e-6)
        assert got[k]["final_n_valid"] == ref[k]["final_n_valid"]
    # frames stay resident in slots 2p / 2p+1: a second registration through the slot API agrees
    again = ctx2.register_pairs(src_idx[:3], trg_idx[:3], guesses[:3])
    assert np.allclose(again["pose"], ref["pose"][:3], atol=1e-6)
    ctx.close(); ctx2.close()


def test_deterministic_run_to_run(r360):
    """The same call twice gives the same BYTES in every field of every record and of every trace entry: the per-pair sums
    are accumulated across CTAs in fixed point (integer atomics: order-independent), everything else is a fixed
    function of the inputs.  Also across a fresh context, through the host-pairs entry point and the evaluation hooks."""
    rows, cols, L, n = 128, 256, 3, 12
    ctx = r360.Context(rows, cols, 2 * n, n, r360.default_params(n_levels=L))
    rgb, dep = ctx.synth_frames(0, 0, 2 * n)
    roles = np.array([r360.ROLE_TARGET, r360.ROLE_SOURCE] * n, np.uint8)
    ctx.set_frames(0, rgb, dep, roles)
    trg_idx = np.arange(0, 2 * n, 2); src_idx = trg_idx + 1
    rng = np.random.default_rng(11)
    guesses = np.stack([r360.pose_to_colmajor(small_pose(*(rng.uniform(-1, 1, 6) * 0.02))) for _ in range(n)])
    a, tr_a = ctx.register_pairs(src_idx, trg_idx, guesses, trace=True)
    for _ in range(3):
        b, tr_b = ctx.register_pairs(src_idx, trg_idx, guesses, trace=True)
        assert a.tobytes() == b.tobytes()
        assert bytes(tr_a) == bytes(tr_b)
    ctx2 = r360.Context(rows, cols, 2 * n, n, r360.default_params(n_levels=L))
    ctx2.set_frames(0, rgb, dep, roles)
    assert ctx2.register_pairs(src_idx, trg_idx, guesses).tobytes() == a.tobytes()
    h1 = ctx2.register_host_pairs(rgb, dep, n, guesses)
    h2 = ctx2.register_host_pairs(rgb, dep, n, guesses)
    assert h1.tobytes() == h2.tobytes() == a.tobytes()          # n <= one internal batch: the same decomposition
    for level in range(L):
        H1, g1, n1 = ctx.eval_hessgrad(1, 0, level, POSES[1]); H2, g2, n2 = ctx.eval_hessgrad(1, 0, level, POSES[1])
        assert H1.tobytes() == H2.tobytes() and g1.tobytes() == g2.tobytes() and n1 == n2
        assert ctx.eval_error(1, 0, level, POSES[1]) == ctx.eval_error(1, 0, level, POSES[1])
    ctx.close(); ctx2.close()


def test_slot_reset_with_narrower_role_invalidates_the_other_pyramid(r360):
    """A slot set as BOTH and later re-set as TARGET only no longer has a source pyramid (the allocation stays, its
    content is another frame's): registering it as a source is a state error, not a silent use of the old frame."""
    ctx = r360.Context(64, 128, 2, 1, r360.default_params(n_levels=2))
    rgb, dep = ctx.synth_frames(0, 0, 3)
    ctx.set_frames(0, rgb[:2], dep[:2])                                      # both slots, both roles
    assert ctx.register_pairs([0], [1])[0]["status"] == 0
    ctx.set_frames(0, rgb[2:3], dep[2:3], [r360.ROLE_TARGET])                # slot 0: another frame, target only
    with pytest.raises(r360.R360Error):
        ctx.register_pairs([0], [1])
    with pytest.raises(r360.R360Error):
        ctx.dump_source_level(0, 0)
    assert ctx.register_pairs([1], [0])[0]["status"] == 0                    # its target pyramid is the new frame's
    ctx.register_host_pairs(rgb[:2], dep[:2], 1)                             # slot 0 target, slot 1 source
    with pytest.raises(r360.R360Error):
        ctx.register_pairs([0], [1])                                         # slot 0 has no source, slot 1 no target
    ctx.close()


def test_python_shape_checks(r360):
    """The ctypes wrappers pass raw pointers: mismatched arrays must raise in Python, not read out of bounds in C."""
    ctx = r360.Context(64, 128, 4, 2, r360.default_params(n_levels=2))
    rgb, dep = ctx.synth_frames(0, 0, 2)
    with pytest.raises(ValueError):
        ctx.set_frames(0, rgb[:, :32], dep)
    with pytest.raises(ValueError):
        ctx.set_frames(0, rgb, dep[:1])
    with pytest.raises(ValueError):
        ctx.set_frames(3, rgb, dep)
    with pytest.raises(ValueError):
        ctx.set_frames(0, rgb, dep, [r360.ROLE_BOTH])
    ctx.set_frames(0, rgb, dep)
    with pytest.raises(ValueError):
        ctx.register_pairs([0, 1], [1])
    with pytest.raises(ValueError):
        ctx.register_pairs([0], [1], np.zeros(15, np.float32))
    with pytest.raises(ValueError):
        ctx.register_pairs([0], [1], out=np.zeros(1, np.float32))
    with pytest.raises(ValueError):
        ctx.register_host_pairs(rgb, dep, 2)
    ctx.close()


def test_errors_are_loud(r360):
    ctx = r360.Context(64, 128, 2, 1, r360.default_params(n_levels=2))
    with pytest.raises(r360.R360Error):
        ctx.register_pairs([0], [1])                  # frames never set
    with pytest.raises(r360.R360Error):
        r360.Context(64, 128, 2, 1, r360.default_params(n_levels=2, occlusion=3))
    with pytest.raises(r360.R360Error):
        r360.Context(65, 128, 2, 1, r360.default_params(n_levels=2))
    ctx.close()


def test_cpp_class_drop_in(tmp_path):
    """The C++ mirror of the reference class (include/RegisterPhotoICP_b200.hpp) gives the same pose
    as the batched C ABI call on the same frames."""
    import os, subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = tmp_path / "class_demo"
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(root, "include"),
                           os.path.join(root, "tests", "cpp", "class_demo.cpp"), "-o", str(exe),
                           "-L", os.path.join(root, "rgbd360_b200"), "-lrgbd360_b200",
                           "-Wl,-rpath," + os.path.join(root, "rgbd360_b200")])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0 and "OK" in out.stdout, out.stdout + out.stderr


def test_sample_pair_config1(orc, r360):
    """Config #1: the reference's sample pair (1920x320, 4 levels, PHOTO_DEPTH, guess Identity):
    GPU vs the oracle run on the same stitched frames and vs the committed golden vector."""
    import json, os
    gold_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    d = np.load(os.path.join(gold_dir, "sample_pair.npz"))
    gold = json.load(open(os.path.join(gold_dir, "sample_pair_oracle.json")))["pinned"]
    P = orc.default_params(n_levels=4)
    trg = orc.Frame(d["trg_rgb"], d["trg_depth"], P, True)
    src = orc.Frame(d["src_rgb"], d["src_depth"], P, False)
    ctx = r360.Context(320, 1920, 2, 1, r360.default_params(n_levels=4))
    ctx.set_frames(0, np.stack([d["src_rgb"], d["trg_rgb"]]), np.stack([d["src_depth"], d["trg_depth"]]),
                   [r360.ROLE_SOURCE, r360.ROLE_TARGET])
    for level in range(4):                                   # real data: holes, sensor joints, borders
        g = ctx.dump_level(1, level); o = trg.level(level)
        for k in o:
            assert np.array_equal(g[k].view(np.int32), o[k].view(np.int32)), (level, k)
        ro, co, vpo, vdo = orc.warp(src, trg, level, POSES[1], P)
        rg, cg, vpg, vdg = ctx.dump_warp(0, 1, level, POSES[1])
        assert np.array_equal(ro, rg) and np.array_equal(co, cg)
        assert np.array_equal(vpo, vpg) and np.array_equal(vdo, vdg)
    res_o, tr_o = orc.align(src, trg, None, P, trace=True)
    res_g, tr_g = ctx.register_pairs([0], [1], None, trace=True)
    _check_align(orc, res_g[0], tr_g, res_o, tr_o, P, src, trg)
    assert list(res_g[0]["iters"][:4]) == gold["iters"]
    Tg = np.array(res_g[0]["pose"], np.float32).reshape(4, 4).T
    ang, dist = pose_err(Tg, np.array(gold["pose"], np.float32).reshape(4, 4).T)
    assert ang <= POSE_RAD and dist <= POSE_M
    ctx.close()


def test_full_size_properties(orc, r360):
    """BASELINE.json's full size (2048x1024, 4 levels): size-independent properties on a small
    batch -- convergence to the analytic ground truth, oracle replay of the final pose (counts
    exact, sum 1e-4), and batch independence (a pair's result does not depend on its neighbours)."""
    rows, cols, L, n = 1024, 2048, 4, 3
    ctx = r360.Context(rows, cols, 2 * n, n, r360.default_params(n_levels=L))
    rgb, dep = ctx.synth_frames(0, 0, 2 * n)
    roles = np.array([r360.ROLE_TARGET, r360.ROLE_SOURCE] * n, np.uint8)
    ctx.set_frames(0, rgb, dep, roles)
    trg_idx = np.arange(0, 2 * n, 2); src_idx = trg_idx + 1
    res = ctx.register_pairs(src_idx, trg_idx)
    P = orc.default_params(n_levels=L)
    for k in range(n):
        assert res[k]["status"] == 0 and res[k]["pair_id"] == k
        T = np.array(res[k]["pose"], np.float32).reshape(4, 4).T
        ang, dist = pose_err(T, orc.synth_gt_pose(0, 2 * k + 1, 2 * k))
        assert ang < 5e-4 and dist < 1.5e-3, (k, ang, dist)
        single = ctx.register_pairs([src_idx[k]], [trg_idx[k]])[0]
        assert np.allclose(single["pose"], res[k]["pose"], atol=1e-6)
        assert list(single["iters"]) == list(res[k]["iters"])
    # packed index path at full size: identical to the scalar pinned sequence on every pixel it keeps
    for level in range(L):
        for T in (POSES[0], POSES[1], POSES[2], np.array(res[0]["pose"], np.float32).reshape(4, 4).T):
            st = ctx.index_stats(1, 0, level, T)
            assert st["mismatch"] == 0 and st["scalar"] < 0.01 * st["valid"], (level, st)
    k = 1
    trg = orc.Frame(rgb[2 * k], dep[2 * k], P, True); src = orc.Frame(rgb[2 * k + 1], dep[2 * k + 1], P, False)
    T = np.array(res[k]["pose"], np.float32).reshape(4, 4).T
    e2, nv = orc.error(src, trg, 0, T, P)
    assert nv == res[k]["final_n_valid"]
    assert abs(e2 - res[k]["final_err2"]) <= REL * e2
    res_o = orc.align(src, trg, None, P)
    ang, dist = pose_err(T, orc.pose_from(res_o.pose))
    assert ang <= POSE_RAD and dist <= POSE_M, (ang, dist)
    assert list(res_o.iters)[:L] == list(res[k]["iters"][:L])
    ctx.close()


def test_edge_cases_empty_ragged_and_limits(orc, r360):
    """Empty batch, frames of one role only, geometry the pyramid cannot halve, out-of-range indices,
    and the smallest legal image (every level still has an even width)."""
    ctx = r360.Context(64, 128, 3, 2, r360.default_params(n_levels=3))
    rgb, dep = ctx.synth_frames(0, 0, 3)
    assert len(ctx.register_pairs([], [])) == 0                                   # empty batch: no work, no error
    ctx.set_frames(0, rgb[:0], dep[:0])                                          # empty frame range
    ctx.set_frames(0, rgb, dep, [r360.ROLE_SOURCE, r360.ROLE_TARGET, r360.ROLE_BOTH])
    with pytest.raises(r360.R360Error):
        ctx.register_pairs([1], [0])                                             # frame 1 has no source pyramid / 0 no target
    with pytest.raises(r360.R360Error):
        ctx.register_pairs([0], [3])                                             # slot out of range
    with pytest.raises((r360.R360Error, ValueError)):
        ctx.register_pairs([0, 2, 0], [1, 1, 2])                                 # more pairs than max_pairs
    s3, t3 = np.array([0, 2, 0], np.int32), np.array([1, 1, 2], np.int32)        # ... and straight at the C ABI
    out3 = np.zeros(3, r360.native.RESULT_DTYPE)
    assert ctx.L.r360_register_pairs(ctx.h, 3, s3.ctypes.data, t3.ctypes.data, None, out3.ctypes.data, None) == -1
    with pytest.raises(r360.R360Error):
        ctx.eval_error(0, 1, 3, np.eye(4))                                       # level out of range
    ok = ctx.register_pairs([0, 2], [1, 2])                                      # ragged roles; a frame against itself
    assert ok[1]["status"] == 0 and np.allclose(np.array(ok[1]["pose"]).reshape(4, 4), np.eye(4), atol=1e-5)
    ctx.close()
    for rows, cols, L in ((64, 130, 3), (66, 128, 3), (64, 128, 9)):             # cols % 2^L, rows % 2^(L-1), too many levels
        with pytest.raises(r360.R360Error):
            r360.Context(rows, cols, 2, 1, r360.default_params(n_levels=L))
    # smallest image with 4 levels: 8 x 16 -> coarsest level 1 x 2
    P = orc.default_params(n_levels=4)
    small = r360.Context(8, 16, 2, 1, r360.default_params(n_levels=4))
    rgb, dep = small.synth_frames(0, 0, 2)
    small.set_frames(0, rgb, dep)
    trg = orc.Frame(rgb[0], dep[0], P, True); src = orc.Frame(rgb[1], dep[1], P, False)
    for level in range(4):
        g = small.dump_level(0, level); o = trg.level(level)
        for k in o:
            assert np.array_equal(g[k].view(np.int32), o[k].view(np.int32)), (level, k)
        ro, co, vpo, vdo = orc.warp(src, trg, level, POSES[1], P)
        rg, cg, vpg, vdg = small.dump_warp(1, 0, level, POSES[1])
        assert np.array_equal(ro, rg) and np.array_equal(co, cg) and np.array_equal(vpo, vpg) and np.array_equal(vdo, vdg)
    res_o = orc.align(src, trg, None, P)
    res_g = small.register_pairs([1], [0])[0]
    assert list(res_g["iters"][:4]) == list(res_o.iters)[:4] and res_g["status"] == res_o.status
    small.close()


def test_large_image_4096x2048(orc, r360):
    """Above BASELINE's size: 4096 x 2048, 5 levels (8.4 Mpixel level 0).  Packed index path == scalar on
    every pixel, final pose replayed through the oracle (counts exact, sum 1e-4), converges to the
    analytic ground truth."""
    rows, cols, L = 2048, 4096, 5
    ctx = r360.Context(rows, cols, 2, 1, r360.default_params(n_levels=L))
    rgb, dep = ctx.synth_frames(0, 10, 2)
    ctx.set_frames(0, rgb, dep, [r360.ROLE_TARGET, r360.ROLE_SOURCE])
    res = ctx.register_pairs([1], [0])[0]
    assert res["status"] == 0
    T = np.array(res["pose"], np.float32).reshape(4, 4).T
    ang, dist = pose_err(T, orc.synth_gt_pose(0, 11, 10))
    assert ang < 5e-4 and dist < 1.5e-3, (ang, dist)
    st = ctx.index_stats(1, 0, 0, T)
    assert st["mismatch"] == 0 and st["valid"] == rows * cols
    P = orc.default_params(n_levels=L)
    trg = orc.Frame(rgb[0], dep[0], P, True); src = orc.Frame(rgb[1], dep[1], P, False)
    e2, nv = orc.error(src, trg, 0, T, P)
    assert nv == res["final_n_valid"]
    assert abs(e2 - res["final_err2"]) <= REL * e2
    ctx.close()
