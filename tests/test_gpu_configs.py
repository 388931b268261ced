"""BASELINE.json configs 4 and 5 at test scale, through the same sharding code bench/production use:
  config 4  sequence odometry -- consecutive frames, frame k is the target of pair k AND the source of
            pair k-1 (one resident pyramid set serves both roles), contiguous shards with a halo frame;
  config 5  loop closure -- all C(n,2) pairs over replicated keyframes, dealt round-robin, initial
            guess = ground truth o exp(delta).
Every pair is checked against the CPU oracle (same iteration counts, pose within 1e-4 rad / 1e-4 m);
the union of the per-rank results equals the single-rank run (sharding changes nothing)."""
import numpy as np
import pytest
from util import pose_err

pytestmark = pytest.mark.gpu
POSE_RAD, POSE_M = 1e-4, 1e-4


def _oracle_pair(orc, P, rgb, dep, s, t, guess):
    trg = orc.Frame(rgb[t], dep[t], P, True)
    src = orc.Frame(rgb[s], dep[s], P, False)
    return orc.align(src, trg, guess, P)


def test_config4_odometry_shards(orc, r360):
    from rgbd360_b200 import shard
    rows, cols, L, n_frames = 128, 256, 3, 9
    gp = r360.default_params(n_levels=L, std_photo=np.float32(3.0 / 255))      # odometry callers use 3/255
    P = orc.default_params(n_levels=L, std_photo=3.0 / 255)
    full = r360.Context(rows, cols, n_frames, n_frames - 1, gp)
    rgb, dep = full.synth_frames(0, 40, n_frames)
    # single rank: all 8 pairs, every inner frame resident once with both roles
    pairs, frames, s, t = shard.odometry_pairs(n_frames, 0, 1)
    roles = shard.frame_roles(s, t, len(frames))
    assert roles[0] == r360.ROLE_TARGET and roles[-1] == r360.ROLE_SOURCE and np.all(roles[1:-1] == 3)
    full.set_frames(0, rgb[frames], dep[frames], roles)
    ref = full.register_pairs(s, t)
    for k in range(len(pairs)):
        o = _oracle_pair(orc, P, rgb, dep, frames[s[k]], frames[t[k]], None)
        assert list(ref[k]["iters"][:L]) == list(o.iters)[:L], k
        ang, dist = pose_err(np.array(ref[k]["pose"]).reshape(4, 4).T, orc.pose_from(o.pose))
        assert ang <= POSE_RAD and dist <= POSE_M, (k, ang, dist)
        gt = orc.synth_gt_pose(0, 40 + k + 1, 40 + k)
        ang, dist = pose_err(np.array(ref[k]["pose"]).reshape(4, 4).T, gt)
        assert ang < 5e-3 and dist < 1e-2, (k, ang, dist)
    # two ranks (run one after the other here): contiguous ranges + halo frame; same records
    got = np.zeros(n_frames - 1, r360.native.RESULT_DTYPE)
    for rank in range(2):
        pairs, frames, s, t = shard.odometry_pairs(n_frames, rank, 2)
        ctx = r360.Context(rows, cols, len(frames), len(pairs), gp)
        ctx.set_frames(0, rgb[frames], dep[frames], shard.frame_roles(s, t, len(frames)))
        got[pairs] = ctx.register_pairs(s, t)
        ctx.close()
    assert np.array_equal(got["iters"], ref["iters"])
    assert np.array_equal(got["final_n_valid"], ref["final_n_valid"])
    assert np.allclose(got["pose"], ref["pose"], atol=1e-6)
    full.close()


def test_config5_loop_closure_all_pairs(orc, r360):
    from rgbd360_b200 import shard
    rows, cols, L, n_kf = 128, 256, 3, 6
    gp = r360.default_params(n_levels=L)
    P = orc.default_params(n_levels=L)
    ap = shard.all_pairs(n_kf)                                   # (src, trg), trg = lower index
    assert len(ap) == n_kf * (n_kf - 1) // 2
    ctx = r360.Context(rows, cols, n_kf, len(ap), gp)
    rgb, dep = ctx.synth_frames(1, 0, n_kf)                      # kind 1: loop-closure keyframes
    ctx.set_frames(0, rgb, dep, np.full(n_kf, 3, np.uint8))      # replicated keyframes, both roles
    guesses = np.stack([shard.loop_closure_guess(i, orc.synth_gt_pose(1, int(sv), int(tv))) for i, (sv, tv) in enumerate(ap)])
    gcm = np.stack([r360.pose_to_colmajor(g) for g in guesses])
    ref = ctx.register_pairs(ap[:, 0], ap[:, 1], gcm)
    n_ok = 0
    for i, (sv, tv) in enumerate(ap):
        o = _oracle_pair(orc, P, rgb, dep, int(sv), int(tv), guesses[i])
        assert ref[i]["status"] == o.status, i
        assert list(ref[i]["iters"][:L]) == list(o.iters)[:L], i
        ang, dist = pose_err(np.array(ref[i]["pose"]).reshape(4, 4).T, orc.pose_from(o.pose))
        assert ang <= POSE_RAD and dist <= POSE_M, (i, ang, dist)
        ang, dist = pose_err(np.array(ref[i]["pose"]).reshape(4, 4).T, orc.synth_gt_pose(1, int(sv), int(tv)))
        n_ok += (ang < 5e-3 and dist < 1e-2)
    assert n_ok >= len(ap) - 2                                  # converges to the ground truth (wide baselines)
    # round-robin shards over 4 "ranks": union == the single-rank batch
    got = np.zeros(len(ap), r360.native.RESULT_DTYPE)
    for rank in range(4):
        ids = shard.round_robin(len(ap), rank, 4)
        got[ids] = ctx.register_pairs(ap[ids, 0], ap[ids, 1], gcm[ids])
    assert np.array_equal(got["iters"], ref["iters"])
    assert np.allclose(got["pose"], ref["pose"], atol=1e-6)
    ctx.close()
