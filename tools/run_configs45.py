#!/usr/bin/env python
"""BASELINE.json configs 4 and 5 at their NAMED scale, through the production sharding code
(rgbd360_b200/shard.py), one process per GPU (torchrun), NCCL all-gather of the result records.

  config 4  sequence odometry: 8 193 consecutive synthetic 2048x1024 sphere frames -> 8 192 pairs
            (target = frame k, source = frame k + 1, guess Identity), contiguous shards with one halo
            frame; every inner frame is resident ONCE with both roles.  A shard is streamed through the
            GPU in blocks (frames rendered on the device, pyramids, registration), so it also fits 2 GPUs.
  config 5  loop closure: all C(128, 2) = 8 128 pairs over 128 synthetic 2048x1024 keyframes replicated on
            every rank, pairs dealt round-robin, guess = ground truth o exp(delta).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        tools/run_configs45.py --config 4|5 [--frames 8193] [--keyframes 128] [--levels 4]

Prints ONE JSON line on rank 0: pairs/s (device-timed, max over ranks, synthesis excluded), how many pairs
converged to the analytic ground truth, and the size-independent checks (every pair OK, gathered ids complete).
Not the bench contract (bench.py is): a recorded run of the two remaining BASELINE configurations.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pose_err(T, G):
    T = np.asarray(T, np.float64); G = np.asarray(G, np.float64)
    dR = T[:3, :3] @ G[:3, :3].T
    v = 0.5 * np.array([dR[2, 1] - dR[1, 2], dR[0, 2] - dR[2, 0], dR[1, 0] - dR[0, 1]])
    return float(np.arctan2(np.linalg.norm(v), (np.trace(dR) - 1) / 2)), float(np.linalg.norm(T[:3, 3] - G[:3, 3]))


def main():
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, required=True, choices=[4, 5])
    ap.add_argument("--frames", type=int, default=8193)
    ap.add_argument("--keyframes", type=int, default=128)
    ap.add_argument("--levels", type=int, default=4)
    ap.add_argument("--rows", type=int, default=1024)
    ap.add_argument("--cols", type=int, default=2048)
    ap.add_argument("--block", type=int, default=512, help="pairs per streamed block (config 4) / per call (config 5)")
    args = ap.parse_args()

    import torch
    import torch.distributed as dist
    import rgbd360_b200 as r360
    from rgbd360_b200 import shard

    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rows, cols, L = args.rows, args.cols, args.levels
    dev_ms = 0.0

    if args.config == 4:
        n_pairs_total = args.frames - 1
        pairs, frames, s_loc, t_loc = shard.odometry_pairs(args.frames, rank, world)
        gp = r360.default_params(n_levels=L, std_photo=np.float32(3.0 / 255))           # OdometryRGBD360.cpp:92
        B = min(args.block, max(len(pairs), 1))
        ctx = r360.Context(rows, cols, B + 1, B, gp, device=local)
        rgb_dev = torch.empty((B + 1, rows, cols, 3), dtype=torch.uint8, device="cuda")
        dep_dev = torch.empty((B + 1, rows, cols), dtype=torch.int16, device="cuda")
        res = np.zeros(len(pairs), r360.native.RESULT_DTYPE)
        # warm-up (untimed): first block once -- frame slots are allocated on first use
        nb = min(B, len(pairs))
        ctx.synth_frames_dev(0, int(frames[0]), nb + 1, rgb_dev.data_ptr(), dep_dev.data_ptr())
        ctx.set_frames_ptr(0, nb + 1, rgb_dev.data_ptr(), dep_dev.data_ptr(), None, device=True)
        ctx.register_pairs(np.arange(1, nb + 1, dtype=np.int32), np.arange(nb, dtype=np.int32))
        torch.cuda.synchronize()
        t_wall = time.perf_counter()
        for b0 in range(0, len(pairs), B):
            nb = min(B, len(pairs) - b0)
            f0 = int(frames[0]) + b0                                                   # block frames f0 .. f0 + nb
            ctx.synth_frames_dev(0, f0, nb + 1, rgb_dev.data_ptr(), dep_dev.data_ptr())
            roles = np.full(nb + 1, r360.ROLE_BOTH, np.uint8)
            roles[0] = r360.ROLE_TARGET; roles[nb] = r360.ROLE_SOURCE                   # block-boundary frames: one role here
            ctx.set_frames_ptr(0, nb + 1, rgb_dev.data_ptr(), dep_dev.data_ptr(), roles, device=True)
            dev_ms += ctx.last_device_ms()
            loc = np.arange(nb, dtype=np.int32)
            res[b0:b0 + nb] = ctx.register_pairs(loc + 1, loc)                         # source k + 1 -> target k
            dev_ms += ctx.last_device_ms()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t_wall
        gt = lambda gid: r360.synth_gt_pose(0, int(gid) + 1, int(gid))
        kind_name = "sequence odometry, %d frames" % args.frames
    else:
        n_kf = args.keyframes
        ap_all = shard.all_pairs(n_kf)
        n_pairs_total = len(ap_all)
        ids = shard.round_robin(n_pairs_total, rank, world)
        pairs = ids
        gp = r360.default_params(n_levels=L)
        B = min(args.block, max(len(ids), 1))
        ctx = r360.Context(rows, cols, n_kf, B, gp, device=local)
        chunk = 64
        rgb_dev = torch.empty((chunk, rows, cols, 3), dtype=torch.uint8, device="cuda")
        dep_dev = torch.empty((chunk, rows, cols), dtype=torch.int16, device="cuda")
        for k0 in range(0, n_kf, chunk):                                               # warm-up (untimed): slot allocation
            m = min(chunk, n_kf - k0)
            ctx.synth_frames_dev(1, k0, m, rgb_dev.data_ptr(), dep_dev.data_ptr())
            ctx.set_frames_ptr(k0, m, rgb_dev.data_ptr(), dep_dev.data_ptr(), None, device=True)
        if len(ids):
            ctx.register_pairs(ap_all[ids[:B], 0], ap_all[ids[:B], 1])
        torch.cuda.synchronize()
        t_wall = time.perf_counter()
        for k0 in range(0, n_kf, chunk):                                               # keyframes replicated on every rank
            m = min(chunk, n_kf - k0)
            ctx.synth_frames_dev(1, k0, m, rgb_dev.data_ptr(), dep_dev.data_ptr())
            ctx.set_frames_ptr(k0, m, rgb_dev.data_ptr(), dep_dev.data_ptr(), None, device=True)
            dev_ms += ctx.last_device_ms()
        guesses = np.stack([r360.pose_to_colmajor(shard.loop_closure_guess(int(i), r360.synth_gt_pose(1, int(ap_all[i, 0]), int(ap_all[i, 1]))))
                            for i in ids]) if len(ids) else np.zeros((0, 16), np.float32)
        res = np.zeros(len(ids), r360.native.RESULT_DTYPE)
        for b0 in range(0, len(ids), B):
            sl = ids[b0:b0 + B]
            res[b0:b0 + len(sl)] = ctx.register_pairs(ap_all[sl, 0], ap_all[sl, 1], guesses[b0:b0 + len(sl)])
            dev_ms += ctx.last_device_ms()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t_wall
        gt = lambda gid: r360.synth_gt_pose(1, int(ap_all[gid, 0]), int(ap_all[gid, 1]))
        kind_name = "loop closure, all pairs over %d keyframes" % n_kf

    if world > 1:                                     # the first NCCL collective builds the communicator: not timed
        w = torch.zeros(world, device="cuda"); dist.all_gather_into_tensor(w, torch.ones(1, device="cuda"))
        torch.cuda.synchronize()
    t0 = time.perf_counter()
    allres = shard.allgather_results(res, pairs, n_pairs_total, device="cuda" if world > 1 else None)
    gather_s = time.perf_counter() - t0
    t = torch.tensor([dev_ms, wall * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms_max, wall_ms_max = [float(x) for x in t.tolist()]
    if rank == 0:
        ok = int((allres["status"] == 0).sum())
        n_gt = 0
        worst = (0.0, 0.0)
        step = max(1, n_pairs_total // 2048)                                          # ground-truth check on a subsample
        checked = 0
        for gid in range(0, n_pairs_total, step):
            ang, dist_m = pose_err(np.array(allres[gid]["pose"]).reshape(4, 4).T, gt(gid))
            good = ang < 5e-3 and dist_m < 1e-2
            n_gt += good
            checked += 1
            if good:
                worst = (max(worst[0], ang), max(worst[1], dist_m))
        line = {
            "config": args.config, "what": kind_name, "rows": rows, "cols": cols, "levels": L, "n_gpus": world,
            "pairs": n_pairs_total, "pairs_per_s_device": n_pairs_total / (dev_ms_max / 1e3),
            "pairs_per_s_wall_incl_synthesis": n_pairs_total / (wall_ms_max / 1e3),
            "device_ms_max_over_ranks": dev_ms_max, "allgather_s": gather_s,
            "pairs_status_ok": ok, "gathered_ids_complete": bool(np.array_equal(allres["pair_id"], np.arange(n_pairs_total))),
            "ground_truth_checked": checked, "ground_truth_within_5mrad_1cm": int(n_gt),
            "worst_converged_err_rad_m": worst,
            "mean_accepted_iters_per_level": [float(x) for x in allres["iters"][:, :L].mean(0)],
        }
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
