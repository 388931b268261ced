"""Multi-GPU host logic on CPU: sharding of independent pairs and the all-gather of result
records, world_size 2 over gloo (the GPU run uses the same code over NCCL)."""
import os
import socket
import numpy as np
import pytest


def test_block_and_odometry_and_roundrobin_shards(r360):
    from rgbd360_b200 import shard
    for n, world in ((512, 8), (8192, 8), (10, 3), (3, 4)):
        covered = []
        for r in range(world):
            pairs, frames, s, t = shard.batch_pairs(n, r, world)
            covered += list(pairs)
            assert len(frames) == 2 * len(pairs)
            assert np.array_equal(frames[s], 2 * pairs + 1) and np.array_equal(frames[t], 2 * pairs)
        assert covered == list(range(n))
    n_frames = 8193
    tot = 0
    for r in range(8):
        pairs, frames, s, t = shard.odometry_pairs(n_frames, r, 8)
        assert len(pairs) == 1024 and len(frames) == 1025              # one halo frame per range
        assert np.array_equal(frames[t], pairs) and np.array_equal(frames[s], pairs + 1)
        roles = shard.frame_roles(s, t, len(frames))
        assert roles[0] == 2 and roles[-1] == 1 and np.all(roles[1:-1] == 3)
        tot += len(pairs)
    assert tot == 8192
    ap = shard.all_pairs(128)
    assert len(ap) == 8128 and np.all(ap[:, 1] < ap[:, 0])
    ids = np.concatenate([shard.round_robin(8128, r, 8) for r in range(8)])
    assert sorted(ids) == list(range(8128))


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from rgbd360_b200 import shard
    from rgbd360_b200.native import RESULT_DTYPE
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = 11                                                    # ragged: 6 + 5
    pairs, frames, s, t = shard.batch_pairs(n, rank, world)
    local = np.zeros(len(pairs), RESULT_DTYPE)
    local["pose"][:, 0] = pairs * 1.5
    local["final_n_valid"] = pairs + 100
    local["iters"][:, 2] = rank + 1
    allr = shard.allgather_results(local, pairs, n)
    ok = (np.array_equal(allr["final_n_valid"], np.arange(n) + 100) and np.allclose(allr["pose"][:, 0], np.arange(n) * 1.5)
          and np.array_equal(allr["pair_id"], np.arange(n)) and set(allr["iters"][:, 2]) == {1, 2})
    # round-robin (loop-closure) sharding gathers into the same global order
    ids = shard.round_robin(n, rank, world)
    loc2 = np.zeros(len(ids), RESULT_DTYPE)
    loc2["final_n_valid"] = ids * 7
    all2 = shard.allgather_results(loc2, ids, n)
    ok = ok and np.array_equal(all2["final_n_valid"], np.arange(n) * 7)
    dist.destroy_process_group()
    q.put((rank, bool(ok)))


def test_allgather_results_world2_gloo(r360):
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def test_single_rank_gather_is_identity(r360):
    from rgbd360_b200 import shard
    from rgbd360_b200.native import RESULT_DTYPE
    loc = np.zeros(4, RESULT_DTYPE); loc["sso"] = [1, 2, 3, 4]
    out = shard.allgather_results(loc, np.array([2, 0, 3, 1]), 4)
    assert list(out["sso"]) == [2, 4, 1, 3]
