"""Host-side sharding of independent sphere pairs across the GPUs of one box.

Every pair is independent (two frames + initial guess in, one r360_result out), so ranks share
nothing on the data path; the only collective is one all-gather of the fixed-size result
records (NCCL over NVLink on GPUs, gloo in the CPU tests).
"""
import numpy as np

from .native import RESULT_DTYPE


def block_range(n, rank, world):
    """Contiguous block [lo, hi) of n items for `rank` (sizes differ by at most one)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def batch_pairs(n_pairs, rank, world):
    """Configs 2/3: pair j = (target frame 2j, source frame 2j+1), block-distributed.
    Returns (global pair ids, frame ids this rank must build, local src idx, local trg idx)."""
    lo, hi = block_range(n_pairs, rank, world)
    pairs = np.arange(lo, hi, dtype=np.int32)
    frames = np.arange(2 * lo, 2 * hi, dtype=np.int32)
    local = np.arange(hi - lo, dtype=np.int32)
    return pairs, frames, 2 * local + 1, 2 * local


def odometry_pairs(n_frames, rank, world):
    """Config 4: pair k = (target frame k, source frame k+1), contiguous ranges per rank with one
    halo frame at the range boundary (frame k is target of pair k and source of pair k-1)."""
    n_pairs = n_frames - 1
    lo, hi = block_range(n_pairs, rank, world)
    pairs = np.arange(lo, hi, dtype=np.int32)
    frames = np.arange(lo, hi + 1, dtype=np.int32) if hi > lo else np.zeros(0, np.int32)
    local = np.arange(hi - lo, dtype=np.int32)
    return pairs, frames, local + 1, local


def all_pairs(n_keyframes):
    """Config 5: all C(n,2) keyframe pairs, target = lower index."""
    i, j = np.triu_indices(n_keyframes, 1)
    return np.stack([j, i], 1).astype(np.int32)          # (src, trg)


def round_robin(n_items, rank, world):
    """Config 5: pair ids dealt round-robin (keyframes are replicated on every rank)."""
    return np.arange(rank, n_items, world, dtype=np.int32)


def frame_roles(src_local, trg_local, n_frames_local):
    """R360_ROLE_* bitmask per local frame from the pair lists."""
    roles = np.zeros(n_frames_local, np.uint8)
    roles[np.asarray(src_local, np.int64)] |= 1
    roles[np.asarray(trg_local, np.int64)] |= 2
    return roles


def allgather_results(local, pair_ids, n_total, device=None):
    """All-gather the per-pair result records of every rank; returns n_total records in global
    pair order on every rank.  `local`: np.ndarray of RESULT_DTYPE; `pair_ids`: their global ids.
    Works with any initialised torch.distributed backend (nccl: pass device='cuda')."""
    import torch
    import torch.distributed as dist

    local = np.ascontiguousarray(local)
    assert local.dtype == RESULT_DTYPE
    world = dist.get_world_size() if dist.is_initialized() else 1
    out = np.zeros(n_total, RESULT_DTYPE)
    if world == 1:
        ids = np.asarray(pair_ids, np.int64)
        out[ids] = local
        out["pair_id"][ids] = ids
        return out
    dev = torch.device(device) if device is not None else torch.device("cpu")
    counts = torch.zeros(world, dtype=torch.int64, device=dev)
    mine = torch.tensor([len(local)], dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(counts, mine)
    cap = int(counts.max().item())
    rec = RESULT_DTYPE.itemsize
    buf = np.zeros(cap * (rec + 4), np.uint8)             # records + int32 global ids, padded to cap
    buf[:len(local) * rec] = local.view(np.uint8).reshape(-1)
    ids = np.full(cap, -1, np.int32)
    ids[:len(local)] = pair_ids
    buf[cap * rec:] = ids.view(np.uint8)
    send = torch.from_numpy(buf).to(dev)
    recv = torch.empty(world * buf.size, dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(recv, send)
    allb = recv.cpu().numpy().reshape(world, buf.size)
    for r in range(world):
        n = int(counts[r].item())
        recs = allb[r, :n * rec].copy().view(RESULT_DTYPE)
        gid = allb[r, cap * rec:].copy().view(np.int32)[:n]
        out[gid] = recs
        out["pair_id"][gid] = gid
    return out


# ---------------------------------------------------------------- config 5: perturbed initial guesses
def _splitmix64(state):
    state = (state + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
    z = state
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
    return state, z ^ (z >> 31)


def loop_closure_guess(pair_id, gt_pose, max_t=0.05, max_w=0.02):
    """Initial guess of loop-closure candidate `pair_id`: the ground-truth
    relative pose composed with exp(delta), delta ~ U(+-max_t m, +-max_w rad) drawn from
    splitmix64(0xC105E + pair_id).  Returns a float32 4x4."""
    s = (0xC105E + int(pair_id)) & 0xFFFFFFFFFFFFFFFF
    u = []
    for _ in range(6):
        s, z = _splitmix64(s)
        u.append((z >> 11) * (1.0 / 9007199254740992.0) * 2.0 - 1.0)
    t = np.array(u[:3]) * max_t
    w = np.array(u[3:]) * max_w
    th = float(np.linalg.norm(w))
    K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    R = np.eye(3) + (np.sin(th) / th if th > 0 else 1.0) * K + ((1 - np.cos(th)) / (th * th) if th > 0 else 0.5) * K @ K
    D = np.eye(4)
    D[:3, :3] = R
    D[:3, 3] = t
    return (D @ np.asarray(gt_pose, np.float64)).astype(np.float32)
