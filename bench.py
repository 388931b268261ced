#!/usr/bin/env python
"""bench.py -- sphere-pair registrations/s of the B200 spherical dense registration path.

One "step" = one pass of the hot path over one batch of synthetic sphere pairs:
pyramid build of every frame (setSourceFrame / setTargetFrame) + batched alignFrames360
(+ at N > 1 the all-gather of the result records through the product's own collective,
r360_allgather_results, back to host memory on every rank).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                  [--workload A|B] [--pairs P] [--no-extra-configs]

Workload A (default, the configuration BASELINE.json's metric is quoted on, configs[2]): synthetic
2048x1024 sphere pairs, 4-level pyramid, photometric + depth with Huber weights, 512 pairs
per GPU (weak scaling).  Workload B (configs[1]): 1024x512, 3 levels, 256 pairs.

Prints ONE JSON line (rank 0).  `value`: inputs resident in HBM; `e2e`: same metric through the
C ABI with HOST (pinned) buffers, H2D of the frames and D2H of the (gathered) results inside the timed
region.  `configs`: the other BASELINE.json configurations measured in the same run, each with its own
roofline object -- config 2 (workload B, N = 1 only), config 4 (sequence odometry, 8192 pairs sharded
over the N GPUs), config 5 (loop closure, all pairs over 128 keyframes).  `verify`: rank 0 re-registers
the shard of rank N-1 and compares the gathered records bit for bit (outside the timed region).
`--impl reference` times the CPU oracle (port of the reference's own CPU path, all host threads) on a
bounded sample of the same workload.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    "A": dict(rows=1024, cols=2048, levels=4, pairs=512,
              name="synthetic 2048x1024 sphere pairs, 4-level pyramid, photo+depth Huber, batch 512 per GPU"),
    "B": dict(rows=512, cols=1024, levels=3, pairs=256,
              name="synthetic 1024x512 sphere pairs, 3-level pyramid, photo+depth Huber, batch 256 per GPU"),
}
METRIC = "sphere-pair registrations/s @2048x1024"
UNIT = "pairs/s"
KERNEL_NAME = ("k_pass<PHOTO_DEPTH, WITH_H> (warp/residual/normal-equation pixel passes: the fused launches and the "
               "speculative error-only ones, 32 B per source pixel each)")


def measured_traffic_ratio():
    """dram bytes / algorithmic bytes of one k_pass launch from the committed ncu --set full capture."""
    p = os.path.join(ROOT, "profiles", "k_pass_traffic.json")
    try:
        d = json.load(open(p))
        return float(d["dram_bytes"]) / float(d["algorithmic_bytes"]), d.get("source", p)
    except Exception:
        return None, None


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx = float(f[2])
            except ValueError:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx,
                "reasons": sorted(reasons), "samples": len(sm)}


def workload(args, key=None):
    w = dict(WORKLOADS[key or args.workload])
    if args.pairs and key is None:
        w["pairs"] = args.pairs
    return w


# ----------------------------------------------------------------------------- CPU reference arm
CPU_SAMPLE = {"A": 96, "B": 384}        # pairs of the workload timed on the host cores (~10-30 s of CPU work)


def cpu_reference_run(w, n_pairs, first_pair=0):
    """Oracle (port of the reference CPU path, FAITHFUL accumulation, glibc math, all host threads)
    on n_pairs pairs of the workload: setTargetFrame + setSourceFrame + alignFrames360 per pair.
    Frame synthesis is outside the timed region.  Returns (pairs/s, threads, seconds)."""
    from oracle import orc
    orc.build()
    orc.use_all_cores()
    orc.set_math(orc.MATH_LIBM)          # what a g++/glibc build of the reference calls
    P = orc.default_params(n_levels=w["levels"])
    dt = 0.0
    for k in range(n_pairs):
        ft = orc.synth_frame(0, 2 * (first_pair + k), w["rows"], w["cols"])
        fs = orc.synth_frame(0, 2 * (first_pair + k) + 1, w["rows"], w["cols"])
        t0 = time.perf_counter()
        trg = orc.Frame(ft[0], ft[1], P, True)        # setTargetFrame
        src = orc.Frame(fs[0], fs[1], P, False)       # setSourceFrame
        orc.align(src, trg, None, P, accum=orc.ACC_FAITHFUL)   # alignFrames360
        dt += time.perf_counter() - t0
    orc.set_math(orc.MATH_PINNED)
    return n_pairs / dt, orc.omp_threads(), dt


def compiled_reference_run(w, n_pairs=2):
    """The reference's own RegisterPhotoICP.h as compiled here against the from-scratch Eigen / OpenCV / MRPT
    stand-ins (oracle/_ref), all host threads, same call sequence.  Reported next to the port for
    transparency: the stand-ins evaluate eagerly through heap temporaries and were written for fidelity, not
    speed, so this figure understates what a real Eigen / OpenCV build of the reference would do."""
    try:
        from oracle import orc, refbind
        if not refbind.available():
            return None
        n_thr = len(os.sched_getaffinity(0))
        refbind.lib(False).ref_set_threads(n_thr)
        dt = 0.0
        for k in range(n_pairs):
            ft = orc.synth_frame(0, 2 * k, w["rows"], w["cols"])
            fs = orc.synth_frame(0, 2 * k + 1, w["rows"], w["cols"])
            R = refbind.Reference(n_levels=w["levels"], pinned=False)
            t0 = time.perf_counter()
            R.set_target(*ft); R.set_source(*fs)
            R.align(None, 2, 0)
            dt += time.perf_counter() - t0
            R.close()
        refbind.lib(False).ref_set_threads(1)
        return {"value": n_pairs / dt, "unit": UNIT, "cores": n_thr, "kind": "reference",
                "sample": "%d pairs, oracle/_ref (reference header compiled against stand-in Eigen/OpenCV/MRPT: "
                          "eager evaluation, heap temporaries -- a lower bound on the reference's CPU speed), %.1f s" % (n_pairs, dt)}
    except Exception as e:                      # noqa: BLE001
        return {"unavailable": repr(e)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = workload(args)
    sample = max(1, CPU_SAMPLE[args.workload] // 4)      # per step; the whole run stays within a few minutes
    for _ in range(args.warmup):
        cpu_reference_run(w, 2)
    cores, dt = 1, 0.0
    for s in range(args.steps):
        _, cores, d = cpu_reference_run(w, sample, first_pair=s * sample)
        dt += d
    v = args.steps * sample / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": w["name"], "rows": w["rows"], "cols": w["cols"], "levels": w["levels"],
                   "pairs_per_step": sample, "note": "CPU oracle: port of the reference, bit-identical to the "
                   "reference header compiled against third-party stand-ins (oracle/_ref, too slow to time fairly); "
                   "FAITHFUL accumulation, glibc math, OpenMP over all host threads"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{sample} pairs per step x {args.steps} steps, frame build + alignFrames360",
                         "compiled_reference": compiled_reference_run(w)},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ----------------------------------------------------------------------------- multi-GPU plumbing
class ProductComm:
    """The ncclComm_t handed to the product's collective (r360_allgather_results, include/r360.h): created with
    ncclCommInitRank through ctypes on the NCCL library torch already loaded; the unique id travels through
    torch.distributed.  torch.distributed itself is only used for barriers and the max-over-ranks of timings."""

    class UID(C.Structure):
        _fields_ = [("internal", C.c_byte * 128)]

    def __init__(self, rank, world, dist):
        path = "libnccl.so.2"
        try:
            for ln in open("/proc/self/maps"):
                if "libnccl.so" in ln:
                    path = ln.split()[-1]
                    break
        except OSError:
            pass
        self.path = path
        self.lib = C.CDLL(path)
        self.lib.ncclGetErrorString.restype = C.c_char_p
        uid = self.UID()
        if rank == 0:
            self._ck(self.lib.ncclGetUniqueId(C.byref(uid)))
        box = [bytes(bytearray(uid.internal))]
        dist.broadcast_object_list(box, src=0)
        C.memmove(C.byref(uid), box[0], 128)
        self.comm = C.c_void_p()
        self.lib.ncclCommInitRank.argtypes = [C.POINTER(C.c_void_p), C.c_int, self.UID, C.c_int]
        self._ck(self.lib.ncclCommInitRank(C.byref(self.comm), world, uid, rank))
        n = C.c_int(0)
        self._ck(self.lib.ncclCommCount(self.comm, C.byref(n)))
        self.n_ranks = n.value

    def _ck(self, rc):
        if rc != 0:
            raise RuntimeError("NCCL: " + self.lib.ncclGetErrorString(rc).decode())

    def close(self):
        if self.comm:
            self.lib.ncclCommDestroy(self.comm)
            self.comm = None


class Dist:
    """Process-group state of one bench run (one process per GPU)."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; the registration path has no CPU fallback")
        torch.cuda.set_device(self.local)
        self.comm = None
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
            self.comm = ProductComm(self.rank, self.world, dist)

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def reduce(self, values, op="max"):
        t = self.torch.tensor(values, dtype=self.torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX if op == "max" else self.dist.ReduceOp.SUM)
        return [float(x) for x in t.tolist()]

    def close(self):
        if self.comm:
            self.comm.close()
        if self.world > 1:
            self.dist.destroy_process_group()


_CUDART = None


def cudart():
    """The CUDA runtime torch loaded, through ctypes (plain cudaMemcpyAsync for the copy-ceiling measurement)."""
    global _CUDART
    if _CUDART is None:
        path = "libcudart.so.12"
        for ln in open("/proc/self/maps"):
            if "libcudart.so" in ln:
                path = ln.split()[-1]
                break
        _CUDART = C.CDLL(path)
        _CUDART.cudaMemcpyAsync.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
    return _CUDART


def roofline_object(pass_bytes, pass_ms, pass_launches, dev_ms):
    peak, peak_src = measured_peak()
    achieved = (pass_bytes / 1e9) / (pass_ms / 1e3) if pass_ms > 0 else 0.0
    t_ratio, t_src = measured_traffic_ratio()
    return {"bound": "hbm", "kernel": KERNEL_NAME, "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": achieved / peak, "peak_source": peak_src,
            "peak_nominal": 8000.0, "frac_nominal": achieved / 8000.0,   # north_star's "~8 TB/s" (report both)
            "traffic": (t_ratio * pass_bytes / max(pass_launches, 1)) if t_ratio else None,
            "traffic_source": t_src,
            "alg_bytes_per_launch": pass_bytes / max(pass_launches, 1),
            "avg_launch_ms": pass_ms / max(pass_launches, 1), "launches": pass_launches,
            "kernel_share_of_step": pass_ms / dev_ms if dev_ms else None}


# ----------------------------------------------------------------------------- GPU arm: batch workloads (configs 2 / 3)
def run_batch(D, args, w, steps, warmup, want_clocks, occlusion=0):
    """Workload A / B on every rank: `steps` timed resident steps, then the e2e leg from pinned host memory.
    Returns the pieces of the JSON line (rank 0) -- None on the other ranks."""
    import rgbd360_b200 as r360
    torch = D.torch
    world, rank, local = D.world, D.rank, D.local
    rows, cols, L, n_pairs = w["rows"], w["cols"], w["levels"], w["pairs"]
    n_frames = 2 * n_pairs
    npx = rows * cols

    params = r360.default_params(n_levels=L, occlusion=occlusion)
    ctx = r360.Context(rows, cols, n_frames, n_pairs, params, device=local)
    # synthetic frames rendered on the device: pair j = (target frame 2j, source frame 2j+1),
    # frame ids offset per rank so every GPU registers different pairs (weak scaling)
    rgb_dev = torch.empty((n_frames, rows, cols, 3), dtype=torch.uint8, device="cuda")
    dep_dev = torch.empty((n_frames, rows, cols), dtype=torch.int16, device="cuda")
    torch.cuda.synchronize()
    ctx.synth_frames_dev(0, rank * n_frames, n_frames, rgb_dev.data_ptr(), dep_dev.data_ptr())
    roles = np.array([r360.ROLE_TARGET, r360.ROLE_SOURCE] * n_pairs, np.uint8)
    trg_idx = np.arange(0, n_frames, 2, dtype=np.int32)
    src_idx = trg_idx + 1
    res = np.zeros(n_pairs, r360.native.RESULT_DTYPE)
    res_bytes = res.nbytes
    gathered = [None]

    def gather_results():
        # the product's collective: host records -> device -> ncclAllGather over NVLink -> host records of ALL ranks
        if world > 1:
            gathered[0] = ctx.allgather_results(D.comm.comm, res, world)

    pyr_ms = [0.0]                          # device time of the pyramid builds (set*Frame) inside the timed steps

    def step_resident():
        ctx.set_frames_ptr(0, n_frames, rgb_dev.data_ptr(), dep_dev.data_ptr(), roles, device=True)
        ms = ctx.last_device_ms()
        pyr_ms[0] += ms
        ctx.register_pairs(src_idx, trg_idx, None, out=res)
        ms += ctx.last_device_ms()
        gather_results()
        return ms

    if args.one_step:
        # profiling aid (ncu launch lists / captures): exactly ONE resident step, no warm-up, no e2e leg, no JSON
        # line -- the launch list of this command is the launch list of a step.  Never a bench number.
        step_resident()
        D.barrier()
        sys.stderr.write("one step: %d launches, %.2f ms on the device (cold)\n" % (ctx.kernel_launches(), ctx.last_device_ms()))
        ctx.close()
        return None
    for _ in range(max(warmup, 3)):
        step_resident()
    D.barrier()
    sampler = ClockSampler(local)
    if rank == 0 and want_clocks:
        sampler.start()
    launches0 = ctx.kernel_launches()
    pyr_ms[0] = 0.0
    dev_ms, pass_ms, pass_bytes, pass_launches = 0.0, 0.0, 0.0, 0
    t0 = time.perf_counter()
    for _ in range(steps):
        dev_ms += step_resident()
        ps = ctx.last_pass_stats()
        pass_ms += ps["ms"]; pass_bytes += ps["alg_bytes"]; pass_launches += ps["launches"]
    D.barrier()
    wall_ms = 1e3 * (time.perf_counter() - t0)
    launches = ctx.kernel_launches() - launches0
    clocks = sampler.stop() if (rank == 0 and want_clocks) else None
    res_resident = res.copy()
    gathered_resident = None if gathered[0] is None else gathered[0].copy()

    # ---- e2e: host pinned frames -> C ABI -> host results (H2D / D2H inside the timed region).  The staging
    #      memory comes from the library (r360_host_alloc: pinned, on the GPU's NUMA node where the platform says which)
    rgb_addr, numa_node, numa_note = ctx.host_alloc(n_frames * npx * 3)
    dep_addr, _, _ = ctx.host_alloc(n_frames * npx * 2)
    rgb_host = np.ctypeslib.as_array((C.c_uint8 * (n_frames * npx * 3)).from_address(rgb_addr))
    dep_host = np.ctypeslib.as_array((C.c_uint16 * (n_frames * npx)).from_address(dep_addr))
    torch.from_numpy(rgb_host).copy_(rgb_dev.view(-1))
    torch.from_numpy(dep_host.view(np.int16)).copy_(dep_dev.view(-1))
    torch.cuda.synchronize()

    def step_e2e():
        # one C-ABI call: host frames in, host results out (uploads, pyramid builds and batched
        # registrations overlap inside; pair p = (target frame 2p, source frame 2p+1)); then the gather
        ctx.register_host_pairs(rgb_addr, dep_addr, n_pairs, None, out=res)
        gather_results()

    step_e2e()
    D.barrier()
    e2e_steps = max(1, min(steps, 5))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_e2e()
    D.barrier()
    e2e_ms = 1e3 * (time.perf_counter() - t0)

    # ---- pure-copy ceiling of the same upload: every rank copies its pinned frames to its device buffers at the same
    #      time, nothing else running.  The e2e leg cannot be faster than this; how close it gets is the software's share.
    copy_ms = 0.0
    if not args.no_copy_ceiling:
        rt = cudart()
        for it in range(3):
            D.barrier()
            t0 = time.perf_counter()
            rt.cudaMemcpyAsync(C.c_void_p(rgb_dev.data_ptr()), C.c_void_p(rgb_addr), C.c_size_t(n_frames * npx * 3), 1, None)
            rt.cudaMemcpyAsync(C.c_void_p(dep_dev.data_ptr()), C.c_void_p(dep_addr), C.c_size_t(n_frames * npx * 2), 1, None)
            D.barrier()
            if it:
                copy_ms += 1e3 * (time.perf_counter() - t0) / 2
        ctx.synth_frames_dev(0, rank * n_frames, n_frames, rgb_dev.data_ptr(), dep_dev.data_ptr())   # (same bytes; keeps the buffers defined)

    # ---- verify (outside every timed region): the gathered records of the LAST rank, recomputed on rank 0 from the
    #      same synthetic frames with the same batch composition, must be identical bit for bit -- the path is
    #      deterministic (order-independent fixed-point accumulation) and the gather returns what was sent.
    #      At N = 1: a second run of the same step against the first.
    verify = None
    other = world - 1
    if rank == 0:
        ctx.synth_frames_dev(0, other * n_frames, n_frames, rgb_dev.data_ptr(), dep_dev.data_ptr())
        again = np.zeros(n_pairs, r360.native.RESULT_DTYPE)
        ctx.set_frames_ptr(0, n_frames, rgb_dev.data_ptr(), dep_dev.data_ptr(), roles, device=True)
        ctx.register_pairs(src_idx, trg_idx, None, out=again)
        want = res_resident if world == 1 else gathered_resident[other * n_pairs:(other + 1) * n_pairs]
        same = want.tobytes() == again.tobytes()
        n_diff = 0 if same else int(sum(want[k].tobytes() != again[k].tobytes() for k in range(n_pairs)))
        verify = {"ok": bool(same), "records": n_pairs, "records_differing": n_diff,
                  "what": ("rank 0 re-registered the %d pairs of rank %d; compared with the records gathered through "
                           "r360_allgather_results, all bytes" % (n_pairs, other)) if world > 1 else
                          "the same resident step run twice, all bytes of all records (determinism)"}
    D.barrier()

    # ---- max over ranks
    dev_ms, wall_ms, e2e_ms, _, copy_ms = D.reduce([dev_ms, wall_ms, e2e_ms, pass_ms, copy_ms])
    launches_all = int(D.reduce([float(launches)], "sum")[0])
    total_pairs = n_pairs * world
    out = None
    if rank == 0:
        iters = res_resident["iters"][:, :L].astype(np.float64)
        passes = res_resident["passes"][:, :L].astype(np.float64)
        h2d = int(n_frames * npx * 5 + n_pairs * 8 * 3)
        d2h = int(res_bytes * (world if world > 1 else 1))
        e2e_s = e2e_ms / 1e3 / e2e_steps
        out = {
            "value": total_pairs * steps / (wall_ms / 1e3), "ms_per_step": wall_ms / steps,
            "device_ms_per_step": dev_ms / steps, "pyramid_ms_per_step": pyr_ms[0] / steps,
            "config": {"workload": w["name"], "rows": rows, "cols": cols, "levels": L,
                       "pairs_per_gpu": n_pairs, "frames_per_gpu": n_frames, "method": "PHOTO_DEPTH",
                       "occlusion": occlusion,
                       "sampling": "nearest-neighbour (reference semantics)",
                       "l2": "inputs larger than L2 (pyramids %.1f GB per GPU)" % (
                           (8 + 24) * rows * cols * sum(0.25 ** l for l in range(L)) * n_pairs / 1e9),
                       "step": "pyramid build of all frames + batched alignFrames360" +
                               (" + r360_allgather_results (NCCL all-gather of the result records, D2H on every rank)" if world > 1 else ""),
                       "collective": ("r360_allgather_results (C ABI; ncclAllGather on a communicator of %d ranks created "
                                      "with ncclCommInitRank, %s)" % (D.comm.n_ranks, os.path.basename(D.comm.path))) if world > 1 else None,
                       "mean_accepted_iters_per_level": [float(x) for x in iters.mean(0)],
                       "mean_passes_per_level": [float(x) for x in passes.mean(0)],
                       "pairs_ok": int((res_resident["status"] == 0).sum()),
                       "host_staging": {"numa_node_rank0": None if numa_node < 0 else numa_node, "how": numa_note}},
            "e2e": {"value": total_pairs / e2e_s, "unit": UNIT,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                    "ms_per_step": 1e3 * e2e_s,
                    "h2d_gbs_per_rank": h2d / e2e_s / 1e9,
                    "h2d_copy_ceiling_gbs_per_rank": (h2d / (copy_ms / 1e3) / 1e9) if copy_ms > 0 else None,
                    "frac_of_copy_ceiling": (copy_ms / 1e3 / e2e_s) if copy_ms > 0 else None,
                    "note": "h2d_copy_ceiling: the same pinned frames copied by all ranks at once with nothing else running "
                            "(slowest rank); the e2e step cannot beat it"},
            "gpu_launches": launches_all, "clocks": clocks,
            "roofline": roofline_object(pass_bytes, pass_ms, pass_launches, dev_ms),
            "verify": verify,
        }
    ctx.host_free(rgb_addr); ctx.host_free(dep_addr)
    ctx.close()
    del rgb_dev, dep_dev
    torch.cuda.empty_cache()
    return out


# ----------------------------------------------------------------------------- GPU arm: configs 4 / 5 at their named scale
def pose_err(T, G):
    T = np.asarray(T, np.float64); G = np.asarray(G, np.float64)
    dR = T[:3, :3] @ G[:3, :3].T
    v = 0.5 * np.array([dR[2, 1] - dR[1, 2], dR[0, 2] - dR[2, 0], dR[1, 0] - dR[0, 1]])
    return float(np.arctan2(np.linalg.norm(v), (np.trace(dR) - 1) / 2)), float(np.linalg.norm(T[:3, 3] - G[:3, 3]))


def gather_ragged(D, ctx, res, gids, n_total, r360):
    """All ranks' records in global pair order through r360_allgather_results: every rank sends the same count
    (a ragged last shard is padded with records whose pair_id is -1, as include/r360.h says)."""
    res = res.copy()
    res["pair_id"] = gids
    if D.world == 1:
        out = np.zeros(n_total, r360.native.RESULT_DTYPE)
        out[gids] = res
        return out
    cap = int(D.reduce([float(len(res))])[0])
    send = np.zeros(cap, r360.native.RESULT_DTYPE)
    send["pair_id"] = -1
    send[:len(res)] = res
    allr = ctx.allgather_results(D.comm.comm, send, D.world)
    allr = allr[allr["pair_id"] >= 0]
    out = np.zeros(n_total, r360.native.RESULT_DTYPE)
    out["pair_id"] = -1
    out[allr["pair_id"]] = allr
    return out


def config_summary(D, name, what, n_total, dev_ms, wall_ms, gather_ms, stats, allres, gt, L, extra):
    dev_ms, wall_ms, gather_ms = D.reduce([dev_ms, wall_ms, gather_ms])
    if D.rank != 0:
        return None
    n_gt, checked = 0, 0
    step = max(1, n_total // 1024)                                            # ground-truth check on a subsample
    for gid in range(0, n_total, step):
        ang, dist_m = pose_err(np.array(allres[gid]["pose"]).reshape(4, 4).T, gt(gid))
        n_gt += int(ang < 5e-3 and dist_m < 1e-2)
        checked += 1
    d = {"workload": what, "n_gpus": D.world, "pairs": n_total, "value": n_total / (dev_ms / 1e3), "unit": UNIT,
         "timing": "sum of the CUDA-event times of the product calls (pyramid builds + registrations), max over ranks; "
                   "frames rendered on the device between the calls",
         "device_ms": dev_ms, "wall_ms_incl_synthesis": wall_ms, "allgather_ms": gather_ms,
         "pairs_ok": int((allres["status"] == 0).sum()),
         "gathered_ids_complete": bool(np.array_equal(allres["pair_id"], np.arange(n_total))),
         "ground_truth_checked": checked, "ground_truth_within_5mrad_1cm": n_gt,
         "mean_accepted_iters_per_level": [float(x) for x in allres["iters"][:, :L].mean(0)],
         "roofline": roofline_object(stats["bytes"], stats["ms"], stats["launches"], dev_ms)}
    d.update(extra)
    return d


def run_config4(D, args, n_frames_total=8193, block=512):
    """BASELINE configs[3]: sequence odometry, 8 193 consecutive synthetic 2048x1024 frames -> 8 192 pairs (target = frame
    k, source = frame k + 1, guess Identity; setNumPyr(4), stdDevPhoto 3/255 as OdometryRGBD360.cpp:92), contiguous shards
    with one halo frame; every inner frame is resident ONCE with both roles; a shard streams through the GPU in blocks."""
    import rgbd360_b200 as r360
    from rgbd360_b200 import shard
    torch = D.torch
    rows, cols, L = 1024, 2048, 4
    n_total = n_frames_total - 1
    pairs, frames, _, _ = shard.odometry_pairs(n_frames_total, D.rank, D.world)
    gp = r360.default_params(n_levels=L, std_photo=np.float32(3.0 / 255))
    B = min(block, max(len(pairs), 1))
    ctx = r360.Context(rows, cols, B + 1, B, gp, device=D.local)
    rgb_dev = torch.empty((B + 1, rows, cols, 3), dtype=torch.uint8, device="cuda")
    dep_dev = torch.empty((B + 1, rows, cols), dtype=torch.int16, device="cuda")
    res = np.zeros(len(pairs), r360.native.RESULT_DTYPE)
    stats = {"ms": 0.0, "bytes": 0.0, "launches": 0}
    dev_ms = 0.0

    def run_block(b0, timed):
        nonlocal dev_ms
        nb = min(B, len(pairs) - b0)
        f0 = int(frames[0]) + b0                                                   # block frames f0 .. f0 + nb
        ctx.synth_frames_dev(0, f0, nb + 1, rgb_dev.data_ptr(), dep_dev.data_ptr())
        roles = np.full(nb + 1, r360.ROLE_BOTH, np.uint8)
        roles[0] = r360.ROLE_TARGET; roles[nb] = r360.ROLE_SOURCE                   # block-boundary frames: one role here
        ctx.set_frames_ptr(0, nb + 1, rgb_dev.data_ptr(), dep_dev.data_ptr(), roles, device=True)
        ms = ctx.last_device_ms()
        loc = np.arange(nb, dtype=np.int32)
        res[b0:b0 + nb] = ctx.register_pairs(loc + 1, loc)                         # source k + 1 -> target k
        ms += ctx.last_device_ms()
        if timed:
            dev_ms += ms
            ps = ctx.last_pass_stats()
            stats["ms"] += ps["ms"]; stats["bytes"] += ps["alg_bytes"]; stats["launches"] += ps["launches"]

    if len(pairs):
        run_block(0, False)                                                        # warm-up: slots are allocated on first use
    D.barrier()
    t0 = time.perf_counter()
    for b0 in range(0, len(pairs), B):
        run_block(b0, True)
    D.barrier()
    wall_ms = 1e3 * (time.perf_counter() - t0)
    t0 = time.perf_counter()
    allres = gather_ragged(D, ctx, res, pairs, n_total, r360)
    gather_ms = 1e3 * (time.perf_counter() - t0)
    out = config_summary(D, "config4", "sequence odometry: %d consecutive synthetic 2048x1024 frames -> %d pairs, 4 levels, "
                         "contiguous shards + halo frame, blocks of %d pairs, results all-gathered (r360_allgather_results)"
                         % (n_frames_total, n_total, B), n_total, dev_ms, wall_ms, gather_ms, stats, allres,
                         lambda gid: r360.synth_gt_pose(0, int(gid) + 1, int(gid)), L,
                         {"frames_per_rank": int(len(frames)), "std_photo": "3/255 (OdometryRGBD360.cpp:92)"})
    ctx.close()
    del rgb_dev, dep_dev
    torch.cuda.empty_cache()
    return out


def run_config5(D, args, n_kf=128, block=512):
    """BASELINE configs[4]: loop closure, all C(128, 2) = 8 128 pairs over 128 synthetic 2048x1024 keyframes replicated on
    every rank, pairs dealt round-robin, guess = ground truth o exp(delta)."""
    import rgbd360_b200 as r360
    from rgbd360_b200 import shard
    torch = D.torch
    rows, cols, L = 1024, 2048, 4
    ap_all = shard.all_pairs(n_kf)
    n_total = len(ap_all)
    ids = shard.round_robin(n_total, D.rank, D.world)
    gp = r360.default_params(n_levels=L)
    B = min(block, max(len(ids), 1))
    ctx = r360.Context(rows, cols, n_kf, B, gp, device=D.local)
    chunk = 64
    rgb_dev = torch.empty((chunk, rows, cols, 3), dtype=torch.uint8, device="cuda")
    dep_dev = torch.empty((chunk, rows, cols), dtype=torch.int16, device="cuda")
    stats = {"ms": 0.0, "bytes": 0.0, "launches": 0}
    dev_ms = 0.0

    def build_keyframes(timed):
        nonlocal dev_ms
        for k0 in range(0, n_kf, chunk):                                           # keyframes replicated on every rank
            m = min(chunk, n_kf - k0)
            ctx.synth_frames_dev(1, k0, m, rgb_dev.data_ptr(), dep_dev.data_ptr())
            ctx.set_frames_ptr(k0, m, rgb_dev.data_ptr(), dep_dev.data_ptr(), None, device=True)
            if timed:
                dev_ms += ctx.last_device_ms()

    build_keyframes(False)                                                         # warm-up: slot allocation
    if len(ids):
        ctx.register_pairs(ap_all[ids[:B], 0], ap_all[ids[:B], 1])
    guesses = np.stack([r360.pose_to_colmajor(shard.loop_closure_guess(int(i), r360.synth_gt_pose(1, int(ap_all[i, 0]), int(ap_all[i, 1]))))
                        for i in ids]) if len(ids) else np.zeros((0, 16), np.float32)
    res = np.zeros(len(ids), r360.native.RESULT_DTYPE)
    D.barrier()
    t0 = time.perf_counter()
    build_keyframes(True)
    for b0 in range(0, len(ids), B):
        sl = ids[b0:b0 + B]
        res[b0:b0 + len(sl)] = ctx.register_pairs(ap_all[sl, 0], ap_all[sl, 1], guesses[b0:b0 + len(sl)])
        dev_ms += ctx.last_device_ms()
        ps = ctx.last_pass_stats()
        stats["ms"] += ps["ms"]; stats["bytes"] += ps["alg_bytes"]; stats["launches"] += ps["launches"]
    D.barrier()
    wall_ms = 1e3 * (time.perf_counter() - t0)
    t0 = time.perf_counter()
    allres = gather_ragged(D, ctx, res, ids, n_total, r360)
    gather_ms = 1e3 * (time.perf_counter() - t0)
    out = config_summary(D, "config5", "loop closure: all C(%d, 2) = %d pairs over %d synthetic 2048x1024 keyframes replicated "
                         "on every rank, 4 levels, round-robin, guess = ground truth o exp(delta), results all-gathered "
                         "(r360_allgather_results)" % (n_kf, n_total, n_kf), n_total, dev_ms, wall_ms, gather_ms, stats, allres,
                         lambda gid: r360.synth_gt_pose(1, int(ap_all[gid, 0]), int(ap_all[gid, 1])), L,
                         {"keyframes_per_rank": n_kf})
    ctx.close()
    del rgb_dev, dep_dev
    torch.cuda.empty_cache()
    return out


def run_latency(D, args, w, reps=20):
    """The call shape of the reference's callers (OdometryRGBD360.cpp:189-193): ONE pair per call, frames handed over in
    host memory -- setTargetFrame + setSourceFrame + alignFrames360.  Small calls run in latency mode: the host watches the
    active lists and stops enqueueing the passes of a level once the pair has left it."""
    import rgbd360_b200 as r360
    rows, cols, L = w["rows"], w["cols"], w["levels"]
    ctx = r360.Context(rows, cols, 2, 1, r360.default_params(n_levels=L), device=D.local)
    out = {}
    for pair_id in (0, 7):                                       # two different pairs (iteration counts differ)
        rgb, dep = ctx.synth_frames(0, 2 * pair_id, 2)
        roles = [r360.ROLE_TARGET, r360.ROLE_SOURCE]
        t_align, t_full, launches = [], [], 0
        for k in range(reps + 3):
            t0 = time.perf_counter()
            ctx.set_frames(0, rgb, dep, roles)
            t1 = time.perf_counter()
            l0 = ctx.kernel_launches()
            res = ctx.register_pairs([1], [0])
            t2 = time.perf_counter()
            if k >= 3:
                t_align.append(1e3 * (t2 - t1)); t_full.append(1e3 * (t2 - t0)); launches = ctx.kernel_launches() - l0
        out["pair_%d" % pair_id] = {"ms_alignFrames360": float(np.median(t_align)),
                                    "ms_setFrames_plus_alignFrames360": float(np.median(t_full)),
                                    "kernel_launches_alignFrames360": int(launches),
                                    "accepted_iters_per_level": [int(x) for x in res[0]["iters"][:L]],
                                    "status": int(res[0]["status"])}
    ctx.close()
    out["what"] = ("one %dx%d pair per call through the C ABI (host wall clock, median of %d calls): frames from pageable host memory, "
                   "pyramids, registration, result back on the host" % (cols, rows, reps))
    return out


# ----------------------------------------------------------------------------- GPU arm: the JSON line
def run_ours(args):
    D = Dist()
    full_affinity = os.sched_getaffinity(0)
    w = workload(args)
    main = run_batch(D, args, w, args.steps, args.warmup, True, args.occlusion)
    if args.one_step:
        D.close()
        return
    extras = {}
    if not args.no_extra_configs and args.occlusion == 0:
        if D.world == 1 and args.workload == "A" and not args.pairs:
            b = run_batch(D, args, workload(args, "B"), min(args.steps, 10), 3, False)
            if b:
                extras["config2"] = {"workload": b["config"]["workload"], "n_gpus": 1, "value": b["value"], "unit": UNIT,
                                     "ms_per_step": b["ms_per_step"], "pyramid_ms_per_step": b["pyramid_ms_per_step"],
                                     "pairs_ok": b["config"]["pairs_ok"],
                                     "mean_passes_per_level": b["config"]["mean_passes_per_level"],
                                     "e2e": b["e2e"], "roofline": b["roofline"], "verify": b["verify"]}
        extras["config4"] = run_config4(D, args)
        extras["config5"] = run_config5(D, args)
    latency = run_latency(D, args, w) if (D.world == 1 and D.rank == 0 and not args.no_extra_configs) else None
    if D.rank == 0:
        line = {
            "metric": METRIC, "value": main["value"], "unit": UNIT, "n_gpus": D.world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": main["ms_per_step"],
            "device_ms_per_step": main["device_ms_per_step"], "pyramid_ms_per_step": main["pyramid_ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": main["config"], "e2e": main["e2e"], "gpu_launches": main["gpu_launches"], "clocks": main["clocks"],
            "roofline": main["roofline"], "verify": main["verify"],
        }
        if extras:
            line["configs"] = extras
        if latency:
            line["latency_batch1"] = latency
        if D.world == 1 and not args.no_cpu_baseline:
            n_cpu = CPU_SAMPLE[args.workload]
            os.sched_setaffinity(0, full_affinity)                     # the CPU baseline gets every host core
            cpu_reference_run(w, 2)                                    # warm-up (library build, page-in)
            v, cores, dt = cpu_reference_run(w, n_cpu)
            if latency:
                latency["cpu_reference_ms_per_pair"] = 1e3 / v
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": "%d pairs of the same workload (frame build + alignFrames360, "
                                              "FAITHFUL accumulation, glibc math, OpenMP), %.1f s" % (n_cpu, dt),
                                    "compiled_reference": compiled_reference_run(w)}
        emit(line)
    D.close()


_REAL_STDOUT = None


def emit(line):
    """The ONE JSON line of the contract, on the process's original stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is not None:
        os.write(_REAL_STDOUT, data)
    else:
        sys.stdout.write(data.decode()); sys.stdout.flush()


def main():
    global _REAL_STDOUT
    # libraries (NCCL's version banner, ...) may write to fd 1: keep it for the JSON line only
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="A", choices=list(WORKLOADS))
    ap.add_argument("--pairs", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra-configs", action="store_true",
                    help="skip the `configs` object (BASELINE configs 2, 4 and 5 measured after the headline workload)")
    ap.add_argument("--no-copy-ceiling", action="store_true", help="skip the pure H2D copy measurement of the e2e object")
    ap.add_argument("--one-step", action="store_true",
                    help="profiling aid: run exactly one resident step (no warm-up, no e2e, no JSON line) and exit")
    ap.add_argument("--occlusion", type=int, default=0, choices=[0, 1, 2],
                    help="alignFrames360's occlusion argument (side measurement; the headline metric is occlusion 0, "
                         "whose fused pass the roofline object describes -- with 1 / 2 that object is empty)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
