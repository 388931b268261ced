#!/usr/bin/env python
"""bench.py -- sphere-pair registrations/s of the B200 spherical dense registration path.

One "step" = one pass of the hot path over one batch of synthetic sphere pairs:
pyramid build of every frame (setSourceFrame / setTargetFrame) + batched alignFrames360.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                  [--workload A|B] [--pairs P]

Workload A (default, the configuration BASELINE.json's metric is quoted on): synthetic
2048x1024 sphere pairs, 4-level pyramid, photometric + depth with Huber weights, 512 pairs
per GPU (weak scaling).  Workload B: 1024x512, 3 levels, 256 pairs.

Prints ONE JSON line (rank 0).  `value`: inputs resident in HBM; `e2e`: same metric through the
C ABI with HOST (pinned) buffers, H2D of the frames and D2H of the results inside the timed
region.  `--impl reference` times the CPU oracle (port of the reference's own CPU path, all host
threads) on a bounded sample of the same workload.
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


def workload(args):
    w = dict(WORKLOADS[args.workload])
    if args.pairs:
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


# ----------------------------------------------------------------------------- GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    import rgbd360_b200 as r360

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the registration path has no CPU fallback")
    torch.cuda.set_device(local)
    full_affinity = os.sched_getaffinity(0)
    numa_node = bind_to_gpu_numa_node(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    w = workload(args)
    rows, cols, L, n_pairs = w["rows"], w["cols"], w["levels"], w["pairs"]
    n_frames = 2 * n_pairs
    npx = rows * cols

    params = r360.default_params(n_levels=L, occlusion=args.occlusion)
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
    gather_in = torch.empty(res_bytes, dtype=torch.uint8, device="cuda") if world > 1 else None
    gather_out = torch.empty(res_bytes * world, dtype=torch.uint8, device="cuda") if world > 1 else None

    def gather_results():
        if world > 1:
            gather_in.copy_(torch.from_numpy(res.view(np.uint8)), non_blocking=False)
            dist.all_gather_into_tensor(gather_out, gather_in)

    pyr_ms = [0.0]                          # device time of the pyramid builds (set*Frame) inside the timed steps

    def step_resident():
        ctx.set_frames_ptr(0, n_frames, rgb_dev.data_ptr(), dep_dev.data_ptr(), roles, device=True)
        ms = ctx.last_device_ms()
        pyr_ms[0] += ms
        ctx.register_pairs(src_idx, trg_idx, None, out=res)
        ms += ctx.last_device_ms()
        gather_results()
        return ms

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if args.one_step:
        # profiling aid (ncu launch lists / captures): exactly ONE resident step, no warm-up, no e2e leg, no JSON
        # line -- the launch list of this command is the launch list of a step.  Never a bench number.
        step_resident()
        barrier()
        sys.stderr.write("one step: %d launches, %.2f ms on the device (cold)\n" % (ctx.kernel_launches(), ctx.last_device_ms()))
        ctx.close()
        return
    for _ in range(max(args.warmup, 3)):
        step_resident()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = ctx.kernel_launches()
    pyr_ms[0] = 0.0
    dev_ms, pass_ms, pass_bytes, pass_launches = 0.0, 0.0, 0.0, 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        dev_ms += step_resident()
        ps = ctx.last_pass_stats()
        pass_ms += ps["ms"]; pass_bytes += ps["alg_bytes"]; pass_launches += ps["launches"]
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - t0)
    launches = ctx.kernel_launches() - launches0
    clocks = sampler.stop() if rank == 0 else None

    # ---- e2e: host pinned frames -> C ABI -> host results (H2D / D2H inside the timed region)
    rgb_host = torch.empty((n_frames, rows, cols, 3), dtype=torch.uint8, pin_memory=True)
    dep_host = torch.empty((n_frames, rows, cols), dtype=torch.int16, pin_memory=True)
    rgb_host.copy_(rgb_dev); dep_host.copy_(dep_dev)
    torch.cuda.synchronize()

    def step_e2e():
        # one C-ABI call: host frames in, host results out (uploads, pyramid builds and batched
        # registrations overlap inside; pair p = (target frame 2p, source frame 2p+1))
        ctx.register_host_pairs(rgb_host.data_ptr(), dep_host.data_ptr(), n_pairs, None, out=res)
        gather_results()

    step_e2e()
    barrier()
    e2e_steps = max(1, min(args.steps, 5))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_e2e()
    barrier()
    e2e_ms = 1e3 * (time.perf_counter() - t0)

    # ---- max over ranks
    t = torch.tensor([dev_ms, wall_ms, e2e_ms, pass_ms], dtype=torch.float64, device="cuda")
    s = torch.tensor([pass_bytes, float(launches)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(s, op=dist.ReduceOp.SUM)
    dev_ms, wall_ms, e2e_ms, pass_ms_max = [float(x) for x in t.tolist()]
    launches = int(s[1].item())                     # all ranks
    total_pairs = n_pairs * world
    value = total_pairs * args.steps / (wall_ms / 1e3)
    e2e_value = total_pairs * e2e_steps / (e2e_ms / 1e3)

    if rank == 0:
        peak, peak_src = measured_peak()
        achieved = (pass_bytes / 1e9) / (pass_ms / 1e3) if pass_ms > 0 else 0.0     # rank 0's kernel
        t_ratio, t_src = measured_traffic_ratio()
        iters = res["iters"][:, :L].astype(np.float64)
        passes = res["passes"][:, :L].astype(np.float64)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": wall_ms / args.steps,
            "device_ms_per_step": dev_ms / args.steps,
            "pyramid_ms_per_step": pyr_ms[0] / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": w["name"], "rows": rows, "cols": cols, "levels": L,
                       "pairs_per_gpu": n_pairs, "frames_per_gpu": n_frames, "method": "PHOTO_DEPTH",
                       "occlusion": args.occlusion,
                       "sampling": "nearest-neighbour (reference semantics)",
                       "l2": "inputs larger than L2 (pyramids %.1f GB per GPU)" % (
                           (8 + 24) * ctx.rows * ctx.cols * sum(0.25 ** l for l in range(L)) * n_pairs / 1e9),
                       "step": "pyramid build of all frames + batched alignFrames360" + (" + NCCL allgather of results" if world > 1 else ""),
                       "mean_accepted_iters_per_level": [float(x) for x in iters.mean(0)],
                       "mean_passes_per_level": [float(x) for x in passes.mean(0)],
                       "pairs_ok": int((res["status"] == 0).sum()),
                       "host_numa_node_rank0": numa_node},
            "e2e": {"value": e2e_value, "unit": UNIT,
                    "h2d_bytes_per_step": int(n_frames * npx * 5 + n_pairs * 8 * 3),
                    "d2h_bytes_per_step": int(res_bytes), "steps": e2e_steps,
                    "ms_per_step": e2e_ms / e2e_steps},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "k_pass<PHOTO_DEPTH, WITH_H> (warp/residual/normal-equation pixel passes: the fused launches and the speculative error-only ones, 32 B per source pixel each)",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "peak_source": peak_src,
                         "peak_nominal": 8000.0, "frac_nominal": achieved / 8000.0,   # north_star's "~8 TB/s" (report both)
                         "traffic": (t_ratio * pass_bytes / max(pass_launches, 1)) if t_ratio else None,
                         "traffic_source": t_src,
                         "alg_bytes_per_launch": pass_bytes / max(pass_launches, 1),
                         "avg_launch_ms": pass_ms / max(pass_launches, 1), "launches": pass_launches,
                         "kernel_share_of_step": pass_ms / dev_ms if dev_ms else None},
        }
        if world == 1 and not args.no_cpu_baseline:
            n_cpu = CPU_SAMPLE[args.workload]
            os.sched_setaffinity(0, full_affinity)                     # the CPU baseline gets every host core
            cpu_reference_run(w, 2)                                    # warm-up (library build, page-in)
            v, cores, dt = cpu_reference_run(w, n_cpu)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": "%d pairs of the same workload (frame build + alignFrames360, "
                                              "FAITHFUL accumulation, glibc math, OpenMP), %.1f s" % (n_cpu, dt),
                                    "compiled_reference": compiled_reference_run(w)}
        emit(line)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


_REAL_STDOUT = None


def emit(line):
    """The ONE JSON line of the contract, on the process's original stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is not None:
        os.write(_REAL_STDOUT, data)
    else:
        sys.stdout.write(data.decode()); sys.stdout.flush()


def bind_to_gpu_numa_node(local_rank):
    """Run this rank (and allocate its pinned staging memory) on the NUMA node its GPU hangs off:
    with 8 ranks uploading 10.7 GB per step each, remote-node pinned buffers halve the H2D rate.
    Best effort: silently does nothing when sysfs / NVML do not expose the topology."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
        if len(bus.split(":")[0]) == 8:
            bus = bus[4:]
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return node
    except Exception as e:                      # noqa: BLE001 -- best effort, but say why on stderr
        print(f"bench.py: NUMA binding skipped for local rank {local_rank}: {e!r}", file=sys.stderr)
    return None


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
