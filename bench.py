#!/usr/bin/env python
"""bench.py -- throughput of the character-collision hot path on N B200s of one node.

  python bench.py --gpus N --steps K --warmup W            our arm (CUDA, through the C-ABI of include/prtc.h)
  python bench.py --impl reference --gpus N --steps K ...  the reference's CPU implementation on the host cores

A "step" is one frame: one call of PhysicsSystem::updateCharacters (reference physics_system.cpp:245-318) over
the whole batch of capsules = `substeps` swept broadphase queries + narrowphase per capsule.  Workload = C4 of
BASELINE.json (1M capsules x 10M-triangle terrain) on every GPU: capsules are
independent units, so ranks hold disjoint capsule sets against replicated static geometry (weak scaling, no
data-path collective).  Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from prototype2_b200 import scenes  # noqa: E402

DT = 1.0 / 30.0
WORKLOADS = {
    #        terrain N, block x, block z, capsules per GPU
    "C4": (2237, 2, 2, 1000000),
    "C3": (708, 4, 2, 64000),
    "C5": (708, 2, 2, 256000),      # GLOBAL capsule count, partitioned over the ranks; dynamic-vs-dynamic pass on
    "C2": (224, 1, 1, 1000),
}
METRIC = "ellipsoid (capsule) sweep queries/sec"
UNIT = "queries/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C4", choices=sorted(WORKLOADS))
    ap.add_argument("--order", default="canonical", choices=["canonical", "by_id"],
                    help="candidate order: canonical = the reference-built tree (headline), by_id = PRTC_ORDER_BY_ID, tree built on the device")
    ap.add_argument("--capsules", type=int, default=0, help="override capsules per GPU")
    ap.add_argument("--terrain-n", type=int, default=0, help="override terrain size")
    ap.add_argument("--cpu-sample", type=int, default=16384, help="capsules in the bounded CPU sample")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU time of the cpu_baseline leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--opt", action="append", default=[], metavar="NAME=VALUE",
                    help="prtc_set_option before the run, e.g. --opt POOL_KERNEL=0 (A/B experiments)")
    return ap.parse_args()


def workload(args):
    n_terrain, bx, bz, n_caps = WORKLOADS[args.workload]
    if args.capsules:
        n_caps = args.capsules
    if args.workload == "C5":                     # strong scaling: the global batch is split over the ranks
        n_caps //= max(1, int(os.environ.get("WORLD_SIZE", "1")))
    if args.terrain_n:
        n_terrain = args.terrain_n
    return n_terrain, bx, bz, n_caps


def make_capsules(n_caps, n_terrain, rank):
    pos, vel = scenes.capsules(n_caps, n_terrain, seed=12345 + rank)
    order = scenes.morton_order(pos, n_terrain)         # input preparation: spatially coherent warps
    return np.ascontiguousarray(pos[order]), np.ascontiguousarray(vel[order])


def describe(args, n_terrain, bx, bz, n_caps, n_tris, world):
    return {
        "workload": "%s: %d capsules/GPU x %d-triangle terrain (%dx%d-cell mesh colliders), 4 substeps/frame, %s, "
                    "%s" % (args.workload, n_caps, n_tris, bx, bz,
                            "static + dynamic-vs-dynamic pass (JACOBI, records all-gathered per frame)" if args.workload == "C5" else "static pass",
                            "candidates in ascending collider id, tree built on the device (PRTC_ORDER_BY_ID; the roofline's node "
                            "tests are those of that tree, not of the reference's)" if args.order == "by_id" else "canonical order"),
        "capsules_per_gpu": n_caps, "global_capsules": n_caps * world, "triangles": n_tris, "terrain_n": n_terrain,
        "substeps": 4, "dt": DT, "abs_mode": "int", "parallelism": "capsules sharded x%d, static geometry replicated" % world,
        "l2_policy": "inputs larger than L2: nodes+triangles+capsule state touched per step exceed 126 MB at C4; "
                     "no explicit flush", "seed": 12345,
    }


# ----------------------------------------------------------------------------------------------------------
# clocks sampling during the timed region (NVML, 20 ms period)
# ----------------------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, device_index):
        self.samples, self.reasons, self.max_mhz, self.ok = [], set(), None, False
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = device_index
            if vis:
                try:
                    phys = int(vis.split(",")[device_index])
                except Exception:
                    phys = device_index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception:
            self.ok = False

    def _loop(self):
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
                 "hw_power_brake": getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80)}
        while not self._stop.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.02)

    def start(self):
        if self.ok:
            self._thread = threading.Thread(target=self._loop, daemon=True)
            self._thread.start()

    def stop(self):
        if self._thread:
            self._stop.set()
            self._thread.join()
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ----------------------------------------------------------------------------------------------------------
# CPU arm: the reference's algorithm on the host cores (oracle port L1; L0 cannot index 1.25M mesh colliders
# with its uint16 tags, SURVEY.md 8c hazard iii)
# ----------------------------------------------------------------------------------------------------------
def cpu_run(xyz, mf, mc, pos, vel, steps, warmup, threads, seconds_cap=None):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oraclelib import L1World
    w = L1World("canonical", dyn="off", abs_mode="int", threads=threads)
    w.add_model(xyz, mf, mc)
    for _ in range(len(pos)):
        w.add_capsule()
    w.set_state(pos=pos, vel=vel)
    for _ in range(warmup):
        w.step(DT)
    w.counters(reset=True)
    t0 = time.perf_counter()
    done = 0
    for _ in range(steps):
        w.step(DT)
        done += 1
        if seconds_cap and time.perf_counter() - t0 > seconds_cap:
            break
    el = time.perf_counter() - t0
    c = w.counters()
    w.close()
    return {"seconds": el, "steps": done, "substeps": c["substeps"], "tri_tests": c["tri_tests"], "node_tests": c["node_tests"]}


def reference_verbatim():
    """The reference's OWN sources (oracle/_ref/libprt_l0.so, compiled unmodified) cannot index C4's 1.25 M mesh
    colliders with their uint16 tags, so they are timed beside the port on the largest shape they hold: C3's terrain
    (62 658 colliders of 16 triangles), 4096 capsules kept apart, 1 thread (the reference is single-threaded)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    try:
        from oraclelib import L0World
        if not L0World.available("plain"):
            return None
        xyz, mf, mc = scenes.terrain(708, 4, 2)
        pos, vel = scenes.capsules_grid(4096, 708, seed=12345)
        w = L0World("plain")
        w.add_model(xyz, mf, mc)
        for _ in range(len(pos)):
            w.add_capsule()
        w.set_state(pos=pos, vel=vel)
        w.step(DT)                       # warm-up frame (first frame re-inserts every capsule leaf)
        frames, t0 = 0, time.perf_counter()
        while frames < 20 and time.perf_counter() - t0 < 8.0:
            w.step(DT)
            frames += 1
        el = time.perf_counter() - t0
        w.close()
        # every capsule executes all 4 substeps here (terrain everywhere below it)
        return {"value": 4096 * 4 * frames / el, "unit": UNIT, "cores": 1, "kind": "reference",
                "sample": "C3 terrain (1 002 528 triangles, 62 658 mesh colliders), 4096 capsules, %d frames, %.1f s; upper "
                          "bound on queries (counts 4 substeps per capsule-frame)" % (frames, el)}
    except Exception as e:             # never let the informational leg break the line
        return {"error": str(e)}


def host_info():
    model = "unknown"
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                model = ln.split(":", 1)[1].strip()
                break
    except OSError:
        pass
    return {"cpu_model": model, "nproc": os.cpu_count(),
            "build": "g++ -std=c++17 -O3 -DNDEBUG -ffp-contract=off (oracle/Makefile), no sanitizers"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_terrain, bx, bz, n_caps = workload(args)
    xyz, mf, mc = scenes.terrain(n_terrain, bx, bz)
    sample = min(args.cpu_sample, n_caps)
    pos, vel = make_capsules(n_caps, n_terrain, 0)
    # bounded sample: a contiguous Morton range would be one patch of terrain; take every k-th capsule instead
    sel = np.linspace(0, n_caps - 1, sample).astype(np.int64)
    threads = os.cpu_count() or 1
    r = cpu_run(xyz, mf, mc, pos[sel], vel[sel], args.steps, args.warmup, threads)
    value = r["substeps"] / r["seconds"]
    verbatim = reference_verbatim()
    sample_txt = "%d of the workload's %d capsules (every %d-th in Morton order) against the full terrain, per step" % (
        sample, n_caps, max(1, n_caps // sample))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": r["steps"],
        "warmup": args.warmup, "ms_per_step": 1e3 * r["seconds"] / max(1, r["steps"]), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": describe(args, n_terrain, bx, bz, n_caps, len(xyz) // 3, max(1, args.gpus)),
        "tri_tests_per_s": r["tri_tests"] / r["seconds"],
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample_txt, "host": host_info(),
                         "note": "oracle/l1_oracle.cpp (restatement pinned bit-identical to the reference's own sources; "
                                 "the verbatim build cannot hold >65534 mesh colliders), std::thread over capsules"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "reference_verbatim": verbatim,
    }
    emit(line)


# ----------------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    from prototype2_b200 import capi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the collision path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    n_terrain, bx, bz, n_caps = workload(args)
    xyz, mf, mc = scenes.terrain(n_terrain, bx, bz)
    n_tris = len(xyz) // 3
    dynamic = args.workload == "C5"
    if dynamic:
        # one global batch in Morton order, contiguous slices per rank (ascending-index neighbour order is global)
        from prototype2_b200 import sharding
        gpos, gvel = make_capsules(n_caps * world, n_terrain, 0)
        first, _ = sharding.partition(n_caps * world, world, rank)
        pos, vel = gpos[first:first + n_caps], gvel[first:first + n_caps]
    else:
        pos, vel = make_capsules(n_caps, n_terrain, rank)

    ctx = capi.PhysicsContext(device=local, dyn_mode=capi.DYN_JACOBI if dynamic else capi.DYN_OFF,
                              order_mode=capi.ORDER_BY_ID if args.order == "by_id" else capi.ORDER_CANONICAL)
    if dynamic:
        ctx.dyn_configure(first, n_caps * world)
    for kv in args.opt:
        name, val = kv.split("=")
        ctx.set_option(getattr(capi, "OPT_" + name), int(val))
    ctx.add_model_collider(xyz, mf, mc)
    cid = ctx.add_capsule_collider()                      # one shared collider shape (h=0.5 r=0.5 off=(0,0.5,0))
    t_commit = time.perf_counter()
    ctx.commit()
    t_commit = time.perf_counter() - t_commit
    info = ctx.tree_info()

    tf0 = capi.make_transforms(pos)
    ph0 = capi.make_physics(vel, collider=np.full(n_caps, cid, np.uint32))
    h_tf = torch.empty(n_caps * 40, dtype=torch.uint8).pin_memory()
    h_ph = torch.empty(n_caps * 44, dtype=torch.uint8).pin_memory()
    tf_np = h_tf.numpy().view(capi.TRANSFORM_DTYPE)
    ph_np = h_ph.numpy().view(capi.PHYSICS_DTYPE)

    def reset_host():
        tf_np[:] = tf0
        ph_np[:] = ph0

    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)
    d_tf = torch.empty(n_caps * 40, dtype=torch.uint8, device=dev)
    d_ph = torch.empty(n_caps * 44, dtype=torch.uint8, device=dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(vals):
        if dist is None:
            return vals
        t = torch.tensor(vals, dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return [float(v) for v in t.tolist()]

    def device_step():
        if dynamic and world > 1:
            # the one exchange of the path: 64-byte records of every character, all-gathered in place over NVLink
            ctx.dyn_prepare(DT, n_caps, d_tf.data_ptr(), d_ph.data_ptr())
            rec, rfirst = sharding.records_tensor(ctx)
            sharding.allgather_records(rec, rfirst, n_caps)
        ctx.update_characters_device(DT, n_caps, d_tf.data_ptr(), d_ph.data_ptr())

    # ---- device-resident leg: `value` --------------------------------------------------------------------
    reset_host()
    d_tf.copy_(h_tf)
    d_ph.copy_(h_ph)
    for _ in range(args.warmup):
        device_step()
    torch.cuda.synchronize()
    ctx.counters(reset=True)
    sampler = ClockSampler(local)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    sampler.start()
    ev[0].record(stream)
    for k in range(args.steps):
        device_step()
        ev[k + 1].record(stream)
    barrier()
    clocks = sampler.stop()
    total_ms = ev[0].elapsed_time(ev[args.steps])
    kernel_ms = [ev[k].elapsed_time(ev[k + 1]) for k in range(args.steps)]
    cnt = ctx.counters(reset=True)
    state_crc = None
    if dynamic:
        # one global batch: the state after W+K frames must not depend on the number of ranks -> checksum over all
        # characters in global order (compare the value printed by an N=1 and an N>1 run)
        import zlib
        mine = torch.cat([d_tf.view(n_caps, 40), d_ph.view(n_caps, 44)], dim=1).contiguous()
        allrows = sharding.gather_rows(mine)
        state_crc = zlib.crc32(allrows.cpu().numpy().tobytes())
    el = max_over_ranks(total_ms * 1e-3)
    substeps, tri_tests, node_tests = sum_over_ranks([cnt["substeps"], cnt["tri_tests"], cnt["node_tests"]])
    value = substeps / el

    # Numerators of the roofline: the REFERENCE ALGORITHM's counts over the same K frames.  The timed run uses the
    # frame cache (fewer node tests than DynamicAABBTree::query performs), so the same frames are replayed untimed
    # with PRTC_OPT_FRAME_CACHE=0, where the device walks the tree per substep and counts exactly what the
    # reference would test; substeps and triangle tests must come out identical in both runs.
    reset_host()
    d_tf.copy_(h_tf)
    d_ph.copy_(h_ph)
    ctx.set_option(capi.OPT_FRAME_CACHE, 0)
    if dynamic:                                      # stored boxes carry over between frames: start the replay from a fresh context state
        ctx2 = capi.PhysicsContext(device=local, dyn_mode=capi.DYN_JACOBI,
                                   order_mode=capi.ORDER_BY_ID if args.order == "by_id" else capi.ORDER_CANONICAL)
        ctx2.add_model_collider(xyz, mf, mc)
        ctx2.add_capsule_collider()
        ctx2.dyn_configure(first, n_caps * world)
        ctx2.set_stream(stream.cuda_stream)
        ctx, ctx_timed = ctx2, ctx
        ctx.set_option(capi.OPT_FRAME_CACHE, 0)
    for _ in range(args.warmup):
        device_step()
    torch.cuda.synchronize()
    ctx.counters(reset=True)
    for _ in range(args.steps):
        device_step()
    torch.cuda.synchronize()
    ref_cnt = ctx.counters(reset=True)
    if dynamic:
        ctx = ctx_timed
    ctx.set_option(capi.OPT_FRAME_CACHE, 1)
    assert ref_cnt["substeps"] == cnt["substeps"] and ref_cnt["tri_tests"] == cnt["tri_tests"], (ref_cnt, cnt)
    gpu_node_tests = cnt["node_tests"]
    cnt = dict(cnt, node_tests=ref_cnt["node_tests"])
    (node_tests,) = sum_over_ranks([cnt["node_tests"]])

    # roofline of the dominant (only) kernel, this rank: algorithmic bytes / measured duration
    alg_bytes = 96.0 * cnt["substeps"] + 32.0 * cnt["node_tests"] + 36.0 * cnt["tri_tests"] + 64.0 * n_caps * args.steps
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    achieved = alg_bytes / (total_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))
            if tj.get("workload") == args.workload and tj.get("capsules") == n_caps:
                traffic = tj.get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "kernel": ("k_step<false, MODE_JACOBI>" if args.workload == "C5" else
                           "k_step_warp<MODE_STATIC>" if n_caps <= 2960 else "k_step_stream<false>"),   # what launch_step picks
                "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes / args.steps,
                "avg_launch_ms": float(np.mean(kernel_ms)), "formula": "96*S + 32*V + 36*T + 64*capsules (SURVEY.md 8d)",
                "per_query": {"V_q": cnt["node_tests"] / max(1, cnt["substeps"]), "T_q": cnt["tri_tests"] / max(1, cnt["substeps"]),
                              "gpu_node_tests_per_query": gpu_node_tests / max(1, cnt["substeps"]),
                              "exact_tests_per_query": cnt["exact_tests"] / max(1, cnt["substeps"]),
                              "contacts_per_query": cnt["contacts"] / max(1, cnt["substeps"])}}

    # ---- end-to-end leg: host buffers through prtc_step_characters (H2D + kernel + D2H every step) ----------
    e2e = None
    if not args.no_e2e and not (dynamic and world > 1):       # the host-buffer entry point has no exchange hook
        own = torch.cuda.Stream()
        ctx.set_stream(own.cuda_stream)
        reset_host()
        for _ in range(min(2, args.warmup)):
            ctx.update_characters(DT, tf_np, ph_np)
        reset_host()
        ctx.counters(reset=True)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            ctx.update_characters(DT, tf_np, ph_np)
        torch.cuda.synchronize()
        el_e = time.perf_counter() - t0
        if dist is not None:
            dist.barrier()
        el_e = max_over_ranks(el_e)
        c2 = ctx.counters(reset=True)
        (sub_e,) = sum_over_ranks([c2["substeps"]])
        e2e = {"value": sub_e / el_e, "unit": UNIT, "h2d_bytes_per_step": n_caps * 84, "d2h_bytes_per_step": n_caps * 84,
               "ms_per_step": 1e3 * el_e / args.steps, "api": "prtc_step_characters (host AoS buffers, pinned)"}
        ctx.set_stream(stream.cuda_stream)

    # ---- CPU baseline beside it (rank 0, N=1 only) ----------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sample = min(args.cpu_sample, n_caps)
        sel = np.linspace(0, n_caps - 1, sample).astype(np.int64)
        threads = os.cpu_count() or 1
        r = cpu_run(xyz, mf, mc, pos[sel], vel[sel], 1000, 1, threads, seconds_cap=args.cpu_seconds)
        r1 = cpu_run(xyz, mf, mc, pos[sel[::8]], vel[sel[::8]], 1000, 1, 1, seconds_cap=4.0)
        cpu = {"value": r["substeps"] / r["seconds"], "unit": UNIT, "cores": threads, "kind": "port", "host": host_info(),
               "single_thread_value": r1["substeps"] / r1["seconds"],
               "sample": "%d of %d capsules (every %d-th in Morton order) x %d frames against the full terrain, %.1f s" % (
                   sample, n_caps, max(1, n_caps // sample), r["steps"], r["seconds"]),
               "tri_tests_per_s": r["tri_tests"] / r["seconds"]}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * el / args.steps, "ms_per_step_median": float(np.median(kernel_ms)), "ms_per_step_min": float(np.min(kernel_ms)),
            "higher_is_better": True, "scaling": "strong" if dynamic else "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": describe(args, n_terrain, bx, bz, n_caps, n_tris, world),
            "tri_tests_per_s": tri_tests / el, "node_tests_per_s": node_tests / el,
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(cnt["launches"]), "clocks": clocks,
            "setup": {"commit_s": t_commit, "nodes": info["nodes"], "leaves": info["leaves"]},
        }
        if state_crc is not None:
            line["state_crc32"] = state_crc
        emit(line)
    if dist is not None:
        dist.destroy_process_group()


_REAL_STDOUT = None


def emit(line):
    """the ONE JSON line goes to the process's real stdout; everything else (NCCL banners, warnings) to stderr"""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    global _REAL_STDOUT
    args = parse_args()
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)                       # libraries that print to fd 1 (e.g. "NCCL version ...") must not pollute the JSON line
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
