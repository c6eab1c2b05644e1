#!/usr/bin/env python
"""bench.py -- throughput of the character-collision hot path on N B200s of one node.

  python bench.py --gpus N --steps K --warmup W            our arm (CUDA, through the C-ABI of include/prtc.h)
  python bench.py --impl reference --gpus N --steps K ...  the reference's CPU implementation on the host cores

A "step" is one frame: one call of PhysicsSystem::updateCharacters (reference physics_system.cpp:245-318) over
the whole batch of capsules = `substeps` swept broadphase queries + narrowphase per capsule.  Workload = C4 of
BASELINE.json (1M capsules x 10M-triangle terrain): ONE batch of 1M capsules in
Morton order, rank r of N steps the contiguous slice r against replicated static geometry -- capsules are independent
units in the static pass, so there is no data-path collective and `scaling` is "strong" (config 4 as BASELINE.json words
it); the state CRC of all capsules is printed and must not depend on N.  Short extra runs on the same line: a crowd that
keeps moving, weak scaling (a full batch per GPU) and config 5 (256k capsules with the dynamic-vs-dynamic pass and its
record exchange).  Prints ONE JSON line (rank 0).
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
    #        terrain N, block x, block z, capsules (GLOBAL count with --scaling strong, per GPU with --scaling weak)
    "C4": (2237, 2, 2, 1000000),
    "C3": (708, 4, 2, 64000),
    "C5": (708, 2, 2, 256000),      # dynamic-vs-dynamic pass on
    "C2": (224, 1, 1, 1000),
}
METRIC = "ellipsoid (capsule) sweep queries/sec"
UNIT = "queries/s"


def parse_args(argv=None):
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
    ap.add_argument("--no-extras", action="store_true", help="skip the short extra runs (moving crowd, weak scaling, config 5)")
    ap.add_argument("--chain", type=int, default=1,
                    help="1: PRTC_OPT_CHAIN_FRAMES for the device-resident leg (consecutive frames overlap on the device); 0: off")
    ap.add_argument("--trace", default="", help="directory for a per-kernel duration table of every rank (diagnostic)")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="strong: ONE global batch split over the GPUs (BASELINE config 4 as written); weak: a full batch per GPU")
    ap.add_argument("--opt", action="append", default=[], metavar="NAME=VALUE",
                    help="prtc_set_option before the run, e.g. --opt POOL_KERNEL=0 (A/B experiments)")
    return ap.parse_args(argv)


def workload(args):
    """(terrain N, block x, block z, GLOBAL capsule count) of the run"""
    n_terrain, bx, bz, n_caps = WORKLOADS[args.workload]
    if args.capsules:
        n_caps = args.capsules
    if args.terrain_n:
        n_terrain = args.terrain_n
    return n_terrain, bx, bz, n_caps


def make_capsules(n_caps, n_terrain, rank):
    pos, vel = scenes.capsules(n_caps, n_terrain, seed=12345 + rank)
    order = scenes.morton_order(pos, n_terrain)         # input preparation: spatially coherent warps
    return np.ascontiguousarray(pos[order]), np.ascontiguousarray(vel[order])


def describe_config(args, name, n, n_global, strong, n_tris, n_terrain, bx, bz, world):
    """the `config` object of the JSON line (both arms print the same one)"""
    dynamic = name == "C5"
    return {
        "workload": "%s: %d capsules %s x %d-triangle terrain (%dx%d-cell mesh colliders), 4 substeps/frame, %s, %s" % (
            name, n_global if strong else n,
            "in ONE Morton-ordered batch, contiguous slices of %d per GPU" % n if strong else "per GPU",
            n_tris, bx, bz,
            "static + dynamic-vs-dynamic pass (JACOBI, records of all characters exchanged per frame)" if dynamic else "static pass",
            "candidates in ascending collider id, tree built on the device (PRTC_ORDER_BY_ID; the roofline's node tests are "
            "those of that tree, not of the reference's)" if args.order == "by_id" else "canonical order"),
        "capsules_per_gpu": n, "global_capsules": n_global, "triangles": n_tris, "terrain_n": n_terrain,
        "substeps": 4, "dt": DT, "abs_mode": "int",
        "parallelism": "capsules sharded x%d (%s), static geometry replicated" % (world, "one global batch" if strong else "a batch per GPU"),
        "l2_policy": "inputs larger than L2: nodes+triangles+capsule state touched per step exceed 126 MB at C4; no explicit flush",
        "seed": 12345,
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
    """The reference's CPU implementation of the path on the host cores (rank 0 only): the L1 port with all host threads on
    a BOUNDED SAMPLE of the workload's capsules (every k-th in Morton order) against the full terrain -- a rate, not the
    whole batch; `reference_fraction` says how much of the batch one step covers."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = max(1, args.gpus)
    strong = args.scaling == "strong"
    n_terrain, bx, bz, n_global = workload(args)
    xyz, mf, mc = scenes.terrain(n_terrain, bx, bz)
    sample = min(args.cpu_sample, n_global)
    pos, vel = make_capsules(n_global, n_terrain, 0)
    # bounded sample: a contiguous Morton range would be one patch of terrain; take every k-th capsule instead
    sel = np.linspace(0, n_global - 1, sample).astype(np.int64)
    threads = os.cpu_count() or 1
    r = cpu_run(xyz, mf, mc, pos[sel], vel[sel], args.steps, args.warmup, threads)
    value = r["substeps"] / r["seconds"]
    verbatim = reference_verbatim()
    sample_txt = "%d of the workload's %d capsules (every %d-th in Morton order) against the full terrain, per step" % (
        sample, n_global, max(1, n_global // sample))
    n_rank = n_global // world if strong else n_global
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": r["steps"],
        "warmup": args.warmup, "ms_per_step": 1e3 * r["seconds"] / max(1, r["steps"]), "higher_is_better": True,
        "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": describe_config(args, args.workload, n_rank, n_global if strong else n_global * world, strong, len(xyz) // 3, n_terrain, bx, bz, world),
        "reference_fraction": sample / float(n_global),
        "reference_sample": sample_txt,
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
class Dist:
    """barrier / reductions over the ranks (NCCL through torch.distributed; plumbing only)"""

    def __init__(self, torch, dev, world):
        self.torch, self.dev, self.world, self.d = torch, dev, world, None
        if world > 1:
            import torch.distributed as dist
            self.d = dist

    def barrier(self):
        if self.d is not None:
            self.d.barrier()
        self.torch.cuda.synchronize()

    def max(self, x):
        if self.d is None:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.d.all_reduce(t, op=self.d.ReduceOp.MAX)
        return float(t.item())

    def per_rank(self, x):
        """x of every rank, in rank order"""
        if self.d is None:
            return [x]
        t = self.torch.zeros(self.world, dtype=self.torch.float64, device=self.dev)
        t[self.d.get_rank()] = x
        self.d.all_reduce(t, op=self.d.ReduceOp.SUM)
        return [float(v) for v in t.tolist()]

    def sum(self, vals):
        if self.d is None:
            return list(vals)
        t = self.torch.tensor(list(vals), dtype=self.torch.float64, device=self.dev)
        self.d.all_reduce(t, op=self.d.ReduceOp.SUM)
        return [float(v) for v in t.tolist()]


def state_crc(torch, sharding, d_tf, d_ph, n):
    """CRC32 of the state of ALL characters in global order: must not depend on the number of ranks"""
    import zlib
    mine = torch.cat([d_tf.view(n, 40), d_ph.view(n, 44)], dim=1).contiguous()
    return zlib.crc32(sharding.gather_rows(mine).cpu().numpy().tobytes())


class Workload:
    """One context + one batch of characters on this rank.
    strong=True : ONE global batch in Morton order, rank r owns the contiguous slice r (BASELINE config 4 as written);
    strong=False: every rank has its own batch of the full size (weak scaling)."""

    def __init__(self, args, name, torch, capi, sharding, local, rank, world, strong, n_override=0, moving=False, ctx=None, geometry=None):
        self.args, self.name, self.torch, self.capi, self.sharding = args, name, torch, capi, sharding
        self.rank, self.world, self.local, self.strong = rank, world, local, strong
        n_terrain, bx, bz, n_global = WORKLOADS[name]
        if args.terrain_n and name == args.workload:
            n_terrain = args.terrain_n
        if n_override:
            n_global = n_override
        self.n_terrain, self.bx, self.bz = n_terrain, bx, bz
        self.dynamic = name == "C5"
        if strong:
            gpos, gvel = make_capsules(n_global, n_terrain, 0)
            self.first, self.n = sharding.partition(n_global, world, rank)
            # diagnostic (one GPU): step the slice rank 0 of a W-rank job would own, next to the frozen records of the others
            self.fake = int(os.environ.get("PRTC_BENCH_FAKE_WORLD", "0")) if world == 1 else 0
            if self.fake:
                self.first, self.n = sharding.partition(n_global, self.fake, 0)
                self.gpos, self.gvel = gpos, gvel
            pos, vel = gpos[self.first:self.first + self.n], gvel[self.first:self.first + self.n]
            self.n_global = n_global
        else:
            pos, vel = make_capsules(n_global, n_terrain, rank)
            self.first, self.n, self.n_global = 0, n_global, n_global * world
            self.fake = 0
        self.pos, self.vel = pos, vel
        dev = torch.device("cuda", local)
        self.t_commit = None
        if ctx is None:
            self.xyz, self.mf, self.mc = geometry if geometry is not None else scenes.terrain(n_terrain, bx, bz)
            ctx = capi.PhysicsContext(device=local, dyn_mode=capi.DYN_JACOBI if self.dynamic else capi.DYN_OFF,
                                      order_mode=capi.ORDER_BY_ID if args.order == "by_id" else capi.ORDER_CANONICAL)
            for kv in args.opt:
                oname, val = kv.split("=")
                ctx.set_option(getattr(capi, "OPT_" + oname), int(val))
            if args.chain and not self.dynamic:
                ctx.set_option(capi.OPT_CHAIN_FRAMES, 1)
            if self.dynamic:
                ctx.dyn_configure(self.first, self.n_global if strong else self.n)
                # the records of the dynamic pass travel by peer stores from the kernel that computes them (CUDA IPC handles
                # swapped once); PRTC_BENCH_EXCHANGE=nccl keeps the all-gather of round 1 for A/B
                self.ipc = False
                if strong and world > 1 and os.environ.get("PRTC_BENCH_EXCHANGE", "ipc") == "ipc":
                    self.ipc = sharding.ipc_connect(ctx, rank, world)
            ctx.add_model_collider(self.xyz, self.mf, self.mc)
            self.cid = ctx.add_capsule_collider()          # one shared collider shape (h=0.5 r=0.5 off=(0,0.5,0))
            t0 = time.perf_counter()
            ctx.commit()
            self.t_commit = time.perf_counter() - t0
        else:
            self.cid = 0
        self.ctx = ctx
        self.n_tris = ctx.tree_info()["triangles"]
        self.tf0 = capi.make_transforms(pos)
        move = None
        if moving:
            # a crowd that keeps walking: movementVector of 0.1 per frame (half the engine's run speed, character.cpp:99)
            # in a per-character direction, so the per-slot leaf lists keep being refilled
            if strong:                         # one global crowd: the state CRC must not depend on the number of ranks
                ang = np.random.RandomState(777).uniform(-np.pi, np.pi, self.n_global)[self.first:self.first + self.n]
            else:
                ang = np.random.RandomState(777 + rank).uniform(-np.pi, np.pi, self.n)
            move = np.stack([0.1 * np.cos(ang), np.zeros(self.n), 0.1 * np.sin(ang)], axis=1).astype(np.float32)
        self.ph0 = capi.make_physics(vel, movement=move, collider=np.full(self.n, self.cid, np.uint32))
        self.h_tf = torch.empty(self.n * 40, dtype=torch.uint8).pin_memory()
        self.h_ph = torch.empty(self.n * 44, dtype=torch.uint8).pin_memory()
        self.tf_np = self.h_tf.numpy().view(capi.TRANSFORM_DTYPE)
        self.ph_np = self.h_ph.numpy().view(capi.PHYSICS_DTYPE)
        self.d_tf = torch.empty(self.n * 40, dtype=torch.uint8, device=dev)
        self.d_ph = torch.empty(self.n * 44, dtype=torch.uint8, device=dev)
        self.stream = torch.cuda.current_stream()
        ctx.set_stream(self.stream.cuda_stream)
        self._rec = None
        if strong and self.fake and self.dynamic:
            full_tf = torch.from_numpy(capi.make_transforms(self.gpos).view(np.uint8).reshape(-1)).to(dev)
            full_ph = torch.from_numpy(capi.make_physics(self.gvel, collider=np.full(self.n_global, self.cid, np.uint32)).view(np.uint8).reshape(-1)).to(dev)
            self.full_tf, self.full_ph = full_tf, full_ph
            self.d_tf, self.d_ph = full_tf[:self.n * 40], full_ph[:self.n * 44]

    def reset(self):
        self.tf_np[:] = self.tf0
        self.ph_np[:] = self.ph0
        self.d_tf.copy_(self.h_tf)
        self.d_ph.copy_(self.h_ph)

    def device_step(self):
        if self.dynamic and self.strong and self.fake:
            self.ctx.dyn_prepare(DT, self.n_global, self.full_tf.data_ptr(), self.full_ph.data_ptr())
        if self.dynamic and self.world > 1 and self.strong and not getattr(self, "ipc", False):
            # the one exchange of the path, collective flavour: the records of every character, all-gathered in place over NVLink
            self.ctx.dyn_prepare(DT, self.n, self.d_tf.data_ptr(), self.d_ph.data_ptr())
            if self._rec is None:                          # the record array never moves: wrap it once
                self._rec = self.sharding.records_tensor(self.ctx)
            self.sharding.allgather_records(self._rec[0], self._rec[1], self.n)
        self.ctx.update_characters_device(DT, self.n, self.d_tf.data_ptr(), self.d_ph.data_ptr())

    def timed(self, dist, steps, warmup, sample_clocks=False, per_frame_events=False):
        """W untimed + K timed frames, device-resident state; CUDA events on the launching stream, max over ranks"""
        torch = self.torch
        self.reset()
        for _ in range(warmup):
            self.device_step()
        torch.cuda.synchronize()
        self.ctx.counters(reset=True)
        sampler = ClockSampler(self.local) if sample_clocks else None
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        dist.barrier()
        if sampler:
            sampler.start()
        # (chained frames overlap on the device: an event between two of them would serialise them again, so only the ends are marked)
        chained = bool(self.args.chain) and not self.dynamic and not per_frame_events
        ev[0].record(self.stream)
        for k in range(steps):
            self.device_step()
            if not chained or k == steps - 1:
                ev[k + 1].record(self.stream)
        dist.barrier()
        clocks = sampler.stop() if sampler else None
        total_ms = ev[0].elapsed_time(ev[steps])
        per = [total_ms / steps] * steps if chained else [ev[k].elapsed_time(ev[k + 1]) for k in range(steps)]
        cnt = self.ctx.counters(reset=True)
        el = dist.max(total_ms * 1e-3)
        self.rank_ms = [v / steps for v in dist.per_rank(total_ms)]
        sums = dist.sum([cnt["substeps"], cnt["tri_tests"], cnt["node_tests"]])
        return {"el": el, "total_ms": total_ms, "per": per, "cnt": cnt, "clocks": clocks, "substeps": sums[0], "tri_tests": sums[1],
                "node_tests": sums[2]}

    def trace(self, steps, path):
        """kernel-by-kernel durations of `steps` more frames (CUPTI through torch.profiler; a diagnostic, never a bench value)"""
        from torch.profiler import profile, ProfilerActivity
        torch = self.torch
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(steps):
                self.device_step()
            torch.cuda.synchronize()
        rows = {}
        t_first, t_last = None, 0
        for e in prof.events():
            if getattr(e.device_type, "name", "") != "CUDA":
                continue
            r = rows.setdefault(e.name, [0, 0.0])
            r[0] += 1
            r[1] += float(getattr(e, "device_time", None) or getattr(e, "cuda_time", 0.0) or (e.time_range.end - e.time_range.start))
            t0 = e.time_range.start
            t_first = t0 if t_first is None else min(t_first, t0)
            t_last = max(t_last, e.time_range.end)
        out = {"rank": self.rank, "steps": steps, "span_us_per_step": (t_last - (t_first or 0)) / steps,
               "kernels": {k: {"launches_per_step": v[0] / steps, "us_per_step": v[1] / steps} for k, v in
                           sorted(rows.items(), key=lambda kv: -kv[1][1])}}
        os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
        with open(path, "w") as f:
            json.dump(out, f, indent=1)

    def crc(self):
        if self.strong and not self.fake:
            return state_crc(self.torch, self.sharding, self.d_tf, self.d_ph, self.n)
        return None

    def describe(self):
        return describe_config(self.args, self.name, self.n, self.n_global, self.strong, self.n_tris, self.n_terrain, self.bx, self.bz, self.world)


def run_ours(args):
    import torch
    from prototype2_b200 import capi, sharding

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the collision path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as tdist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        tdist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    dist = Dist(torch, dev, world)
    strong = args.scaling == "strong"

    w = Workload(args, args.workload, torch, capi, sharding, local, rank, world, strong, n_override=args.capsules)
    ctx, n_caps = w.ctx, w.n
    info = ctx.tree_info()
    cfg, commit_s = w.describe(), w.t_commit

    # ---- device-resident leg: `value` --------------------------------------------------------------------
    r = w.timed(dist, args.steps, args.warmup, sample_clocks=True)
    cnt, el, total_ms, kernel_ms = r["cnt"], r["el"], r["total_ms"], r["per"]
    value = r["substeps"] / el
    crc = w.crc()
    rank_ms = w.rank_ms
    unchained = None
    if args.chain and not w.dynamic:
        # the same frames one launch after the other (option off), with an event per frame
        ctx.set_option(capi.OPT_CHAIN_FRAMES, 0)
        ru = w.timed(dist, args.steps, args.warmup, per_frame_events=True)
        ctx.set_option(capi.OPT_CHAIN_FRAMES, 1)
        assert w.crc() == crc, "chained and unchained frames must leave the same state"
        unchained = {"value": ru["substeps"] / ru["el"], "ms_per_step": 1e3 * ru["el"] / args.steps,
                     "ms_per_step_median": float(np.median(ru["per"])), "ms_per_step_min": float(np.min(ru["per"]))}
    if args.trace:
        w.trace(args.steps, os.path.join(args.trace, "trace_%s_n%d_rank%d.json" % (args.workload, world, rank)))

    # Numerators of the roofline: the REFERENCE ALGORITHM's counts over the same K frames.  The timed run uses the
    # frame cache (fewer node tests than DynamicAABBTree::query performs), so the same frames are replayed untimed
    # with PRTC_OPT_FRAME_CACHE=0, where the device walks the tree per substep and counts exactly what the
    # reference would test; substeps and triangle tests must come out identical in both runs.
    if w.dynamic:                                    # stored boxes carry over between frames: replay on a fresh context
        w2 = Workload(args, args.workload, torch, capi, sharding, local, rank, world, strong, n_override=args.capsules,
                      geometry=(w.xyz, w.mf, w.mc))
    else:
        w2 = w
    w2.ctx.set_option(capi.OPT_FRAME_CACHE, 0)
    w2.reset()
    for _ in range(args.warmup):
        w2.device_step()
    torch.cuda.synchronize()
    w2.ctx.counters(reset=True)
    for _ in range(args.steps):
        w2.device_step()
    torch.cuda.synchronize()
    ref_cnt = w2.ctx.counters(reset=True)
    w2.ctx.set_option(capi.OPT_FRAME_CACHE, 1)
    if w.dynamic:
        del w2
    assert ref_cnt["substeps"] == cnt["substeps"] and ref_cnt["tri_tests"] == cnt["tri_tests"], (ref_cnt, cnt)
    gpu_node_tests = cnt["node_tests"]
    (node_tests,) = dist.sum([ref_cnt["node_tests"]])

    # roofline of the dominant kernel, this rank.  `achieved` is what the contract defines: ALGORITHMIC bytes (SURVEY.md 8d,
    # with the node tests the REFERENCE's traversal performs) / measured duration.  Two more fractions beside it say what
    # the number is made of: the same formula with the node tests the DEVICE performed (its leaf lists are cached across
    # substeps and frames), and the DRAM bytes ncu measured for one launch -- the working set is cache resident and the
    # kernel is bound by instruction issue and latency, not by HBM.
    def alg(nodes):
        return 96.0 * cnt["substeps"] + 32.0 * nodes + 36.0 * cnt["tri_tests"] + 64.0 * n_caps * args.steps
    alg_bytes = alg(ref_cnt["node_tests"])
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
    kernel = ("k_step_warp<MODE_STATIC>" if n_caps <= 2368 and not w.dynamic else
              "k_step_pool<false, %s, %s>" % ("MODE_JACOBI" if w.dynamic else "MODE_STATIC",
                                               "true" if (args.chain and not w.dynamic and int(cnt["launches"]) == args.steps) else "false"))   # what launch_step picks
    avg_ms = float(np.mean(kernel_ms))
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "kernel": kernel, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes / args.steps,
                "avg_launch_ms": avg_ms, "formula": "96*S + 32*V + 36*T + 64*capsules (SURVEY.md 8d)",
                "frac_device_node_tests": alg(gpu_node_tests) / (total_ms * 1e-3) / 1e9 / peak,
                "frac_dram": (traffic / (avg_ms * 1e-3) / 1e9 / peak) if traffic else None,
                "note": "frac = reference-equivalent algorithmic bytes (the contract's definition); frac_device_node_tests = same "
                        "formula with the node tests the device performed; frac_dram = measured DRAM bytes of one launch / time / peak: "
                        "the kernel is issue- and latency-bound, not HBM-bound",
                "per_query": {"V_q": ref_cnt["node_tests"] / max(1, cnt["substeps"]), "T_q": cnt["tri_tests"] / max(1, cnt["substeps"]),
                              "gpu_node_tests_per_query": gpu_node_tests / max(1, cnt["substeps"]),
                              "exact_tests_per_query": cnt["exact_tests"] / max(1, cnt["substeps"]),
                              "contacts_per_query": cnt["contacts"] / max(1, cnt["substeps"])}}

    # ---- end-to-end legs: through the public host API, transfers inside the timed region, wall clock ----------
    # e2e      prtc_resident_step: the records stay on the device (as the engine's caller only gathers / scatters positions
    #          around updateCharacters, character_system.cpp:55-78); every step the kernels read this frame's movement input
    #          (12 B per character) from pinned host memory and write position + isGrounded (16 B) back to pinned host
    #          memory, in place over the bus; the host reads the result array after every step.
    # e2e_aos  prtc_step_characters: the full AoS records (84 B per character each way) from and to pinned host buffers.
    e2e = e2e_aos = None
    if not args.no_e2e and not (w.dynamic and world > 1):       # (config 5 over several ranks has its exchange hook in the device leg)
        own = torch.cuda.Stream()
        ctx.set_stream(own.cuda_stream)
        ctx.resident_begin(w.tf0, w.ph0)
        move, result = ctx.resident_io()
        for _ in range(min(2, args.warmup)):
            ctx.resident_step(DT)
        ctx.resident_begin(w.tf0, w.ph0)                         # same start state as the device leg
        move, result = ctx.resident_io()
        ctx.counters(reset=True)
        dist.barrier()
        t0 = time.perf_counter()
        seen = 0.0
        for _ in range(args.steps):
            ctx.resident_step(DT)
            seen += float(result[n_caps // 2, 1]) + float(result[n_caps - 1, 3])      # the caller scatters positions back: read the result
        el_r = time.perf_counter() - t0
        dist.barrier()
        el_r = dist.max(el_r)
        c3 = ctx.counters(reset=True)
        (sub_r,) = dist.sum([c3["substeps"]])
        e2e = {"value": sub_r / el_r, "unit": UNIT, "h2d_bytes_per_step": n_caps * 12, "d2h_bytes_per_step": n_caps * 16,
               "ms_per_step": 1e3 * el_r / args.steps,
               "api": "prtc_resident_step (records resident on the device; movement in / position + isGrounded out through mapped "
                      "pinned host arrays, read and written by the kernels in place, host reads the result every step)",
               "result_checksum": seen}
        ctx.resident_end()
        w.tf_np[:] = w.tf0
        w.ph_np[:] = w.ph0
        for _ in range(min(2, args.warmup)):
            ctx.update_characters(DT, w.tf_np, w.ph_np)
        w.tf_np[:] = w.tf0
        w.ph_np[:] = w.ph0
        ctx.counters(reset=True)
        dist.barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            ctx.update_characters(DT, w.tf_np, w.ph_np)
        torch.cuda.synchronize()
        el_e = time.perf_counter() - t0
        dist.barrier()
        el_e = dist.max(el_e)
        c2 = ctx.counters(reset=True)
        (sub_e,) = dist.sum([c2["substeps"]])
        e2e_aos = {"value": sub_e / el_e, "unit": UNIT, "h2d_bytes_per_step": n_caps * 84, "d2h_bytes_per_step": n_caps * 84,
                   "ms_per_step": 1e3 * el_e / args.steps, "api": "prtc_step_characters (host AoS buffers, pinned)"}
        ctx.set_stream(w.stream.cuda_stream)

    # ---- extra measurements on the same line (short runs) -----------------------------------------------------
    extras = {}
    if args.workload == "C4" and not args.capsules and not args.no_extras:
        steps_x = max(3, min(args.steps, 10))
        # a crowd that keeps moving (movement input != 0): the per-slot leaf lists are refilled all the time
        wm = Workload(args, "C4", torch, capi, sharding, local, rank, world, strong, moving=True, ctx=ctx)
        rm = wm.timed(dist, steps_x, 3)
        extras["moving_crowd"] = {"ms_per_step": 1e3 * rm["el"] / steps_x, "value": rm["substeps"] / rm["el"], "unit": UNIT,
                                  "movement": "0.1 per frame in a per-character direction", "state_crc32": wm.crc()}
        del wm
        if world > 1 and strong:
            # weak scaling beside the strong number: every rank its own 1 M capsules, no exchange -- trivially linear
            ww = Workload(args, "C4", torch, capi, sharding, local, rank, world, False, ctx=ctx)
            rw = ww.timed(dist, steps_x, 3)
            extras["weak"] = {"ms_per_step": 1e3 * rw["el"] / steps_x, "value": rw["substeps"] / rw["el"], "unit": UNIT,
                              "capsules_per_gpu": ww.n, "scaling": "weak"}
            del ww
        # config 5: 256 k moving capsules with the dynamic-vs-dynamic pass, one global batch over the ranks
        del w
        torch.cuda.empty_cache()
        w5 = Workload(args, "C5", torch, capi, sharding, local, rank, world, True)
        r5 = w5.timed(dist, max(steps_x, 10), 3)
        k5 = max(steps_x, 10)
        if args.trace:
            w5.trace(k5, os.path.join(args.trace, "trace_C5_n%d_rank%d.json" % (world, rank)))
        extras["c5"] = {"workload": w5.describe()["workload"], "ms_per_step": 1e3 * r5["el"] / k5, "value": r5["substeps"] / r5["el"],
                        "unit": UNIT, "state_crc32": w5.crc(), "scaling": "strong", "ms_per_step_by_rank": w5.rank_ms,
                        "exchange_bytes_per_step": (w5.n_global * 64) if world > 1 else 0,
                        "exchange": ("none (one rank)" if world == 1 else
                                     "peer stores of the 64-byte records from k_dyn_prep into every rank's array (CUDA IPC), arrival words in device memory"
                                     if getattr(w5, "ipc", False) else "ncclAllGather of 64-byte records, in place")}
        del w5

    # ---- CPU baseline beside it (rank 0, N=1 only) ----------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        n_terrain, bx, bz, _ = WORKLOADS[args.workload]
        xyz, mf, mc = scenes.terrain(args.terrain_n or n_terrain, bx, bz)
        pos, vel = make_capsules(n_caps, args.terrain_n or n_terrain, 0)
        sample = min(args.cpu_sample, n_caps)
        sel = np.linspace(0, n_caps - 1, sample).astype(np.int64)
        threads = os.cpu_count() or 1
        rc = cpu_run(xyz, mf, mc, pos[sel], vel[sel], 1000, 1, threads, seconds_cap=args.cpu_seconds)
        r1 = cpu_run(xyz, mf, mc, pos[sel[::8]], vel[sel[::8]], 1000, 1, 1, seconds_cap=4.0)
        cpu = {"value": rc["substeps"] / rc["seconds"], "unit": UNIT, "cores": threads, "kind": "port", "host": host_info(),
               "single_thread_value": r1["substeps"] / r1["seconds"],
               "sample": "%d of %d capsules (every %d-th in Morton order) x %d frames against the full terrain, %.1f s" % (
                   sample, n_caps, max(1, n_caps // sample), rc["steps"], rc["seconds"]),
               "tri_tests_per_s": rc["tri_tests"] / rc["seconds"]}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * el / args.steps, "ms_per_step_median": float(np.median(kernel_ms)), "ms_per_step_min": float(np.min(kernel_ms)),
            "ms_per_step_by_rank": rank_ms,
            "frames": ("chained: consecutive frames overlap on the device (PRTC_OPT_CHAIN_FRAMES; per-claim ordering, same bytes)"
                       if (args.chain and unchained is not None and int(cnt["launches"]) == args.steps) else
                       "one launch after the other (the library chains batches of up to 8 rounds of claims; this one has more)"
                       if (args.chain and unchained is not None) else "one launch after the other"),
            "unchained": unchained,
            "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": cfg,
            "tri_tests_per_s": r["tri_tests"] / el, "node_tests_per_s": node_tests / el,
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "e2e_aos": e2e_aos, "gpu_launches": int(cnt["launches"]), "clocks": r["clocks"],
            "setup": {"commit_s": commit_s, "nodes": info["nodes"], "leaves": info["leaves"]},
        }
        if crc is not None:
            line["state_crc32"] = crc
        line.update(extras)
        emit(line)
    if world > 1:
        import torch.distributed as tdist
        tdist.destroy_process_group()



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
