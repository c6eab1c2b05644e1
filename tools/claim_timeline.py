#!/usr/bin/env python
"""Diagnostic: when every claim (a warp's group of characters) of one k_step_pool launch started and ended, and on which SM.
Needs the diagnostic twin of the library (`make -C prototype2_b200/csrc timeline` -> libprtc_timeline.so) and
PRTC_TIMELINE=1; both are arranged below before the library is loaded.  GPU box.

  python tools/claim_timeline.py [--world W] [--workload C4] [--frames 8]
steps rank 0's slice of a W-rank job of the bench workload and prints what the launch's time is made of."""
import argparse
import json
import os
import sys

os.environ["PRTC_TIMELINE"] = "1"
os.environ.setdefault("PRTC_LIB", "libprtc_timeline.so")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--world", type=int, default=8)
    ap.add_argument("--workload", default="C4")
    ap.add_argument("--frames", type=int, default=8)
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    import torch
    import bench
    from prototype2_b200 import capi, sharding
    if a.world > 1:
        os.environ["PRTC_BENCH_FAKE_WORLD"] = str(a.world)
    args = bench.parse_args(["--workload", a.workload])
    w = bench.Workload(args, a.workload, torch, capi, sharding, 0, 0, 1, True)
    w.reset()
    for _ in range(a.frames):
        w.device_step()
    torch.cuda.synchronize()
    tl = w.ctx.debug_timeline().astype(np.int64)
    tl = tl[tl[:, 0] > 0]
    start, end, sm = tl[:, 0], tl[:, 1], tl[:, 2]
    t0, t1 = start.min(), end.max()
    dur = (end - start) * 1e-3
    span = (t1 - t0) * 1e-3
    slots = 148 * 16
    busy = dur.sum()
    rep = {"workload": a.workload, "world": a.world, "characters": w.n, "claims": int(len(tl)), "span_us": span,
           "claim_us": {k: float(np.percentile(dur, q)) for k, q in (("min", 0), ("p10", 10), ("p50", 50), ("p90", 90), ("p99", 99), ("max", 100))},
           "claim_us_mean": float(dur.mean()), "slot_busy_fraction": float(busy / (span * slots)),
           "last_start_us": float((start.max() - t0) * 1e-3),
           "first_end_us": float((end.min() - t0) * 1e-3)}
    # how many claims are in flight over time (20 bins)
    edges = np.linspace(t0, t1, 21)
    mid = 0.5 * (edges[:-1] + edges[1:])
    rep["in_flight"] = [int(((start <= m) & (end > m)).sum()) for m in mid]
    # duration against start time: are late claims shorter (less contention)?
    order = np.argsort(start)
    q = len(order) // 4
    rep["claim_us_mean_by_start_quarter"] = [float(dur[order[k * q:(k + 1) * q]].mean()) for k in range(4)]
    # per SM: when its last claim ended
    last = np.array([end[sm == s].max() for s in np.unique(sm)])
    rep["sm_finish_us"] = {"min": float((last.min() - t0) * 1e-3), "p50": float((np.median(last) - t0) * 1e-3), "max": float((last.max() - t0) * 1e-3)}
    print(json.dumps(rep, indent=1))
    if a.out:
        np.save(a.out, tl)


if __name__ == "__main__":
    main()
