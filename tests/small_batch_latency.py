"""Fixed cost of one engine-sized host-buffer step (prtc_step_characters, n = 1): a character far above the shipped
level finds no candidate and leaves the kernel at once, so the call time is launch + synchronisation + the kernel's
accesses to the staged records (+ ctypes); next to it the same character walking on the level (tests/c1_frame_latency.py
has the scripted 300 frames).  Usage on the GPU box:  PYTHONPATH=$PWD python tests/small_batch_latency.py"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H  # noqa: E402
from prototype2_b200 import capi  # noqa: E402

g = np.load(os.path.join(ROOT, "tests", "golden", "c1_level.npz"))
for mode, name in ((capi.ORDER_LITERAL, "literal"), (capi.ORDER_CANONICAL, "canonical")):
    for where, y in (("far above the level (no candidates)", 1000.0), ("standing on the level", None)):
        ctx, tf, ph = H.build_gpu(H.c1_scene(g), order_mode=mode)
        if y is not None:
            tf["position"][0, 1] = np.float32(y)
        ts = []
        for f in range(400):
            if y is not None:
                tf["position"][0, 1] = np.float32(y)
                ph["velocity"][0] = 0.0
            t = time.perf_counter()
            ctx.update_characters(float(g["dt"]), tf, ph)
            ts.append(time.perf_counter() - t)
        ts = np.array(ts[100:]) * 1e3
        print("%-9s %-38s median %.3f ms  p95 %.3f" % (name, where, np.median(ts), np.percentile(ts, 95)))
