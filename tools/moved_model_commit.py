"""Cost of a moved model on a level of C4's size (1 252 161 mesh colliders, 10 M triangles): prtc_update_models + prtc_commit
for two nine-mesh platforms, (a) moved inside their leaves' fat boxes, (b) moved beyond them, in CANONICAL and BY_ID order.
Usage on the GPU box:  PYTHONPATH=$PWD python tools/moved_model_commit.py"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from prototype2_b200 import capi, scenes  # noqa: E402

xyz, mf, mc = scenes.terrain(2237, 2, 2)
px, pmf, pmc = scenes.terrain(6, 2, 2)
px = px.copy(); px[:, 1] = np.float32(0.9)
out = {}
for name, mode in (("canonical", capi.ORDER_CANONICAL), ("by_id", capi.ORDER_BY_ID)):
    ctx = capi.PhysicsContext(order_mode=mode)
    ctx.add_model_collider(xyz, mf, mc)
    ids = [ctx.add_model_collider(px, pmf, pmc, [100.0 + 300 * k, 0.0, 200.0, 1, 0, 0, 0, 1, 1, 1])[0] for k in range(2)]
    ctx.add_capsule_collider()
    t = time.perf_counter(); ctx.commit(); first = time.perf_counter() - t
    tf = np.zeros(2, capi.TRANSFORM_DTYPE)
    tf["rotation"][:, 3] = 1; tf["scale"] = 1
    pos = np.array([[100.0, 0.0, 200.0], [400.0, 0.0, 200.0]], np.float32)
    res = {"first_commit_s": first}
    for label, step in (("inside_margin_ms", 0.004), ("beyond_margin_ms", 0.4)):
        ts = []
        for _ in range(6):
            pos[:, 0] += np.float32(step)
            tf["position"] = pos
            t = time.perf_counter()
            ctx.update_model_colliders(np.array(ids, np.uint32), tf)
            ctx.commit()
            ctx.synchronize()
            ts.append(1e3 * (time.perf_counter() - t))
        res[label] = {"median": float(np.median(ts)), "all": [round(x, 3) for x in ts]}
    res["commit_stats"] = ctx.commit_stats()
    out[name] = res
    ctx.close()
print(json.dumps(out, indent=1))
