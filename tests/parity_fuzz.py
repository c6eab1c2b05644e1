"""Parity fuzz of the CUDA path against the CPU checker (test infrastructure, like tests/): random levels, crowds, capsule
shapes, rotations and inputs, run free for many frames in CANONICAL order with and without the JACOBI dynamic pass (vs the L1 port) and in LITERAL order (vs the
reference's own code when oracle/_ref is there); states compared bit for bit every frame, counters at the end.
CANONICAL rounds take the mesh candidates in ascending collider id (PRTC_ORDER_BY_ID, tree built on the device) half of the time.
Usage on the GPU box:  PYTHONPATH=$PWD python tests/parity_fuzz.py [rounds] [seed] [big]
`big`: crowds of 24 000 - 70 000 characters (persistent kernel, per-slot leaf lists, staged tree top), fewer frames."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H  # noqa: E402
from oraclelib import L0World  # noqa: E402
from prototype2_b200 import capi  # noqa: E402

rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 20
rs = np.random.RandomState(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
big = len(sys.argv) > 3 and sys.argv[3] == "big"
DT = np.float32(1.0 / 30.0)
bad = 0
for r in range(rounds):
    n_terrain = int(rs.choice([96, 160, 256] if big else [12, 24, 40, 64]))
    bx, bz = int(rs.choice([1, 2, 4])), int(rs.choice([1, 2]))
    n = int(rs.choice([24000, 40000, 70000] if big else [1, 7, 40, 200, 700]))
    literal = bool(rs.rand() < 0.4) and n <= 200 and L0World.available("traced")
    jacobi = (not literal) and bool(rs.rand() < 0.35) and n > 1
    by_id = (not literal) and bool(rs.rand() < 0.5)
    cap = (float(rs.choice([0.5, 0.8, 0.25])), float(rs.choice([0.5, 0.3, 0.75])), (0.0, float(rs.choice([0.5, 0.3])), 0.0))
    scene = H.make_scene(n_terrain, bx, bz, n, seed=int(rs.randint(1 << 30)), capsule=cap)
    kind = rs.randint(3)
    if kind == 1:
        scene["rot"] = H.random_rotations(n, seed=int(rs.randint(1 << 30)))
    elif kind == 2:
        scene["rot"] = H.make_scene(n_terrain, bx, bz, n, seed=int(rs.randint(1 << 30)), yaw=True)["rot"]
    if rs.rand() < 0.5:
        a = rs.uniform(-0.4, 0.4)
        scene["model_transform"] = [rs.uniform(-1, 1), rs.uniform(-0.3, 0.3), rs.uniform(-1, 1), np.cos(a / 2), 0.0, np.sin(a / 2), 0.0,
                                    rs.uniform(0.8, 1.6), rs.uniform(0.7, 1.3), rs.uniform(0.8, 1.6)]
    substeps = int(rs.choice([4, 4, 3, 6]))
    abs_mode = int(rs.rand() < 0.3)
    if literal:
        if abs_mode and not L0World.available("fabs"):
            abs_mode = 0
        substeps = 4
        orc = H.build_oracle(scene, "l0", variant="fabs" if abs_mode else "plain")
        ctx, tf, ph = H.build_gpu(scene, order_mode=capi.ORDER_LITERAL, abs_mode=abs_mode)
    else:
        orc = H.build_oracle(scene, "l1", order="by_id" if by_id else "canonical", dyn="jacobi" if jacobi else "off", substeps=substeps,
                             abs_mode="float" if abs_mode else "int", threads=os.cpu_count() or 8)
        ctx, tf, ph = H.build_gpu(scene, substeps=substeps, abs_mode=abs_mode, dyn_mode=capi.DYN_JACOBI if jacobi else capi.DYN_OFF,
                                  order_mode=capi.ORDER_BY_ID if by_id else capi.ORDER_CANONICAL)
    frames = int(rs.choice([8, 14] if big else [60, 120, 200]))
    desc = "round %d: %s terrain %d (%dx%d) n=%d cap=%s rot=%d xform=%s substeps=%d abs=%d frames=%d" % (
        r, "LITERAL" if literal else (("by_id" if by_id else "canonical") + ("+JACOBI" if jacobi else "")), n_terrain, bx, bz, n, cap, kind, scene["model_transform"] is not None, substeps, abs_mode, frames)
    ok = True
    for f in range(frames):
        if f % 12 == 0:
            move = np.zeros((n, 3), np.float32)
            move[:, [0, 2]] = rs.uniform(-0.2, 0.2, (n, 2)).astype(np.float32)
            move[rs.rand(n) < 0.3] = 0.0
            ph["movement"][:] = move
            orc.set_state(move=move)
        if f % 40 == 25:
            s = H.gpu_state(tf, ph)
            jump = s["grounded"].astype(bool) & (rs.rand(n) < 0.5)
            v = s["vel"].copy()
            v[jump, 1] = np.float32(0.35)
            ph["velocity"][:] = v
            orc.set_state(vel=v)
        orc.step(DT)
        ctx.update_characters(DT, tf, ph)
        d = H.state_diff(orc.get_state(), H.gpu_state(tf, ph))
        if d:
            print("MISMATCH", desc, "frame", f, d)
            ok = False
            bad += 1
            break
    if ok and not literal:
        co, cg = orc.counters(), ctx.counters()
        for k in ("substeps", "tri_tests", "contacts") + (("cc_tests",) if jacobi else ()):
            if co[k] != cg[k]:
                print("COUNTER MISMATCH", desc, k, co[k], cg[k])
                ok = False
                bad += 1
                break
    print(("ok   " if ok else "FAIL ") + desc, flush=True)
print("parity fuzz: %d rounds, %d failures" % (rounds, bad))
sys.exit(1 if bad else 0)
