"""Generates the known-answer vectors under tests/golden/ from L0 = the REFERENCE'S OWN collision sources compiled
verbatim (oracle/_ref/libprt_l0*.so, built from /root/reference by oracle/Makefile).  The reference ships no
physics tests, so these vectors are what pins the oracle and the GPU path.  Run in the dev
container (needs /root/reference):   python tests/golden/make_golden.py
Each .npz holds the inputs and, per frame, position / velocity / groundNormal / isGrounded for both resolutions of
the unqualified abs() (collision_system.cpp:72): `int_*` (libstdc++ build) and `flt_*` (float abs), plus the
frame-0 trace of the int build, plus `canonical_ok` = whether the static-only canonical order reproduces the
literal result on this scene (decided with the L1 restatement).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import helpers as H  # noqa: E402
from oraclelib import L0World, Trace  # noqa: E402
from prototype2_b200 import scenes  # noqa: E402

DT = np.float32(1.0 / 30.0)


def quad(p0, p1, p2, p3):
    """two triangles (p0,p1,p2),(p0,p2,p3)"""
    return [p0, p1, p2, p0, p2, p3]


def floor(size=10.0, y=0.0):
    # normal (b-a)x(c-a) must face +y: a=(0,y,0) b=(0,y,s) c=(s,y,s) d=(s,y,0)
    return quad([0, y, 0], [0, y, size], [size, y, size], [size, y, 0])


def scene(tris_per_mesh, pos, vel=None, rot=None, move=None, capsule=(0.5, 0.5, (0.0, 0.5, 0.0)), frames=10, transform=None):
    xyz = np.array([v for m in tris_per_mesh for v in m], np.float32).reshape(-1, 3)
    counts = np.array([len(m) for m in tris_per_mesh], np.uint32)
    first = np.concatenate([[0], np.cumsum(counts)[:-1]]).astype(np.uint32)
    n = len(pos)
    r = np.zeros((n, 4), np.float32)
    r[:, 0] = 1.0
    return {"xyz": xyz, "mesh_first": first, "mesh_count": counts, "model_transform": transform, "capsule": capsule,
            "pos": np.array(pos, np.float32).reshape(n, 3), "vel": np.zeros((n, 3), np.float32) if vel is None else np.array(vel, np.float32).reshape(n, 3),
            "rot": r if rot is None else np.array(rot, np.float32).reshape(n, 4),
            "move": np.zeros((n, 3), np.float32) if move is None else np.array(move, np.float32).reshape(n, 3),
            "n": n, "frames": frames}


def cases():
    c = {}
    c["floor_drop"] = scene([floor()], [[3.0, 1.0, 6.0]], frames=40)
    wall = quad([5, -1, 0], [5, -1, 10], [5, 4, 10], [5, 4, 0])          # normal faces -x
    c["wall_slide"] = scene([floor(), wall], [[3.5, 0.2, 6.0]], move=[[0.2, 0.0, 0.05]], frames=30)
    small = [[4, 0, 4], [4, 0, 5], [5, 0, 5]]
    c["edge_contact"] = scene([small], [[3.7, -0.3, 4.5]], vel=[[0.02, 0.0, 0.0]], frames=6)
    c["vertex_contact"] = scene([small], [[3.8, -0.2, 3.8]], vel=[[0.02, 0.0, 0.02]], frames=6)
    ceiling = quad([0, 1.2, 0], [10, 1.2, 0], [10, 1.2, 10], [0, 1.2, 10])   # normal faces -y: back face for a capsule below
    c["backface_cull"] = scene([ceiling], [[3.0, 0.0, 6.0]], vel=[[0.0, 0.3, 0.0]], frames=4)
    degen = [[1, 0, 1], [2, 0, 2], [2, 0, 2], [1, 0, 1], [1, 0, 1], [1, 0, 1]]
    c["degenerate_skip"] = scene([degen + floor()], [[1.5, 0.1, 4.0]], frames=8)
    c["tunnel_fast_body"] = scene([floor()], [[3.0, 1.2, 6.0]], vel=[[0.0, -2.4, 0.0]], frames=6)
    tall_leaf = floor() + [[0, 3, 0], [0, 3, 10], [10, 3, 10]]           # one mesh whose AABB is tall
    c["abs_trap_hover"] = scene([tall_leaf], [[3.0, 0.2, 6.0]], frames=3)
    c["two_capsules_push"] = scene([floor()], [[3.0, 0.05, 6.0], [3.4, 0.05, 6.1]], frames=12)
    c["tilted_capsule"] = scene([floor()], [[3.0, 0.6, 6.0]], rot=[[0.9659258, 0.2588190, 0.0, 0.0]], frames=25)
    c["scaled_rotated_model"] = scene([floor(4.0)], [[1.0, 1.0, -2.0]], frames=20,
                                      transform=[0.0, 0.0, 0.0, 0.7071068, 0.0, 0.7071068, 0.0, 1.5, 1.5, 1.5])
    t = H.make_scene(24, 2, 2, 30, seed=77, separated=True)
    t["frames"] = 10
    c["terrain_separated"] = t
    t = H.make_scene(20, 1, 1, 40, seed=78)
    t["frames"] = 8
    c["terrain_crowded"] = t
    t = H.make_scene(24, 2, 1, 24, seed=79, separated=True)
    t["rot"] = H.random_rotations(24, seed=5)
    t["frames"] = 8
    c["terrain_rotated"] = t
    return c


def run(world, frames):
    out = {k: [] for k in ("pos", "vel", "gnormal", "grounded")}
    for _ in range(frames):
        world.step(DT)
        s = world.get_state()
        for k in out:
            out[k].append(s[k])
    return {k: np.stack(v) for k, v in out.items()}


def main():
    for name, sc in cases().items():
        data = {"xyz": sc["xyz"], "mesh_first": sc["mesh_first"], "mesh_count": sc["mesh_count"],
                "model_transform": np.array([0, 0, 0, 1, 0, 0, 0, 1, 1, 1] if sc["model_transform"] is None else sc["model_transform"], np.float32),
                "capsule": np.array([sc["capsule"][0], sc["capsule"][1], *sc["capsule"][2]], np.float32),
                "pos": sc["pos"], "vel": sc["vel"], "rot": sc["rot"], "move": sc["move"], "frames": np.int32(sc["frames"]),
                "dt": DT}
        sc = dict(sc, model_transform=list(data["model_transform"]))
        for tag, variant, absm in (("int", "plain", "int"), ("flt", "fabs", "float")):
            r = run(H.build_oracle(sc, "l0", variant=variant), sc["frames"])
            for k, v in r.items():
                data["%s_%s" % (tag, k)] = v
            rc = run(H.build_oracle(sc, "l1", order="canonical", abs_mode=absm), sc["frames"])
            data["canonical_ok_%s" % tag] = np.bool_(all(H.same(r[k], rc[k]) for k in r))
        tr = H.build_oracle(sc, "l0", variant="traced").step(DT, trace=True)
        for k in Trace.FIELDS:
            data["trace_" + k] = getattr(tr, k)
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **data)
        print("%-24s frames %3d  final pos %s grounded(int) %s  canonical_ok %s/%s  %d B" % (
            name, sc["frames"], data["int_pos"][-1][0], data["int_grounded"][-1], data["canonical_ok_int"],
            data["canonical_ok_flt"], os.path.getsize(path)))


if __name__ == "__main__":
    main()
