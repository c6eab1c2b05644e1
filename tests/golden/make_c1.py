"""C1 = the repo's shipped level with one player capsule (BASELINE.json configs[0]), as a fixture.

Input definition: res/models/level1/level1.obj read in FILE order by a minimal OBJ reader (the file
is already triangulated: 3007 `f` lines of arity 3, 172 `o` groups -> one mesh collider per group); entity transform
and player capsule from res/scenes/cavern.prt:8-9 (position 0, rotation w=0.7071068 y=0.7071068, scale 1.5; capsule
h=0.5 r=0.5 offset (0,0.5,0), start (5.6, 8.0, 5.0)); the 9 NPC entities are dropped (they all start at the same
point -> 0/0 in collideCapsuleCapsule, collision_system.cpp:210-211).  assimp's face reordering is NOT reproduced
(assimp is absent) -- triangle order is "parity unpinned" against a real engine build, pinned against L0 here.
Script: 30 idle frames, 60 frames running +x at 0.2/frame (character.cpp:99), a jump (velocity.y = 0.35,
character_system.cpp:92), then 30 frames running -z, idle to frame 300.  Outputs come from L0 = the reference's own
collision sources (oracle/_ref).  Run in the dev container:  python tests/golden/make_c1.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import helpers as H  # noqa: E402

OBJ = "/root/reference/res/models/level1/level1.obj"
FRAMES = 300
DT = np.float32(1.0 / 30.0)


def read_obj(path):
    verts, tris, first, count = [], [], [], []
    for line in open(path):
        if line.startswith("v "):
            verts.append([float(x) for x in line.split()[1:4]])
        elif line.startswith("o "):
            first.append(3 * len(tris))
            count.append(0)
        elif line.startswith("f "):
            idx = [int(tok.split("/")[0]) - 1 for tok in line.split()[1:]]
            assert len(idx) == 3
            tris.append(idx)
            count[-1] += 3
    v = np.array(verts, np.float32)
    xyz = v[np.array(tris).reshape(-1)]
    return xyz, np.array(first, np.uint32), np.array(count, np.uint32)


def script():
    move = np.zeros((FRAMES, 3), np.float32)
    move[30:90, 0] = 0.2
    move[95:125, 2] = -0.2
    jump = np.zeros(FRAMES, bool)
    jump[90] = True
    jump[200] = True
    return move, jump


def run(world, move, jump):
    out = {k: [] for k in ("pos", "vel", "gnormal", "grounded")}
    for f in range(FRAMES):
        s = world.get_state()
        vel = s["vel"].copy()
        if jump[f]:
            vel[0, 1] = np.float32(0.35)
        world.set_state(vel=vel, move=move[f:f + 1])
        world.step(DT)
        s = world.get_state()
        for k in out:
            out[k].append(s[k])
    return {k: np.stack(v) for k, v in out.items()}


def main():
    xyz, mf, mc = read_obj(OBJ)
    print("level:", len(xyz) // 3, "triangles,", len(mf), "mesh colliders")
    scene = {"xyz": xyz, "mesh_first": mf, "mesh_count": mc,
             "model_transform": [0.0, 0.0, 0.0, 0.7071068, 0.0, 0.7071068, 0.0, 1.5, 1.5, 1.5],
             "capsule": (0.5, 0.5, (0.0, 0.5, 0.0)), "pos": np.array([[5.6, 8.0, 5.0]], np.float32),
             "vel": np.zeros((1, 3), np.float32), "rot": np.array([[1, 0, 0, 0]], np.float32),
             "move": np.zeros((1, 3), np.float32), "n": 1}
    move, jump = script()
    data = {"xyz": xyz, "mesh_first": mf, "mesh_count": mc, "model_transform": np.array(scene["model_transform"], np.float32),
            "capsule": np.array([0.5, 0.5, 0.0, 0.5, 0.0], np.float32), "pos": scene["pos"], "move_script": move,
            "jump_script": jump, "dt": DT, "frames": np.int32(FRAMES)}
    for tag, variant, absm in (("int", "plain", "int"), ("flt", "fabs", "float")):
        r = run(H.build_oracle(scene, "l0", variant=variant), move, jump)
        for k, v in r.items():
            data["%s_%s" % (tag, k)] = v
        rc = run(H.build_oracle(scene, "l1", order="canonical", abs_mode=absm), move, jump)
        data["canonical_ok_%s" % tag] = np.bool_(all(H.same(r[k], rc[k]) for k in r))
        rl = run(H.build_oracle(scene, "l1", order="literal", abs_mode=absm), move, jump)
        assert all(H.same(r[k], rl[k]) for k in r), "L1 literal must equal L0 on C1"
        p = r["pos"][:, 0]
        print(tag, "frame 29 pos", p[29], "frame 89", p[89], "frame 299", p[299], "grounded frames", int(r["grounded"].sum()),
              "canonical_ok", bool(data["canonical_ok_%s" % tag]))
    path = os.path.join(HERE, "c1_level.npz")
    np.savez_compressed(path, **data)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
