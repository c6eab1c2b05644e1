"""GPU parity tests proper: every call goes through the C-ABI (libprtc.so) and is compared with the CPU oracle
on identical inputs.  Bar: BIT-EXACT positions, velocities, ground normals, flags, candidate mesh ids (ordered),
hit-triangle indices (ordered), push vectors and post-push positions (north_star asks for bit-exact indices
and 1e-5 relative on floats; this path delivers bit-exact floats too, so the tolerance written here is 0)."""
import os

import numpy as np
import pytest

import helpers as H
from oraclelib import L0World

pytestmark = pytest.mark.gpu

DT = 1.0 / 30.0


def _run_pair(scene, frames, oracle_kw, gpu_kw, trace_frames=(0, 1), frame_cache=True):
    from prototype2_b200 import capi
    orc = H.build_oracle(scene, "l1", **oracle_kw)
    ctx, tf, ph = H.build_gpu(scene, **gpu_kw)
    ctx.commit()
    if not frame_cache:
        ctx.set_option(capi.OPT_FRAME_CACHE, 0)
    assert np.array_equal(ctx.leaf_order(), orc.dump_order()[0]), "flattened tree order differs from the oracle's DFS order"
    for f in range(frames):
        if f in trace_frames:
            to = orc.step(DT, trace=True)
            tg = ctx.update_characters_traced(DT, tf, ph)
            assert H.trace_diff(to, tg) == [], "frame %d trace" % f
        else:
            orc.step(DT)
            ctx.update_characters(DT, tf, ph)
        assert H.state_diff(orc.get_state(), H.gpu_state(tf, ph)) == [], "frame %d state" % f
    # the work counters are the numerators of bench.py's metrics: they must be the reference algorithm's counts
    co, cg = orc.counters(), ctx.counters()
    for k in ("substeps", "tri_tests", "contacts") + (() if frame_cache else ("node_tests",)):
        assert co[k] == cg[k], "counter %s: oracle %d gpu %d" % (k, co[k], cg[k])
    return orc, ctx, tf, ph


@pytest.mark.parametrize("abs_mode", ["int", "float"])
def test_c2_shape_canonical_bit_exact(abs_mode):
    """C2 shape (2-triangle leaves), 1000 capsules, 6 frames, vs L1 canonical."""
    from prototype2_b200 import capi
    scene = H.make_scene(224, 1, 1, 1000)
    _run_pair(scene, 6, dict(order="canonical", abs_mode=abs_mode),
              dict(abs_mode=capi.ABS_INT if abs_mode == "int" else capi.ABS_FLOAT))


@pytest.mark.parametrize("frame_cache,quick_reject,persistent", [(0, 0, 0), (0, 1, 0), (1, 0, 0), (1, 1, 0), (0, 0, 1), (0, 1, 1), (1, 0, 1)])
def test_work_saving_options_do_not_change_results(frame_cache, quick_reject, persistent):
    """PRTC_OPT_FRAME_CACHE / PRTC_OPT_QUICK_REJECT off: same bits; with the frame cache off the node-test counter
    is the reference's own count"""
    from prototype2_b200 import capi
    scene = H.make_scene(96, 2, 2, 800, seed=8)
    scene["vel"] = scene["vel"] * np.float32(1.7)           # some boxes escape the frame region
    orc = H.build_oracle(scene, "l1", order="canonical")
    ctx, tf, ph = H.build_gpu(scene)
    ctx.set_option(capi.OPT_FRAME_CACHE, frame_cache)
    ctx.set_option(capi.OPT_QUICK_REJECT, quick_reject)
    ctx.set_option(capi.OPT_PERSISTENT, persistent)
    for f in range(5):
        if f % 2 == 0:
            to = orc.step(DT, trace=True)
            tg = ctx.update_characters_traced(DT, tf, ph)                 # traced launches use the lock-step kernel
            assert H.trace_diff(to, tg) == []
        else:
            orc.step(DT)
            ctx.update_characters(DT, tf, ph)                             # untraced: persistent kernel when enabled
        assert H.state_diff(orc.get_state(), H.gpu_state(tf, ph)) == []
    co, cg = orc.counters(), ctx.counters()
    for k in ("substeps", "tri_tests", "contacts") + (("node_tests",) if not frame_cache else ()):
        assert co[k] == cg[k], k


def test_rotated_capsules_bit_exact():
    """non-identity, non-normalised rotations: exercises quaternion normalisation + rotation matrix on the device"""
    scene = H.make_scene(96, 2, 2, 400, seed=3)
    scene["rot"] = H.random_rotations(400)
    _run_pair(scene, 5, dict(order="canonical"), dict())


def test_three_substeps_16tri_leaves():
    """C3 shape (4x2-cell = 16-triangle leaves), 3 collide-and-slide iterations"""
    scene = H.make_scene(128, 4, 2, 2000, seed=11)
    _run_pair(scene, 4, dict(order="canonical", substeps=3), dict(substeps=3))


def test_model_transform_scale_rotation():
    """mesh colliders honour T*R*S (physics_system.cpp:83-92): K8 vertex transform against the oracle's cache"""
    from prototype2_b200 import capi
    scene = H.make_scene(48, 2, 2, 100, seed=5)
    scene["model_transform"] = [1.0, -0.5, 2.0, 0.9914449, 0.0, 0.1305262, 0.0, 1.5, 1.25, 1.5]
    # keep capsules over the transformed terrain: rotate/scale their xz the same way (approximately) and lift them
    orc = H.build_oracle(scene, "l1", order="canonical")
    ctx, tf, ph = H.build_gpu(scene)
    ctx.commit()
    cache = np.zeros((len(scene["xyz"]), 3), np.float32)
    orc.lib.l1_world_cache(orc.h, cache.ctypes.data, cache.size)
    assert H.same(cache, ctx.world_vertices())
    for f in range(3):
        orc.step(DT)
        ctx.update_characters(DT, tf, ph)
        assert H.state_diff(orc.get_state(), H.gpu_state(tf, ph)) == []


@pytest.mark.skipif(not L0World.available("traced"), reason="oracle/_ref not built")
def test_against_literal_reference_l0():
    """The reference's own code (L0).  Capsules are kept apart so its capsule-capsule pass never fires and its
    capsule leaves' churn is the only difference: states must match wherever the candidate SEQUENCE matches
    (it always does on the first frame: static leaves are inserted before any capsule moves)."""
    scene = H.make_scene(120, 2, 2, 600, seed=21, separated=True)
    l0 = H.build_oracle(scene, "l0", variant="traced")
    ctx, tf, ph = H.build_gpu(scene)
    t0 = l0.step(DT, trace=True)
    tg = ctx.update_characters_traced(DT, tf, ph)
    assert len(t0.caps_ids) == 0, "scene must not trigger capsule-capsule"
    stats = H.compare_with_literal(t0, tg, l0.get_state(), H.gpu_state(tf, ph), scene["n"])
    print("GPU canonical vs L0 literal:", stats)
    assert stats["chars_in_sync"] > scene["n"] // 2


def test_empty_and_ragged_inputs():
    from prototype2_b200 import capi
    # no colliders at all: every character takes the early break (physics_system.cpp:391-393) and free-falls
    ctx = capi.PhysicsContext()
    ctx.add_capsule_collider()
    tf = capi.make_transforms(np.array([[0.0, 5.0, 0.0]], np.float32))
    ph = capi.make_physics(np.array([[0.1, -0.2, 0.3]], np.float32))
    ctx.update_characters(DT, tf, ph)
    o = H.L1World("canonical")
    o.add_capsule()
    o.set_state(pos=[[0.0, 5.0, 0.0]], vel=[[0.1, -0.2, 0.3]])
    o.step(DT)
    assert H.state_diff(o.get_state(), H.gpu_state(tf, ph)) == []
    # zero characters
    ctx.update_characters(DT, tf[:0], ph[:0])
    # ragged meshes (1..7 triangles per mesh, one empty mesh) + a degenerate triangle
    rs = np.random.RandomState(4)
    xyz, mf, mc = H.scenes.terrain(16, 1, 1)
    tri = xyz.reshape(-1, 3, 3)
    tri[5, 2] = tri[5, 1]                                    # degenerate: two equal vertices
    sizes, first, count, at = rs.randint(1, 8, 200), [], [], 0
    for s in sizes:
        if at + s > len(tri):
            break
        first.append(3 * at); count.append(3 * s); at += s
    first.insert(3, first[3]); count.insert(3, 0)            # an empty mesh collider
    scene = H.make_scene(16, 1, 1, 40, seed=9)
    scene["xyz"], scene["mesh_first"], scene["mesh_count"] = tri.reshape(-1, 3), np.array(first, np.uint32), np.array(count, np.uint32)
    _run_pair(scene, 4, dict(order="canonical"), dict())


def test_candidate_overflow_falls_back_to_the_ordered_walk():
    """more overlapping leaves than the per-lane candidate buffer (12) and a deep frontier: tiny 1x1-cell leaves under
    a fat capsule.  Exercises the ordered multi-pass walk and its agreement with the fast order-free walk."""
    scene = H.make_scene(40, 1, 1, 96, seed=41)
    scene["capsule"] = (2.0, 2.5, (0.0, 2.5, 0.0))
    scene["pos"][:, 1] += 1.0
    _run_pair(scene, 3, dict(order="canonical"), dict())


def test_fast_bodies_tunnel_like_the_reference():
    """T6: bodies faster than the radius per frame fall through; the early `break` applies the rest of the motion
    un-collided.  Must be reproduced, not fixed."""
    scene = H.make_scene(64, 2, 2, 300, seed=33)
    scene["vel"] = scene["vel"] * np.float32(6.0)
    _run_pair(scene, 5, dict(order="canonical"), dict())


def test_movement_and_grounded_inputs():
    scene = H.make_scene(64, 1, 1, 256, seed=17)
    rs = np.random.RandomState(1)
    scene["move"] = (rs.uniform(-0.2, 0.2, (256, 3))).astype(np.float32)
    orc, ctx, tf, ph = _run_pair(scene, 12, dict(order="canonical"), dict(), trace_frames=(0, 5, 11))
    assert ph["grounded"].sum() > 0


# ------------------------------------------------------------------------------------------------------------
# LITERAL order: the reference's unified mesh+capsule tree is maintained on the host every frame with the
# reference's own algorithm and characters keep their sequential visibility of one another -> the result must be
# bit-identical to L0, the reference's own code, end to end (parity level P4), capsule-capsule included.
# ------------------------------------------------------------------------------------------------------------
LITERAL_KEYS = H.TRACE_KEYS + ("q_caps_off", "caps_ids")


def _run_literal_vs_l0(scene, frames, variant="traced", abs_mode=0, trace_frames=(0, 1, 2)):
    from prototype2_b200 import capi
    l0 = H.build_oracle(scene, "l0", variant=variant)
    ctx, tf, ph = H.build_gpu(scene, order_mode=capi.ORDER_LITERAL, abs_mode=abs_mode)
    cc = 0
    for f in range(frames):
        if variant == "traced" and f in trace_frames:
            t0 = l0.step(DT, trace=True)
            tg = ctx.update_characters_traced(DT, tf, ph)
            assert H.trace_diff(t0, tg, LITERAL_KEYS) == [], "frame %d trace" % f
            cc += int((t0.h_poly < 0).sum())
        else:
            l0.step(DT)
            ctx.update_characters(DT, tf, ph)
        assert H.state_diff(l0.get_state(), H.gpu_state(tf, ph)) == [], "frame %d state" % f
    return cc


@pytest.mark.skipif(not L0World.available("traced"), reason="oracle/_ref not built")
def test_literal_mode_equals_reference_crowded():
    scene = H.make_scene(40, 2, 1, 80, seed=101)             # random placement: capsules overlap and push each other
    cc = _run_literal_vs_l0(scene, 25, trace_frames=(0, 1, 2, 3, 10, 24))
    assert cc > 0, "scene was meant to exercise the capsule-capsule pass"


@pytest.mark.skipif(not L0World.available("traced"), reason="oracle/_ref not built")
def test_literal_mode_equals_reference_rotated_moving():
    scene = H.make_scene(48, 1, 1, 60, seed=102)
    scene["rot"] = H.random_rotations(60, seed=9)
    rs = np.random.RandomState(3)
    scene["move"] = rs.uniform(-0.2, 0.2, (60, 3)).astype(np.float32)
    _run_literal_vs_l0(scene, 20)


@pytest.mark.skipif(not L0World.available("traced"), reason="oracle/_ref not built")
def test_literal_mode_with_models_moved_and_added_mid_run():
    """the unified tree's shape depends on the ORDER of its operations: a model moved (updateModelColliders,
    physics_system.cpp:161-202) or added between two frames changes leaves before that frame's capsule leaves move.
    States and traces (candidate order included) against the reference's own code for 120 frames."""
    from prototype2_b200 import capi
    scene = H.make_scene(24, 2, 2, 16, seed=77, yaw=True)
    scene["model_transform"] = [0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 1.5, 1.0, 1.5]
    orc = H.build_oracle(scene, "l0", variant="traced")
    ctx, tf, ph = H.build_gpu(scene, order_mode=capi.ORDER_LITERAL)
    rs = np.random.RandomState(8)
    slab = np.array([[0, 0, 0], [0, 0, 6], [6, 0, 6], [0, 0, 0], [6, 0, 6], [6, 0, 0]], np.float32)
    extra = None
    for f in range(120):
        if f % 10 == 0:
            move = np.zeros((16, 3), np.float32)
            move[:, [0, 2]] = rs.uniform(-0.15, 0.15, (16, 2)).astype(np.float32)
            ph["movement"][:] = move
            orc.set_state(move=move)
        if f in (20, 45, 70):                                     # the map slides / rescales
            t = [0.25 * (f // 20), -0.1, 0.1 * (f // 20), 1.0, 0.0, 0.0, 0.0, 1.5, 1.0 + 0.01 * f, 1.5]
            orc.update_model(0, t)
            g = np.zeros(1, capi.TRANSFORM_DTYPE)
            g["position"][0] = t[0:3]; g["rotation"][0] = (t[4], t[5], t[6], t[3]); g["scale"][0] = t[7:10]
            ctx.update_model_colliders([0], g)
        if f == 30:                                               # a second model appears
            t2 = [10.0, 1.0, 10.0, 0.9961947, 0.0871557, 0.0, 0.0, 1.0, 1.0, 1.0]
            orc.add_model(slab, [0], [6], t2)
            extra, _ = ctx.add_model_collider(slab, np.uint32([0]), np.uint32([6]), t2)
        if f == 60:                                               # ... and moves
            t2 = [12.0, 1.5, 9.0, 0.9961947, 0.0871557, 0.0, 0.0, 1.0, 1.0, 2.0]
            orc.update_model(1, t2)
            g = np.zeros(1, capi.TRANSFORM_DTYPE)
            g["position"][0] = t2[0:3]; g["rotation"][0] = (t2[4], t2[5], t2[6], t2[3]); g["scale"][0] = t2[7:10]
            ctx.update_model_colliders([extra], g)
        if f % 6 == 0 or f in (20, 21, 30, 31, 45, 46, 60, 61, 70, 71):
            to = orc.step(DT, trace=True)
            tg = ctx.update_characters_traced(DT, tf, ph)
            assert H.trace_diff(to, tg, LITERAL_KEYS) == [], "frame %d trace" % f
        else:
            orc.step(DT)
            ctx.update_characters(DT, tf, ph)
        assert H.state_diff(orc.get_state(), H.gpu_state(tf, ph)) == [], "frame %d" % f
        mo, co = orc.dump_order()
        assert np.array_equal(mo, ctx.leaf_order()), "frame %d: mesh leaves in a different tree order" % f


@pytest.mark.skipif(not L0World.available("fabs"), reason="oracle/_ref not built")
def test_literal_mode_equals_reference_float_abs_build():
    from prototype2_b200 import capi
    scene = H.make_scene(64, 4, 2, 50, seed=103)
    scene["model_transform"] = [0.5, 0.25, -1.0, 0.9914449, 0.0, 0.1305262, 0.0, 1.25, 1.0, 1.25]
    _run_literal_vs_l0(scene, 12, variant="fabs", abs_mode=capi.ABS_FLOAT)


@pytest.mark.skipif(not L0World.available("traced"), reason="oracle/_ref not built")
def test_literal_mode_c2_shape_600_capsules():
    """the 12 % of characters that see a different candidate sequence in canonical order (finding 3 / H4) are
    bit-identical here"""
    scene = H.make_scene(120, 2, 2, 600, seed=21, separated=True)
    _run_literal_vs_l0(scene, 6, trace_frames=(0, 5))


import glob
import os

GOLDEN = sorted(p for p in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")) if not os.path.basename(p).startswith("c1_"))
C1_PATH = os.path.join(os.path.dirname(__file__), "golden", "c1_level.npz")


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_golden_vectors_from_the_reference(path):
    """known-answer vectors generated from the reference's own code (tests/golden/make_golden.py): LITERAL order
    must reproduce every one, CANONICAL order those flagged canonical_ok, in both abs resolutions"""
    from prototype2_b200 import capi
    g = np.load(path)
    cap = g["capsule"]
    scene = {"xyz": g["xyz"], "mesh_first": g["mesh_first"], "mesh_count": g["mesh_count"],
             "model_transform": list(g["model_transform"]), "capsule": (float(cap[0]), float(cap[1]), tuple(float(x) for x in cap[2:5])),
             "pos": g["pos"], "vel": g["vel"], "rot": g["rot"], "move": g["move"], "n": len(g["pos"])}
    for tag, absm in (("int", capi.ABS_INT), ("flt", capi.ABS_FLOAT)):
        modes = [capi.ORDER_LITERAL] + ([capi.ORDER_CANONICAL] if bool(g["canonical_ok_%s" % tag]) else [])
        for mode in modes:
            ctx, tf, ph = H.build_gpu(scene, order_mode=mode, abs_mode=absm)
            for f in range(int(g["frames"])):
                if f == 0 and tag == "int" and mode == capi.ORDER_LITERAL:
                    tg = ctx.update_characters_traced(float(g["dt"]), tf, ph)
                    for k in LITERAL_KEYS:
                        assert H.same(tg[k], g["trace_" + k]), "trace %s" % k
                else:
                    ctx.update_characters(float(g["dt"]), tf, ph)
                s = H.gpu_state(tf, ph)
                for k in ("pos", "vel", "gnormal", "grounded"):
                    assert H.same(s[k], g["%s_%s" % (tag, k)][f]), "%s mode %d frame %d %s" % (tag, mode, f, k)


@pytest.mark.parametrize("tag", ["int", "flt"])
@pytest.mark.parametrize("order", ["literal", "canonical"])
def test_c1_shipped_level_300_frames(tag, order):
    """C1 (BASELINE.json configs[0]): shipped level + 1 player, 300 scripted frames, against the reference's output"""
    from prototype2_b200 import capi
    g = np.load(C1_PATH)
    ctx, tf, ph = H.build_gpu(H.c1_scene(g), order_mode=capi.ORDER_LITERAL if order == "literal" else capi.ORDER_CANONICAL,
                              abs_mode=capi.ABS_INT if tag == "int" else capi.ABS_FLOAT)
    assert bool(g["canonical_ok_%s" % tag])
    for f in range(int(g["frames"])):
        if g["jump_script"][f]:
            ph["velocity"][0, 1] = np.float32(0.35)
        ph["movement"][0] = g["move_script"][f]
        ctx.update_characters(float(g["dt"]), tf, ph)
        s = H.gpu_state(tf, ph)
        for k in ("pos", "vel", "gnormal", "grounded"):
            assert H.same(s[k], g["%s_%s" % (tag, k)][f]), "frame %d %s" % (f, k)


def test_raycast_matches_the_reference():
    """f1: PhysicsSystem::raycast (physics_system.cpp:205-242) -- hit flag and hit point bit-exact against L1 (and the
    reference's own code when oracle/_ref is present) for rays through the shipped level and a terrain"""
    from prototype2_b200 import capi
    g = np.load(C1_PATH)
    scene = H.c1_scene(g)
    rs = np.random.RandomState(12)
    n = 2000
    origins = np.stack([rs.uniform(-20, 20, n), rs.uniform(0.5, 10, n), rs.uniform(-20, 20, n)], 1).astype(np.float32)
    dirs = rs.normal(size=(n, 3)).astype(np.float32)
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True).astype(np.float32)
    dirs[::7, 0] = 0.0                                         # axis-parallel components take the slab test's other branch
    dirs[::11, 1] = 0.0
    maxd = rs.uniform(1.0, 30.0, n).astype(np.float32)
    orcs = [H.build_oracle(scene, "l1", order="literal")]
    if L0World.available("plain"):
        orcs.append(H.build_oracle(scene, "l0", variant="plain"))
    for mode in (capi.ORDER_CANONICAL, capi.ORDER_LITERAL):
        ctx, tf, ph = H.build_gpu(scene, order_mode=mode)
        mask, hit = ctx.raycast(origins, dirs, maxd)
        assert mask.sum() > 100
        for orc in orcs:
            for i in range(n):
                ok, h = orc.raycast(origins[i], dirs[i], float(maxd[i]))
                assert ok == bool(mask[i]), i
                if ok:
                    assert H.same(h, hit[i]), i


def test_remove_and_move_model_colliders():
    """f2 + removeCollider: moving a model re-derives its world cache / leaf boxes (physics_system.cpp:161-202),
    removing one takes its leaves out of the tree; compared with an oracle world built directly in the final state"""
    from prototype2_b200 import capi
    xyz, mf, mc = H.scenes.terrain(24, 2, 2)
    pos, vel = H.scenes.capsules_grid(40, 24, seed=4)
    ctx = capi.PhysicsContext()
    m0, _ = ctx.add_model_collider(xyz, mf, mc)
    m1, _ = ctx.add_model_collider(xyz, mf, mc, [0, 3.0, 0, 1, 0, 0, 0, 1, 1, 1])   # a second terrain 3 units up
    for _ in range(40):
        ctx.add_capsule_collider()
    tf = capi.make_transforms(pos)
    ph = capi.make_physics(vel)
    ctx.update_characters(DT, tf, ph)                          # both present
    ctx.remove_model_collider(m1)
    for f in range(3):                                          # (bit equality under the same history: see the next test)
        ctx.update_characters(DT, tf, ph)
    # move the remaining model: world cache must equal a fresh registration at the new transform
    t = np.zeros(1, capi.TRANSFORM_DTYPE)
    t["position"][0] = (1.0, 0.5, -2.0); t["rotation"][0] = (0.0, 0.1305262, 0.0, 0.9914449); t["scale"][0] = (1.0, 1.5, 1.0)
    ctx.update_model_colliders([m0], t)
    ctx.commit()
    fresh = capi.PhysicsContext()
    fresh.add_model_collider(xyz, mf, mc, [1.0, 0.5, -2.0, 0.9914449, 0.0, 0.1305262, 0.0, 1.0, 1.5, 1.0])
    fresh.commit()
    nv = len(xyz)
    assert H.same(ctx.world_vertices()[:nv], fresh.world_vertices()[:nv])


# ------------------------------------------------------------------------------------------------------------
# CANONICAL order + JACOBI dynamic pass (C5 shape): neighbours from the stored fat boxes in ascending index, seen at
# their start-of-frame pose.  Oracle = L1 canonical/jacobi (brute force over all pairs).
# ------------------------------------------------------------------------------------------------------------
def _patch(x0, z0, y, n=6):
    """one mesh collider: an n x n-cell horizontal patch at height y with its corner at (x0, z0)"""
    xyz, mf, mc = H.scenes.terrain(n, n, n)
    xyz = xyz.copy()
    xyz[:, 1] = np.float32(y)
    xyz[:, 0] += np.float32(x0)
    xyz[:, 2] += np.float32(z0)
    assert len(mf) == 1
    return xyz, mf, mc


@pytest.mark.parametrize("order", ["canonical", "by_id"])
def test_moved_models_take_the_incremental_commit(order):
    """a level plus two small platforms that are moved every frame (updateModelColliders, physics_system.cpp:161-202) --
    sometimes inside their leaves' fat boxes (the tree is untouched: only the moved meshes are re-transformed and their
    triangle records rebuilt), sometimes beyond (CANONICAL: the host tree re-inserts and is re-flattened; BY_ID: the device
    rebuilds the tree from the changed leaf records) -- bit-identical to the oracle every frame, and the commit statistics
    show that the level was not rebuilt for the small moves"""
    from prototype2_b200 import capi
    scene = H.make_scene(64, 2, 2, 3000, seed=91)
    px, pmf, pmc = H.scenes.terrain(6, 2, 2)                      # a 6x6-cell platform of nine meshes
    px = px.copy(); px[:, 1] = np.float32(0.9)
    orc = H.build_oracle(scene, "l1", order=order)
    ctx, tf, ph = H.build_gpu(scene, order_mode=capi.ORDER_BY_ID if order == "by_id" else capi.ORDER_CANONICAL)
    plats = []
    for k, at in enumerate(((10.0, 0.0, 12.0), (40.0, 0.2, 30.0))):
        t0 = [at[0], at[1], at[2], 1, 0, 0, 0, 1, 1, 1]
        plats.append((orc.add_model(px, pmf, pmc, t0), ctx.add_model_collider(px, pmf, pmc, t0)[0], list(at)))
    ctx.set_option(capi.OPT_WARP_KERNEL, 0)
    full_before = None
    for f in range(24):
        if f > 0:
            step = 0.004 if f % 6 else 0.3                        # mostly inside the 0.05 margin, now and then beyond it
            ids, tfs = [], np.zeros(len(plats), capi.TRANSFORM_DTYPE)
            for k, (oid, gid, at) in enumerate(plats):
                at[0] += step; at[1] += 0.5 * step * (1 if k else -1)
                orc.update_model(oid, [at[0], at[1], at[2], 1, 0, 0, 0, 1, 1, 1])
                ids.append(gid)
                tfs["position"][k] = at; tfs["rotation"][k] = (0, 0, 0, 1); tfs["scale"][k] = (1, 1, 1)
            ctx.update_model_colliders(np.array(ids, np.uint32), tfs)
        orc.step(DT)
        ctx.update_characters(DT, tf, ph)
        assert H.state_diff(orc.get_state(), H.gpu_state(tf, ph)) == [], "frame %d" % f
        if f == 1:
            full_before = ctx.commit_stats()["full"]
    st = ctx.commit_stats()
    assert st["incremental"] >= 15, st                            # the small moves (and, by id, the big ones too)
    if order == "by_id":
        assert st["full"] == full_before, st                       # the level's tree is never rebuilt on the host
    else:
        assert st["full"] - full_before <= 5, st                   # only the moves that re-insert a leaf re-flatten


class _L1Slots:
    """L1 canonical world behind the reference's model-index reuse (LIFO free list, physics_system.cpp:49-56,122-123)"""

    def __init__(self):
        self.w, self.slots, self.free = H.L1World("canonical"), [], []

    def add_model(self, xyz, mf, mc):
        mid = self.w.add_model(xyz, mf, mc)
        if self.free:
            k = self.free.pop()
            self.slots[k] = mid
            return k
        self.slots.append(mid)
        return len(self.slots) - 1

    def remove_model(self, k):
        self.w.remove_model(self.slots[k])
        self.free.append(k)

    def __getattr__(self, name):
        return getattr(self.w, name)


@pytest.mark.skipif(not L0World.available("plain"), reason="oracle/_ref not built")
@pytest.mark.parametrize("order", ["literal", "canonical"])
def test_remove_and_re_add_model_colliders_same_history_as_the_reference(order):
    """PhysicsSystem::removeCollider + addModelCollider (physics_system.cpp:108-149, 44-106; caller scene.cpp:354-359) against
    the reference's own code under the SAME history: three one-mesh models, the last one removed, frames, re-added somewhere
    else, frames, removed again together with the one before it, frames, both re-added.  (The reference is only well defined
    for that envelope -- last model first, one mesh per model -- see oracle/l0_driver.cpp l0_remove_model.)  LITERAL order:
    bit-equal states every frame (the unified tree goes through the same removals and insertions); CANONICAL order: the
    mesh-only tree is rebuilt the same way, compared where the candidate sequences agree -- here always, the patches do
    not overlap."""
    from prototype2_b200 import capi
    patches = [_patch(0, 0, 0.0), _patch(6, 0, 0.25), _patch(0, 6, -0.25)]
    # LITERAL order against the reference's own code; CANONICAL (mesh-only tree, same removals and insertions) against the
    # restatement, whose remove is pinned to the reference's in tests/test_oracle_pinning.py
    orc = H.L0World("plain") if order == "literal" else _L1Slots()
    ctx = capi.PhysicsContext(order_mode=capi.ORDER_LITERAL if order == "literal" else capi.ORDER_CANONICAL)
    tags, ids = [], []
    for xyz, mf, mc in patches:
        tags.append(orc.add_model(xyz, mf, mc))
        ids.append(ctx.add_model_collider(xyz, mf, mc)[0])
    assert tags == [0, 1, 2]
    n = 60
    rs = np.random.RandomState(11)
    pos = np.stack([rs.uniform(0.8, 11.2, n), rs.uniform(0.8, 1.6, n), rs.uniform(0.8, 11.2, n)], axis=1).astype(np.float32)
    vel = rs.uniform(-0.05, 0.05, (n, 3)).astype(np.float32)
    for _ in range(n):
        orc.add_capsule()
        ctx.add_capsule_collider()
    orc.set_state(pos=pos, vel=vel)
    tf, ph = capi.make_transforms(pos), capi.make_physics(vel)

    def frames(k, what):
        for f in range(k):
            orc.step(DT)
            ctx.update_characters(DT, tf, ph)
            assert H.state_diff(orc.get_state(), H.gpu_state(tf, ph)) == [], "%s, frame %d" % (what, f)
    frames(6, "three models")
    orc.remove_model(tags[2]); ctx.remove_model_collider(ids[2])
    frames(6, "last model removed")                               # capsules over the hole fall
    x2 = _patch(0, 6, -0.75)
    assert orc.add_model(*x2) == 2                                # the freed index comes back (physics_system.cpp:49-51)
    ids[2] = ctx.add_model_collider(*x2)[0]
    frames(6, "re-added lower")
    orc.remove_model(2); ctx.remove_model_collider(ids[2])
    orc.remove_model(1); ctx.remove_model_collider(ids[1])
    frames(5, "two models removed")
    x1 = _patch(6, 0, -0.5)
    assert orc.add_model(*x1) == 1 and orc.add_model(*x2) == 2    # LIFO: the last freed first
    ids[1] = ctx.add_model_collider(*x1)[0]
    ids[2] = ctx.add_model_collider(*x2)[0]
    frames(8, "both re-added")


def _run_jacobi(scene, frames, trace_frames=(0, 1, 2)):
    from prototype2_b200 import capi
    orc = H.build_oracle(scene, "l1", order="canonical", dyn="jacobi")
    ctx, tf, ph = H.build_gpu(scene, dyn_mode=capi.DYN_JACOBI)
    cc_hits = 0
    for f in range(frames):
        if f in trace_frames:
            to = orc.step(DT, trace=True)
            tg = ctx.update_characters_traced(DT, tf, ph)
            assert H.trace_diff(to, tg, LITERAL_KEYS) == [], "frame %d trace" % f
            cc_hits += int((to.h_poly < 0).sum())
        else:
            orc.step(DT)
            ctx.update_characters(DT, tf, ph)
        assert H.state_diff(orc.get_state(), H.gpu_state(tf, ph)) == [], "frame %d state" % f
    co, cg = orc.counters(), ctx.counters()
    for k in ("substeps", "tri_tests", "contacts", "cc_tests"):
        assert co[k] == cg[k], k
    return cc_hits


def test_jacobi_dynamic_pass_crowded():
    scene = H.make_scene(48, 2, 2, 700, seed=55)             # ~0.35 capsules per unit area: plenty of contacts
    assert _run_jacobi(scene, 8, trace_frames=(0, 1, 4, 7)) > 0


def test_jacobi_dynamic_pass_more_neighbours_than_the_buffer():
    """a pile of 90 capsules on a 6x6 patch: far more than 16 neighbours per query -> multi-pass neighbour lists"""
    scene = H.make_scene(32, 2, 2, 90, seed=56)
    rs = np.random.RandomState(2)
    scene["pos"][:, 0] = 12.0 + rs.uniform(0, 6, 90).astype(np.float32)
    scene["pos"][:, 2] = 12.0 + rs.uniform(0, 6, 90).astype(np.float32)
    scene["pos"][:, 1] = H.scenes.height(scene["pos"][:, 0], scene["pos"][:, 2]) + rs.uniform(0, 2, 90).astype(np.float32)
    assert _run_jacobi(scene, 5) > 50


def test_jacobi_grid_of_a_million_cells():
    """a stray triangle 3 km away stretches the world, so the uniform grid of the dynamic pass has ~10^6 cells: the prefix
    sum over the cell counts runs over several hundred tiles (look-back over more than one window of 32) with a ragged last
    tile; some capsules stand next to the stray triangle, i.e. in the last cells"""
    scene = H.make_scene(48, 2, 2, 600, seed=58)
    far = np.array([[3001.0, 0.0, 2777.0], [3001.0, 0.0, 2779.0], [3003.0, 0.0, 2779.0]], np.float32)
    scene["xyz"] = np.concatenate([scene["xyz"], far])
    scene["mesh_first"] = np.concatenate([scene["mesh_first"], [len(scene["xyz"]) - 3]]).astype(np.uint32)
    scene["mesh_count"] = np.concatenate([scene["mesh_count"], [3]]).astype(np.uint32)
    rs = np.random.RandomState(4)
    scene["pos"][:40, 0] = 3000.5 + rs.uniform(0, 2.5, 40).astype(np.float32)
    scene["pos"][:40, 2] = 2776.5 + rs.uniform(0, 2.5, 40).astype(np.float32)
    scene["pos"][:40, 1] = rs.uniform(0.0, 1.0, 40).astype(np.float32)
    assert _run_jacobi(scene, 6, trace_frames=(0, 3)) > 0


def test_jacobi_rotated_capsules_and_no_static_geometry():
    from prototype2_b200 import capi
    scene = H.make_scene(40, 2, 2, 300, seed=57)
    scene["rot"] = H.random_rotations(300, seed=3)
    _run_jacobi(scene, 5)
    # capsules only (no mesh collider at all): the pass still runs before the early break (T7)
    orc = H.L1World("canonical", dyn="jacobi")
    ctx = capi.PhysicsContext(dyn_mode=capi.DYN_JACOBI)
    rs = np.random.RandomState(9)
    pos = rs.uniform(0, 6, (80, 3)).astype(np.float32)
    vel = rs.uniform(-0.1, 0.1, (80, 3)).astype(np.float32)
    for _ in range(80):
        orc.add_capsule()
        ctx.add_capsule_collider()
    orc.set_state(pos=pos, vel=vel)
    tf, ph = capi.make_transforms(pos), capi.make_physics(vel)
    for f in range(4):
        orc.step(DT)
        ctx.update_characters(DT, tf, ph)
        assert H.state_diff(orc.get_state(), H.gpu_state(tf, ph)) == []
    assert ctx.counters()["cc_tests"] > 0


def test_c5_full_size_jacobi_against_the_oracle():
    """config 5 at full size: 256 000 capsules x 1 002 528 triangles with the dynamic-vs-dynamic pass, 4 frames, every
    character compared with the CPU oracle (grid-accelerated, all host threads) bit for bit, counters included; the same
    batch split over two contexts behind prtc_multi gives the same bytes"""
    from prototype2_b200 import capi
    n = 256000
    scene = H.make_scene(708, 2, 2, n, seed=12345)
    morton = H.scenes.morton_order(scene["pos"], 708)
    for k in ("pos", "vel", "rot", "move"):
        scene[k] = np.ascontiguousarray(scene[k][morton])
    orc = H.L1World("canonical", dyn="jacobi", threads=os.cpu_count() or 1)
    orc.add_model(scene["xyz"], scene["mesh_first"], scene["mesh_count"], None)
    ctx = capi.PhysicsContext(dyn_mode=capi.DYN_JACOBI)
    ctx.add_model_collider(scene["xyz"], scene["mesh_first"], scene["mesh_count"])
    multi = capi.MultiContext([0, 0], dyn_mode=capi.DYN_JACOBI)
    multi.add_model_collider(scene["xyz"], scene["mesh_first"], scene["mesh_count"])
    cid = ctx.add_capsule_collider()
    multi.add_capsule_collider()
    for _ in range(n):
        orc.add_capsule()
    orc.set_state(pos=scene["pos"], vel=scene["vel"])
    tf = capi.make_transforms(scene["pos"])
    ph = capi.make_physics(scene["vel"], collider=np.full(n, cid, np.uint32))
    tf2, ph2 = tf.copy(), ph.copy()
    for f in range(4):
        orc.step(DT)
        ctx.update_characters(DT, tf, ph)
        multi.update_characters(DT, tf2, ph2)
        assert H.state_diff(orc.get_state(), H.gpu_state(tf, ph)) == [], "frame %d" % f
        assert H.state_diff(H.gpu_state(tf, ph), H.gpu_state(tf2, ph2)) == [], "frame %d, two contexts" % f
    co, cg = orc.counters(), ctx.counters()
    for k in ("substeps", "tri_tests", "contacts", "cc_tests"):
        assert co[k] == cg[k], k
    assert cg["cc_tests"] > 100000


def test_jacobi_two_ranks_emulated_on_one_gpu_equal_one_rank():
    """multi-rank dynamic pass without a second GPU: two contexts on cuda:0 each own half of the characters, prepare
    their records, exchange them (device copies standing in for the NCCL all-gather), step -- and the concatenated
    result must be byte-identical to one context stepping everybody (results independent of G)."""
    import torch
    from prototype2_b200 import capi, sharding
    scene = H.make_scene(48, 2, 2, 600, seed=61)
    n = scene["n"]
    one, tf1, ph1 = H.build_gpu(scene, dyn_mode=capi.DYN_JACOBI)
    ranks = []
    for r in range(2):
        first, count = sharding.partition(n, 2, r)
        ctx, tf, ph = H.build_gpu(scene, dyn_mode=capi.DYN_JACOBI)        # every rank registers the same colliders
        ctx.dyn_configure(first, n)
        d_tf = torch.from_numpy(tf[first:first + count].view(np.uint8).reshape(-1)).cuda()
        d_ph = torch.from_numpy(ph[first:first + count].view(np.uint8).reshape(-1)).cuda()
        ranks.append((ctx, first, count, d_tf, d_ph))
    for f in range(6):
        one.update_characters(DT, tf1, ph1)
        recs = []
        for ctx, first, count, d_tf, d_ph in ranks:
            ctx.dyn_prepare(DT, count, d_tf.data_ptr(), d_ph.data_ptr())
            ctx.synchronize()
            recs.append(sharding.records_tensor(ctx)[0])
        for r, (ctx, first, count, _, _) in enumerate(ranks):              # "all-gather": copy every slice everywhere
            for q, (_, qfirst, qcount, _, _) in enumerate(ranks):
                if q != r:
                    recs[r][qfirst:qfirst + qcount].copy_(recs[q][qfirst:qfirst + qcount])
        torch.cuda.synchronize()
        for ctx, first, count, d_tf, d_ph in ranks:
            ctx.update_characters_device(DT, count, d_tf.data_ptr(), d_ph.data_ptr())
            ctx.synchronize()
        got_tf = np.concatenate([d_tf.cpu().numpy().view(capi.TRANSFORM_DTYPE) for _, _, _, d_tf, _ in ranks])
        got_ph = np.concatenate([d_ph.cpu().numpy().view(capi.PHYSICS_DTYPE) for _, _, _, _, d_ph in ranks])
        assert H.state_diff(H.gpu_state(tf1, ph1), H.gpu_state(got_tf, got_ph)) == [], "frame %d" % f
    assert one.counters()["cc_tests"] > 0


@pytest.mark.parametrize("pinned,streamed,opts", [(False, 1, {}), (True, 1, {}), (True, 0, {}),
                                                  (True, 1, {"OPT_CROSS_FRAME_CACHE": 0}), (True, 1, {"OPT_FRAME_CACHE": 0}),
                                                  (True, 1, {"OPT_QUICK_REJECT": 0})])
def test_streamed_host_step_equals_device_step(pinned, streamed, opts):
    """big host batches run as one persistent launch fed by the copy-in stream while it runs, results leaving chunk by
    chunk: same bits as one device-resident launch -- from pageable and pinned buffers, with an odd tail chunk, and when
    the caller edits the state between frames (the per-slot leaf lists were refreshed for a frame that did not come).
    streamed=0 is the older arrangement (chunks copied in, stepped and copied out on three streams)."""
    import torch
    from prototype2_b200 import capi
    n = 300000 + 37
    scene = H.make_scene(200, 2, 2, n, seed=71)
    ctx, tf, ph = H.build_gpu(scene)
    ctx.set_option(capi.OPT_STREAMED_COPY, streamed)
    for k, v in opts.items():
        ctx.set_option(getattr(capi, k), v)
    if pinned:
        t_tf = torch.from_numpy(tf.view(np.uint8).reshape(-1).copy()).pin_memory()
        t_ph = torch.from_numpy(ph.view(np.uint8).reshape(-1).copy()).pin_memory()
        tf, ph = t_tf.numpy().view(capi.TRANSFORM_DTYPE), t_ph.numpy().view(capi.PHYSICS_DTYPE)
    d_tf = torch.from_numpy(tf.view(np.uint8).reshape(-1).copy()).cuda()
    d_ph = torch.from_numpy(ph.view(np.uint8).reshape(-1).copy()).cuda()
    rs = np.random.RandomState(5)
    for f in range(6):
        if f == 3:                                                # a third of the crowd is told to run / teleported
            who = rs.rand(n) < 0.33
            ph["movement"][who, 0] = np.float32(0.2)
            tf["position"][who, 0] += rs.uniform(-3, 3, int(who.sum())).astype(np.float32)
            d_tf.copy_(torch.from_numpy(tf.view(np.uint8).reshape(-1)))
            d_ph.copy_(torch.from_numpy(ph.view(np.uint8).reshape(-1)))
        ctx.update_characters(DT, tf, ph)                         # host buffers: streamed
        ctx.update_characters_device(DT, n, d_tf.data_ptr(), d_ph.data_ptr())
        ctx.synchronize()
        got_tf = d_tf.cpu().numpy().view(capi.TRANSFORM_DTYPE)
        got_ph = d_ph.cpu().numpy().view(capi.PHYSICS_DTYPE)
        assert H.state_diff(H.gpu_state(tf, ph), H.gpu_state(got_tf, got_ph)) == [], "frame %d" % f
        assert np.array_equal(tf["rotation"], got_tf["rotation"]) and np.array_equal(ph["collider"], got_ph["collider"])


# ------------------------------------------------------------------------------------------------------------
# BASELINE.json sizes
# ------------------------------------------------------------------------------------------------------------
def test_c3_full_size_bit_exact():
    """C3 at full size: 64 000 capsules x 1 002 528 triangles (62 658 leaves of 16), 3 collide-and-slide iterations,
    against the multi-threaded L1 port"""
    scene = H.make_scene(708, 4, 2, 64000, seed=12345)
    orc = H.build_oracle(scene, "l1", order="canonical", substeps=3, threads=os.cpu_count() or 1)
    ctx, tf, ph = H.build_gpu(scene, substeps=3)
    for f in range(3):
        orc.step(DT)
        ctx.update_characters(DT, tf, ph)
        assert H.state_diff(orc.get_state(), H.gpu_state(tf, ph)) == [], "frame %d" % f
    co, cg = orc.counters(), ctx.counters()
    for k in ("substeps", "tri_tests", "contacts"):
        assert co[k] == cg[k], k


@pytest.mark.parametrize("order", ["canonical", "by_id"])
def test_c4_full_size_properties(order):
    """C4 at full size (1 M capsules x 10 M triangles), CANONICAL order and the device-built BY_ID tree:
      * a 65 536-character sample replayed by the CPU oracle against the full terrain for 12 frames is bit-identical
        (the persistent kernel's cross-frame cache is in its steady state by then; characters are independent, so a
        sample's frames are the full run's);
      * work-saving devices off (walk per substep, no pre-cull) == on, bit for bit, and the counters obey
        substeps / tri_tests / contacts equal, fewer exact tests with the cull;
      * batch independence: stepping two halves separately == stepping the whole batch;
      * permutation equivariance: permuting the characters permutes the result."""
    from prototype2_b200 import capi
    n, frames = 1000000, 12
    scene = H.make_scene(2237, 2, 2, n, seed=12345)
    morton = H.scenes.morton_order(scene["pos"], 2237)
    for k in ("pos", "vel", "rot", "move"):
        scene[k] = np.ascontiguousarray(scene[k][morton])
    ctx = capi.PhysicsContext(order_mode=capi.ORDER_BY_ID if order == "by_id" else capi.ORDER_CANONICAL)
    ctx.add_model_collider(scene["xyz"], scene["mesh_first"], scene["mesh_count"])
    cid = ctx.add_capsule_collider()
    ctx.commit()
    tf0 = capi.make_transforms(scene["pos"])
    ph0 = capi.make_physics(scene["vel"], collider=np.full(n, cid, np.uint32))

    def run(nframes, tf, ph, **opts):
        for o, v in opts.items():
            ctx.set_option(getattr(capi, o), v)
        ctx.counters(reset=True)
        for _ in range(nframes):
            ctx.update_characters(DT, tf, ph)
        c = ctx.counters(reset=True)
        ctx.set_option(capi.OPT_FRAME_CACHE, 1)
        ctx.set_option(capi.OPT_QUICK_REJECT, 1)
        ctx.set_option(capi.OPT_TIGHT_REJECT, 1)
        return c
    # oracle on a sample against the full terrain, every frame
    sel = np.linspace(0, n - 1, 65536).astype(np.int64)
    sub = dict(scene, pos=scene["pos"][sel], vel=scene["vel"][sel], rot=scene["rot"][sel], move=scene["move"][sel], n=len(sel))
    orc = H.build_oracle(sub, "l1", order=order, threads=os.cpu_count() or 1)
    tf_a, ph_a = tf0.copy(), ph0.copy()
    for f in range(frames):
        orc.step(DT)
        ctx.update_characters(DT, tf_a, ph_a)
        assert H.state_diff(orc.get_state(), H.gpu_state(tf_a[sel], ph_a[sel])) == [], "frame %d" % f
    # options
    tf_a, ph_a = tf0.copy(), ph0.copy()
    ca = run(3, tf_a, ph_a)
    tf_b, ph_b = tf0.copy(), ph0.copy()
    cb = run(3, tf_b, ph_b, OPT_FRAME_CACHE=0, OPT_QUICK_REJECT=0, OPT_TIGHT_REJECT=0)
    assert H.state_diff(H.gpu_state(tf_a, ph_a), H.gpu_state(tf_b, ph_b)) == []
    for k in ("substeps", "tri_tests", "contacts"):
        assert ca[k] == cb[k], k
    assert ca["exact_tests"] < cb["exact_tests"] and ca["exact_tests"] < ca["tri_tests"]
    assert ca["node_tests"] < cb["node_tests"]
    if order == "by_id":
        return
    # halves
    tf_c, ph_c = tf0.copy(), ph0.copy()
    for _ in range(3):
        ctx.update_characters(DT, tf_c[:n // 2], ph_c[:n // 2])
        ctx.update_characters(DT, tf_c[n // 2:], ph_c[n // 2:])
    assert H.state_diff(H.gpu_state(tf_a, ph_a), H.gpu_state(tf_c, ph_c)) == []
    # permutation
    perm = np.random.RandomState(3).permutation(n)
    tf_d, ph_d = tf0[perm].copy(), ph0[perm].copy()
    for _ in range(3):
        ctx.update_characters(DT, tf_d, ph_d)
    assert H.state_diff(H.gpu_state(tf_a[perm], ph_a[perm]), H.gpu_state(tf_d, ph_d)) == []


@pytest.mark.parametrize("dyn,n", [("off", 64), ("jacobi", 64), ("off", 5000), ("jacobi", 5000)])
def test_unknown_collider_id_is_reported_not_dereferenced(dyn, n):
    """a character naming a capsule that was never added: the step returns PRTC_ERR_INVALID, nothing out of range is read
    (static pass and the dynamic pass's record kernel alike, small and pooled batches), the others are stepped"""
    from prototype2_b200 import capi
    scene = H.make_scene(48, 2, 2, n, seed=2)
    ctx, tf, ph = H.build_gpu(scene, dyn_mode=capi.DYN_JACOBI if dyn == "jacobi" else capi.DYN_OFF)
    before = tf["position"].copy()
    ph["collider"][5] = 0x7FFFFFF0
    with pytest.raises(capi.PrtcError) as e:
        ctx.update_characters(DT, tf, ph)
    assert e.value.code == capi.PRTC_ERR_INVALID and "never added" in str(e.value)
    assert not H.same(tf["position"][6], before[6])                # the others were stepped
    ph["collider"][5] = 5
    ctx.update_characters(DT, tf, ph)                              # and the context keeps working


def test_pre_cull_survives_empty_meshes_and_removed_models():
    """a mesh without vertices (and a removed model) has the fold's +-FLT_MAX start values as its box; that must not be
    taken for the extent of the world, or the pre-cull switches itself off for good"""
    from prototype2_b200 import capi
    scene = H.make_scene(48, 2, 2, 3000, seed=9)
    mc = scene["mesh_count"].copy()
    mc[3] = 0                                                      # an empty mesh collider
    scene["mesh_count"] = mc
    orc = H.build_oracle(scene, "l1", order="by_id")               # (by-id order: adding and removing a model leaves the others' order alone)
    ctx, tf, ph = H.build_gpu(scene, order_mode=capi.ORDER_BY_ID)
    extra = ctx.add_model_collider(scene["xyz"][:36] + np.float32(500.0), np.array([0], np.uint32), np.array([36], np.uint32))
    ctx.commit()
    ctx.remove_model_collider(extra[0])
    for f in range(3):
        orc.step(DT)
        ctx.update_characters(DT, tf, ph)
        assert H.state_diff(orc.get_state(), H.gpu_state(tf, ph)) == [], "frame %d" % f
    c = ctx.counters()
    assert c["exact_tests"] < c["tri_tests"] // 2                  # the cull is at work


# ------------------------------------------------------------------------------------------------------------
# robustness: heterogeneous capsules, odd substep counts, far-away worlds, NaN input
# ------------------------------------------------------------------------------------------------------------
def _hetero_worlds(scene, caps, oracle_kind, oracle_kw, gpu_kw):
    from prototype2_b200 import capi
    orc = H.L0World(oracle_kw.pop("variant")) if oracle_kind == "l0" else H.L1World(**oracle_kw)
    ctx = capi.PhysicsContext(**gpu_kw)
    orc.add_model(scene["xyz"], scene["mesh_first"], scene["mesh_count"], scene["model_transform"])
    ctx.add_model_collider(scene["xyz"], scene["mesh_first"], scene["mesh_count"], scene["model_transform"])
    for h, r, off in caps:
        orc.add_capsule(h, r, off)
        ctx.add_capsule_collider(h, r, off)
    orc.set_state(pos=scene["pos"], rot=scene["rot"], vel=scene["vel"], move=scene["move"])
    tf = capi.make_transforms(scene["pos"], rot_xyzw=scene["rot"][:, [1, 2, 3, 0]])
    ph = capi.make_physics(scene["vel"], movement=scene["move"])
    return orc, ctx, tf, ph


def _random_capsules(n, seed):
    rs = np.random.RandomState(seed)
    return [(float(rs.uniform(0.2, 1.2)), float(rs.uniform(0.2, 0.8)), (float(rs.uniform(-0.1, 0.1)), float(rs.uniform(0.2, 0.8)), float(rs.uniform(-0.1, 0.1))))
            for _ in range(n)]


@pytest.mark.skipif(not L0World.available("plain"), reason="oracle/_ref not built")
def test_heterogeneous_capsules_literal_vs_reference():
    """every capsule its own height / radius / offset (capsule-capsule uses both radii, collision_system.cpp:212)"""
    from prototype2_b200 import capi
    scene = H.make_scene(36, 2, 2, 70, seed=81)
    caps = _random_capsules(70, 5)
    orc, ctx, tf, ph = _hetero_worlds(scene, caps, "l0", dict(variant="plain"), dict(order_mode=capi.ORDER_LITERAL))
    for f in range(15):
        orc.step(DT)
        ctx.update_characters(DT, tf, ph)
        assert H.state_diff(orc.get_state(), H.gpu_state(tf, ph)) == [], "frame %d" % f
    assert ctx.counters()["cc_tests"] > 0


@pytest.mark.parametrize("dyn", ["off", "jacobi"])
def test_heterogeneous_capsules_canonical(dyn):
    from prototype2_b200 import capi
    scene = H.make_scene(36, 2, 2, 200, seed=82)
    caps = _random_capsules(200, 6)
    orc, ctx, tf, ph = _hetero_worlds(scene, caps, "l1", dict(order="canonical", dyn=dyn),
                                      dict(dyn_mode=capi.DYN_JACOBI if dyn == "jacobi" else capi.DYN_OFF))
    for f in range(8):
        orc.step(DT)
        ctx.update_characters(DT, tf, ph)
        assert H.state_diff(orc.get_state(), H.gpu_state(tf, ph)) == [], "frame %d" % f


@pytest.mark.parametrize("substeps", [1, 2, 7])
def test_other_substep_counts(substeps):
    scene = H.make_scene(48, 2, 2, 300, seed=83)
    _run_pair(scene, 4, dict(order="canonical", substeps=substeps), dict(substeps=substeps))


def test_world_far_from_the_origin_disables_the_pre_cull():
    """coordinates around 3e6: the pre-cull's margin argument no longer holds and it switches itself off
    (StepParams::qr_enable); float spacing there is 0.25, results must still equal the oracle's bit for bit"""
    scene = H.make_scene(32, 2, 2, 120, seed=84)
    scene["model_transform"] = [3.0e6, 0.0, -2.0e6, 1, 0, 0, 0, 1, 1, 1]
    scene["pos"] = (scene["pos"] + np.array([3.0e6, 0.0, -2.0e6], np.float32)).astype(np.float32)
    _run_pair(scene, 4, dict(order="canonical"), dict())


def test_nan_and_inf_characters_do_not_hang_or_poison_the_others():
    """a NaN box overlaps every node (AABB::intersect is written with negated compares, aabb.cpp:3-10, trap T14), so
    such a character tests every triangle of the scene -- like in the reference -- and its neighbours are unaffected"""
    scene = H.make_scene(20, 2, 2, 48, seed=85)
    scene["pos"][3] = np.nan
    scene["vel"][7, 1] = np.inf
    scene["pos"][11, 0] = np.float32(3.0e38)
    _run_pair(scene, 3, dict(order="canonical"), dict(), trace_frames=())      # (a NaN query has more candidates than a trace slot holds)


def test_cross_frame_cache_does_not_care_who_sits_in_a_slot():
    """the persistent kernel's per-slot leaf cache survives between frames; it is a pure function of (tree, region), so
    shuffling the characters between frames (different character in every slot), teleporting some of them and moving
    the model (tree re-flattened) must change nothing -- checked against the oracle and against a context with the
    cache switched off"""
    from prototype2_b200 import capi
    n = 30000                                                 # more than 8 characters per resident warp: persistent kernel
    scene = H.make_scene(200, 2, 2, n, seed=91)
    scene["vel"] *= np.float32(2.0)
    orc = H.build_oracle(scene, "l1", order="canonical", threads=os.cpu_count() or 1)
    on, tf, ph = H.build_gpu(scene)
    off, tf2, ph2 = H.build_gpu(scene)
    off.set_option(capi.OPT_CROSS_FRAME_CACHE, 0)
    rs = np.random.RandomState(0)
    for f in range(8):
        if f in (2, 5):                                       # shuffle slots
            perm = rs.permutation(n)
            tf, ph, tf2, ph2 = tf[perm].copy(), ph[perm].copy(), tf2[perm].copy(), ph2[perm].copy()
            s = orc.get_state()
            orc.set_state(pos=s["pos"][perm], vel=s["vel"][perm], gnormal=s["gnormal"][perm], grounded=s["grounded"][perm],
                          rot=scene["rot"][perm])
            scene["rot"] = scene["rot"][perm]
        if f == 4:                                            # teleport a tenth of them
            idx = rs.choice(n, n // 10, replace=False)
            newpos = tf["position"].copy()
            newpos[idx, 0] = rs.uniform(4, 190, len(idx)).astype(np.float32)
            newpos[idx, 2] = rs.uniform(4, 190, len(idx)).astype(np.float32)
            tf["position"] = newpos
            tf2["position"] = newpos
            orc.set_state(pos=newpos)
        orc.step(DT)
        on.update_characters(DT, tf, ph)
        off.update_characters(DT, tf2, ph2)
        assert H.state_diff(H.gpu_state(tf, ph), H.gpu_state(tf2, ph2)) == [], "frame %d on/off" % f
        assert H.state_diff(orc.get_state(), H.gpu_state(tf, ph)) == [], "frame %d oracle" % f
    # a moved model re-flattens the tree: cached node indices must be dropped (tree_version)
    t = np.zeros(1, capi.TRANSFORM_DTYPE)
    t["position"][0] = (0.5, -0.25, 0.25); t["rotation"][0] = (0, 0, 0, 1); t["scale"][0] = (1, 1, 1)
    on.update_model_colliders([0], t)
    off.update_model_colliders([0], t)
    for f in range(3):
        on.update_characters(DT, tf, ph)
        off.update_characters(DT, tf2, ph2)
        assert H.state_diff(H.gpu_state(tf, ph), H.gpu_state(tf2, ph2)) == [], "after move, frame %d" % f
    assert on.counters()["launches"] == off.counters()["launches"] > 0


def test_warp_kernel_overflow_and_equivalence_with_the_lock_step_kernel():
    """small batches run one warp per character (k_step_warp).  A huge capsule and a NaN one overlap far more than
    the 128 leaves / frontier nodes a warp buffers -> ordered passes; results must equal the oracle and the thinned
    lock-step kernel (PRTC_OPT_WARP_KERNEL = 0), counters included"""
    from prototype2_b200 import capi
    xyz, mf, mc = H.scenes.terrain(32, 1, 1)                  # 1024 two-triangle leaves
    rs = np.random.RandomState(7)
    n = 24
    pos = np.stack([rs.uniform(4, 28, n), rs.uniform(0.3, 1.5, n), rs.uniform(4, 28, n)], 1).astype(np.float32)
    vel = rs.uniform(-0.1, 0.1, (n, 3)).astype(np.float32)
    pos[5] = np.nan
    caps = [(0.5, 0.5, (0.0, 0.5, 0.0))] * n
    caps[2] = (1.0, 6.0, (0.0, 6.0, 0.0))                     # 13 units wide: ~200 candidate leaves
    caps[9] = (0.2, 3.0, (0.0, 3.0, 0.0))
    scene = {"xyz": xyz, "mesh_first": mf, "mesh_count": mc, "model_transform": None, "pos": pos, "vel": vel,
             "rot": np.tile(np.array([1, 0, 0, 0], np.float32), (n, 1)), "move": np.zeros((n, 3), np.float32), "n": n}
    orc, warp, tf, ph = _hetero_worlds(scene, caps, "l1", dict(order="canonical"), dict())
    _, lock, tf2, ph2 = _hetero_worlds(scene, caps, "l1", dict(order="canonical"), dict())
    lock.set_option(capi.OPT_WARP_KERNEL, 0)
    _, walk, tf3, ph3 = _hetero_worlds(scene, caps, "l1", dict(order="canonical"), dict())
    walk.set_option(capi.OPT_FRAME_CACHE, 0)                  # a frontier walk per substep instead of the frame's cached leaves
    for f in range(4):
        orc.step(DT)
        warp.update_characters(DT, tf, ph)
        lock.update_characters(DT, tf2, ph2)
        walk.update_characters(DT, tf3, ph3)
        assert H.state_diff(orc.get_state(), H.gpu_state(tf, ph)) == [], "frame %d warp kernel vs oracle" % f
        assert H.state_diff(H.gpu_state(tf, ph), H.gpu_state(tf2, ph2)) == [], "frame %d warp vs lock-step" % f
        assert H.state_diff(H.gpu_state(tf, ph), H.gpu_state(tf3, ph3)) == [], "frame %d cached vs walked" % f
    co, cw, cl, ck = orc.counters(), warp.counters(), lock.counters(), walk.counters()
    for k in ("substeps", "tri_tests", "contacts"):
        assert co[k] == cw[k] == cl[k] == ck[k], k
    assert ck["node_tests"] == co["node_tests"]               # the per-substep frontier walk tests exactly the nodes the reference tests


def test_randomised_long_runs():
    """tests/parity_fuzz.py: 16 random levels / crowds / capsule shapes / rotations / inputs, 60-200 free-running frames
    each, CANONICAL vs the L1 port and LITERAL vs the reference's own code, bit for bit every frame"""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, PYTHONPATH=root + os.pathsep + os.environ.get("PYTHONPATH", ""))
    r = subprocess.run([sys.executable, os.path.join(root, "tests", "parity_fuzz.py"), "16", "11"], env=env, capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "16 rounds, 0 failures" in r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("cross_frame", [0, 1])
def test_top_staging_does_not_change_results(cross_frame):
    """PRTC_OPT_TOP_STAGING: k_xc_refresh walking the top levels out of shared memory (TMA bulk copy per block) fills the
    slots with the same leaf lists as the plain walk -- states and counters equal with the option on and off and equal to
    the oracle; with the cross-frame cache off every slot is refilled every frame, and teleports force refills with it on"""
    from prototype2_b200 import capi
    n = 40000
    scene = H.make_scene(160, 2, 2, n, seed=17)                # 6400 leaves, 12799 nodes: staged (>= 4096)
    scene["vel"] *= np.float32(1.5)
    orc = H.build_oracle(scene, "l1", order="canonical", threads=os.cpu_count() or 1)
    on, tf, ph = H.build_gpu(scene)
    off, tf2, ph2 = H.build_gpu(scene)
    off.set_option(capi.OPT_TOP_STAGING, 0)
    for c in (on, off):
        c.set_option(capi.OPT_CROSS_FRAME_CACHE, cross_frame)
    rs = np.random.RandomState(5)
    for f in range(6):
        if f in (2, 4):                                        # teleport a fifth; one character goes non-finite
            idx = rs.choice(n, n // 5, replace=False)
            newpos = tf["position"].copy()
            newpos[idx, 0] = rs.uniform(4, 150, len(idx)).astype(np.float32)
            newpos[idx, 2] = rs.uniform(4, 150, len(idx)).astype(np.float32)
            tf["position"] = newpos
            tf2["position"] = newpos
            orc.set_state(pos=newpos)
        orc.step(DT)
        on.update_characters(DT, tf, ph)
        off.update_characters(DT, tf2, ph2)
        assert H.state_diff(H.gpu_state(tf, ph), H.gpu_state(tf2, ph2)) == [], "frame %d on/off" % f
        assert H.state_diff(orc.get_state(), H.gpu_state(tf, ph)) == [], "frame %d oracle" % f
    ca, cb = on.counters(), off.counters()
    for k in ("substeps", "tri_tests", "node_tests", "contacts", "launches"):
        assert ca[k] == cb[k], k


# ------------------------------------------------------------------------------------------------------------
# PRTC_ORDER_BY_ID: the tree is built on the device; mesh candidates in ascending collider id.
# Oracle = L1 "by_id": the reference's own candidate set (its tree, its stored fat boxes), sorted by id.
# ------------------------------------------------------------------------------------------------------------
def _tf10(t):
    from prototype2_b200 import capi
    g = np.zeros(1, capi.TRANSFORM_DTYPE)
    g["position"][0] = t[0:3]; g["rotation"][0] = (t[4], t[5], t[6], t[3]); g["scale"][0] = t[7:10]
    return g


@pytest.mark.parametrize("n,dyn", [(500, "off"), (500, "jacobi"), (36000, "off")])
def test_by_id_order_device_built_tree_bit_exact(n, dyn):
    """two overlapping models (a terrain and a second, finer one 0.3 above part of it -- characters touch both, so the
    order of the candidates decides the result), one of them nudged inside its margin (stored fat boxes must survive,
    aabb_tree.cpp:102-111) and then moved for real; states every frame, traces (candidate ids in ascending order, hit
    indices, pushes) on some; box queries against brute force over the fat boxes"""
    from prototype2_b200 import capi
    scene = H.make_scene(96, 2, 2, n, seed=31)
    xyz2, mf2, mc2 = H.scenes.terrain(40, 1, 3)
    t2 = [20.0, 0.3, 25.0, 1.0, 0.0, 0.0, 0.0, 1.0, 1.0, 1.0]
    kw = dict(order="by_id", dyn=dyn)
    if n > 1000:
        kw["threads"] = os.cpu_count() or 1
    orc = H.build_oracle(scene, "l1", **kw)
    orc.add_model(xyz2, mf2, mc2, t2)
    ctx, tf, ph = H.build_gpu(scene, order_mode=capi.ORDER_BY_ID, dyn_mode=capi.DYN_JACOBI if dyn == "jacobi" else capi.DYN_OFF)
    m2, _ = ctx.add_model_collider(xyz2, mf2, mc2, t2)
    ctx.commit()
    n_leaves = len(scene["mesh_first"]) + len(mf2)
    assert np.array_equal(ctx.leaf_order(), np.arange(n_leaves, dtype=np.uint32))
    # the device-built tree answers box queries with exactly the leaves whose fat box (tight box -/+ 0.05) overlaps,
    # in ascending id
    wv = ctx.world_vertices()
    first = np.concatenate([scene["mesh_first"], np.asarray(mf2) + len(scene["xyz"])]).astype(np.int64)
    lo = np.minimum.reduceat(wv, first, axis=0) - np.float32(0.05)
    hi = np.maximum.reduceat(wv, first, axis=0) + np.float32(0.05)
    rs = np.random.RandomState(3)
    for q in range(40):
        c = rs.uniform(0, 96, 3).astype(np.float32); c[1] = rs.uniform(-1, 2)
        e = rs.uniform(0.3, 9.0, 3).astype(np.float32)
        qlo, qhi = c - e, c + e
        brute = np.nonzero(~((lo > qhi).any(axis=1) | (hi < qlo).any(axis=1)))[0]
        assert np.array_equal(ctx.query_box(np.concatenate([qlo, qhi])), brute.astype(np.uint32)), "query %d" % q
    keys = LITERAL_KEYS if dyn == "jacobi" else H.TRACE_KEYS
    for f in range(8):
        if f == 3:
            t2 = [20.01, 0.32, 25.0, 1.0, 0.0, 0.0, 0.0, 1.0, 1.0, 1.0]          # inside the margin
        if f == 5:
            t2 = [22.0, 0.1, 24.0, 0.9961947, 0.0, 0.0871557, 0.0, 1.0, 1.2, 1.0]
        if f in (3, 5):
            orc.update_model(1, t2)
            ctx.update_model_colliders([m2], _tf10(t2))
        if f in (0, 3, 5) and n <= 1000:
            to = orc.step(DT, trace=True)
            tg = ctx.update_characters_traced(DT, tf, ph)
            assert H.trace_diff(to, tg, keys) == [], "frame %d trace" % f
        else:
            orc.step(DT)
            ctx.update_characters(DT, tf, ph)
        assert H.state_diff(orc.get_state(), H.gpu_state(tf, ph)) == [], "frame %d state" % f
    co, cg = orc.counters(), ctx.counters()
    for k in ("substeps", "tri_tests", "contacts") + (("cc_tests",) if dyn == "jacobi" else ()):
        assert co[k] == cg[k], k
    # some queries did see both models at once
    assert co["tri_tests"] > 0


def test_by_id_order_removed_model_and_empty_world():
    """a removed model's leaves drop out of the device-built tree; an empty world steps like the reference (no candidates:
    `if (meshColIDs.empty()) break;`, physics_system.cpp:391-393)"""
    from prototype2_b200 import capi
    scene = H.make_scene(48, 2, 2, 300, seed=12)
    xyz2, mf2, mc2 = H.scenes.terrain(48, 2, 2)
    t2 = [0.0, 0.4, 0.0, 1.0, 0.0, 0.0, 0.0, 1.0, 1.0, 1.0]
    ctx, tf, ph = H.build_gpu(scene, order_mode=capi.ORDER_BY_ID)
    m2, _ = ctx.add_model_collider(xyz2, mf2, mc2, t2)
    ctx.update_characters(DT, tf, ph)                          # both present
    ctx.remove_model_collider(m2)
    ctx.commit()
    assert np.array_equal(ctx.leaf_order(), np.arange(len(scene["mesh_first"]), dtype=np.uint32))
    orc = H.build_oracle(scene, "l1", order="by_id")
    s = H.gpu_state(tf, ph)
    orc.set_state(pos=s["pos"], vel=s["vel"], gnormal=s["gnormal"], grounded=s["grounded"])
    for f in range(4):
        orc.step(DT)
        ctx.update_characters(DT, tf, ph)
        assert H.state_diff(orc.get_state(), H.gpu_state(tf, ph)) == [], "frame %d" % f    # exact: no tree history in this order
    ctx.remove_model_collider(0)
    ctx.commit()
    assert ctx.tree_info()["nodes"] == 0
    before = tf["position"].copy()
    ctx.update_characters(DT, tf, ph)
    assert np.isfinite(tf["position"]).all() and not np.array_equal(before, tf["position"])


@pytest.mark.parametrize("mode", [0, 1, 2, 3])
def test_shared_division_and_square_root_fast_paths_are_correctly_rounded(mode):
    """prt_math.cuh computes vec3 / float, the barycentric pair and the three edge lengths / inverse lengths of the exact
    test through written-out copies of the div.rn / sqrt.rn fast paths (shared reciprocal, interleaved chains).  On the
    device, against __fdiv_rn / __fsqrt_rn: 2^22 threads x 16 operand sets x 14 results per seed, bits must agree
    (random bit patterns incl. NaN / inf / denormals; magnitudes around 1; magnitudes at the edges of the fast ranges)"""
    from prototype2_b200 import capi
    ctx = capi.PhysicsContext()
    for seed in (1, 2, 3):
        assert ctx.selftest_math(1 << 22, seed, mode) == 0
