"""Pins the oracle: the L1 restatement (oracle/l1_oracle.cpp) in LITERAL mode must be BIT-identical -- states and
traces (ordered candidate lists, hit indices, push vectors, node/triangle test counts) -- to L0, the reference's own
collision sources compiled verbatim (oracle/_ref, built from /root/reference by oracle/Makefile), in both
resolutions of the abs() trap; and both must reproduce the committed golden vectors."""
import glob
import os

import numpy as np
import pytest

import helpers as H
from oraclelib import L0World, L1World, Trace

DT = 1.0 / 30.0
GOLDEN = sorted(p for p in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")) if not os.path.basename(p).startswith("c1_"))
C1_PATH = os.path.join(os.path.dirname(__file__), "golden", "c1_level.npz")
needs_l0 = pytest.mark.skipif(not L0World.available("traced"), reason="oracle/_ref not built (needs /root/reference)")


def _lockstep(scene, frames, variant, abs_mode, traced):
    l0 = H.build_oracle(scene, "l0", variant=variant)
    l1 = H.build_oracle(scene, "l1", order="literal", abs_mode=abs_mode)
    assert H.same(l0.dump_order()[0], l1.dump_order()[0]) and H.same(l0.dump_order()[1], l1.dump_order()[1])
    cc_hits = 0
    for f in range(frames):
        t0 = l0.step(DT, trace=traced)
        t1 = l1.step(DT, trace=traced)
        assert H.state_diff(l0.get_state(), l1.get_state()) == [], "frame %d" % f
        if traced:
            assert H.trace_diff(t0, t1, Trace.FIELDS) == [], "frame %d" % f
            cc_hits += int((t0.h_poly < 0).sum())
        assert H.same(l0.dump_order()[0], l1.dump_order()[0]), "tree order diverged at frame %d" % f
    return cc_hits


@needs_l0
@pytest.mark.parametrize("variant,abs_mode,traced", [("traced", "int", True), ("plain", "int", False), ("fabs", "float", False)])
def test_l1_literal_equals_reference_crowded(variant, abs_mode, traced):
    scene = H.make_scene(40, 2, 1, 80, seed=101)            # random placement: capsule-capsule contacts occur
    cc = _lockstep(scene, 25, variant, abs_mode, traced)
    if traced:
        assert cc > 0, "scene was meant to exercise collideCapsuleCapsule"


@needs_l0
def test_l1_literal_equals_reference_rotated_and_moving():
    scene = H.make_scene(48, 1, 1, 60, seed=102)
    scene["rot"] = H.random_rotations(60, seed=9)
    rs = np.random.RandomState(3)
    scene["move"] = rs.uniform(-0.2, 0.2, (60, 3)).astype(np.float32)
    _lockstep(scene, 20, "traced", "int", True)


@needs_l0
def test_l1_literal_equals_reference_transformed_model_16tri_leaves():
    scene = H.make_scene(64, 4, 2, 50, seed=103, separated=True)
    scene["model_transform"] = [0.5, 0.25, -1.0, 0.9914449, 0.0, 0.1305262, 0.0, 1.25, 1.0, 1.25]
    _lockstep(scene, 10, "traced", "int", True)


@needs_l0
def test_abs_trap_is_real():
    """`abs(dist)` (collision_system.cpp:72) truncates to int under libstdc++: a capsule hovering 0.7 above a floor
    inside a tall leaf box is 'grounded' there and is not with the float overload."""
    g = np.load([p for p in GOLDEN if p.endswith("abs_trap_hover.npz")][0])
    assert g["int_grounded"][0][0] == 1 and g["flt_grounded"][0][0] == 0


def _scene_from_golden(g):
    cap = g["capsule"]
    return {"xyz": g["xyz"], "mesh_first": g["mesh_first"], "mesh_count": g["mesh_count"],
            "model_transform": list(g["model_transform"]), "capsule": (float(cap[0]), float(cap[1]), tuple(float(x) for x in cap[2:5])),
            "pos": g["pos"], "vel": g["vel"], "rot": g["rot"], "move": g["move"], "n": len(g["pos"])}


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_l1_reproduces_golden(path):
    g = np.load(path)
    scene = _scene_from_golden(g)
    for tag, absm in (("int", "int"), ("flt", "float")):
        w = H.build_oracle(scene, "l1", order="literal", abs_mode=absm)
        for f in range(int(g["frames"])):
            tr = w.step(float(g["dt"]), trace=(f == 0 and tag == "int"))
            s = w.get_state()
            for k in ("pos", "vel", "gnormal", "grounded"):
                assert H.same(s[k], g["%s_%s" % (tag, k)][f]), "%s frame %d %s" % (tag, f, k)
            if tr is not None:
                for k in Trace.FIELDS:
                    assert H.same(getattr(tr, k), g["trace_" + k]), k


@needs_l0
@pytest.mark.parametrize("path", GOLDEN[:4], ids=[os.path.basename(p)[:-4] for p in GOLDEN[:4]])
def test_reference_build_still_reproduces_golden(path):
    """guards the L0 build recipe itself (flags, shim headers)"""
    g = np.load(path)
    scene = _scene_from_golden(g)
    w = H.build_oracle(scene, "l0", variant="plain")
    for f in range(int(g["frames"])):
        w.step(float(g["dt"]))
        assert H.same(w.get_state()["pos"], g["int_pos"][f])


def test_known_answers():
    """physical sanity of the golden vectors themselves (what the survey observed with the literal oracle)"""
    g = {os.path.basename(p)[:-4]: np.load(p) for p in GOLDEN}
    fd = g["floor_drop"]
    assert abs(fd["int_pos"][-1][0][1]) < 1e-5 and fd["int_grounded"][-1][0] == 1     # rests on y = 0
    assert np.allclose(fd["int_gnormal"][-1][0], [0, 1, 0])
    assert fd["int_vel"][-1][0][1] == 0.0
    assert g["tunnel_fast_body"]["int_pos"][-1][0][1] < -5.0                          # T6: falls through
    assert g["backface_cull"]["int_grounded"].sum() == 0                              # back faces never collide
    ws = g["wall_slide"]["int_pos"][:, 0]
    assert ws[:, 0].max() < 4.51 and ws[-1, 2] > ws[0, 2]                             # stopped by the wall (x=5-r), slides in z


@needs_l0
def test_canonical_vs_literal_deviation_report():
    """P3: canonical order (static-only tree) against the literal unified tree.  Sets always agree; sequences may
    not (finding 3/H4); characters whose sequences agree are bit-identical."""
    scene = H.make_scene(120, 2, 2, 600, seed=21, separated=True)
    l0 = H.build_oracle(scene, "l0", variant="traced")
    l1 = H.build_oracle(scene, "l1", order="canonical")
    for f in range(3):
        t0, t1 = l0.step(DT, trace=True), l1.step(DT, trace=True)
        assert len(t0.caps_ids) == 0
        st = H.compare_with_literal(t0, t1, l0.get_state(), l1.get_state(), scene["n"])
        assert st["chars_in_sync"] >= scene["n"] * 0.6
        s = l0.get_state()
        l1.set_state(pos=s["pos"], vel=s["vel"], gnormal=s["gnormal"], grounded=s["grounded"])


@pytest.mark.parametrize("tag,absm", [("int", "int"), ("flt", "float")])
def test_l1_reproduces_c1_shipped_level(tag, absm):
    """C1: the repo's shipped level (res/models/level1, 3007 triangles in 172 mesh colliders) with the player capsule
    of res/scenes/cavern.prt, 300 scripted frames (fall, land, run, jump) recorded from the reference's own code"""
    g = np.load(C1_PATH)
    w = H.build_oracle(H.c1_scene(g), "l1", order="literal", abs_mode=absm)
    for f in range(int(g["frames"])):
        s = w.get_state()
        vel = s["vel"].copy()
        if g["jump_script"][f]:
            vel[0, 1] = np.float32(0.35)
        w.set_state(vel=vel, move=g["move_script"][f:f + 1])
        w.step(float(g["dt"]))
        s = w.get_state()
        for k in ("pos", "vel", "gnormal", "grounded"):
            assert H.same(s[k], g["%s_%s" % (tag, k)][f]), "frame %d %s" % (f, k)
    assert g["int_grounded"].sum() > 200 and abs(float(g["int_pos"][-1][0][1])) < 1e-5    # ends standing on the ground floor


def test_sweep_checker_known_answers():
    """Tier-B checker (oracle/l0_sweep_driver.cpp: the reference's Nettle helpers under our loop; parity unpinned):
    an ellipsoid dropped on a floor stops where its lowest point touches, keeps the tangential part of its motion,
    and reports the impact at the fraction of the motion that geometry dictates."""
    import oraclelib as O
    if not O.sweep_reference_available():
        pytest.skip("oracle/_ref/libprt_l0_sweep.so not built")
    floor = np.array([[-10, 0, -10], [-10, 0, 10], [10, 0, 10], [-10, 0, -10], [10, 0, 10], [10, 0, -10]], np.float32)
    r = O.sweep_reference(floor, [[0, 2, 0], [0, 0.8, 0], [3, 5, 1], [0, 2, 0]], [[0.5, 1.0, 0.5]] * 4,
                          [[0.3, -2.0, 0.0], [0.5, -0.5, 0], [0, -0.1, 0], [0, 1, 0]], iterations=3, very_close=0.005)
    # 1: bottom of the ellipsoid (y = 1) reaches the floor after half of the motion; slides the rest of the way in x
    assert r["hit_tri"][0].tolist()[0] in (0, 1) and r["hit_tri"][0, 1] == -1
    assert abs(r["hit_toi"][0, 0] - 0.5) < 1e-6 and abs(r["pos"][0, 1] - 1.005) < 1e-3 and abs(r["pos"][0, 0] - 0.3) < 2e-3
    assert abs(r["hit_normal"][0, 0, 1] - 1.0) < 1e-5
    # 2: starts cutting the floor: impact at time 0, vertical motion removed, tangential motion kept
    assert r["hit_toi"][1, 0] == 0.0 and np.allclose(r["pos"][1], [0.5, 0.8, 0.0], atol=1e-6)
    # 3: never reaches the floor; 4: moves away from it
    assert (r["hit_tri"][2:] == -1).all() and np.allclose(r["pos"][2], [3, 4.9, 1]) and np.allclose(r["pos"][3], [0, 3, 0])
    assert np.all(r["vel"] == 0.0)


def test_l1_by_id_is_canonical_with_sorted_candidates():
    """checker of PRTC_ORDER_BY_ID: on the first frame (same start state) every query of the by_id arrangement has the
    candidate SET of the canonical one -- the reference's own tree query -- in ascending collider id"""
    import helpers as H
    scene = H.make_scene(48, 1, 3, 120, seed=5)
    a = H.build_oracle(scene, "l1", order="canonical")
    b = H.build_oracle(scene, "l1", order="by_id")
    ta, tb = a.step(1.0 / 30.0, trace=True), b.step(1.0 / 30.0, trace=True)
    assert np.array_equal(ta.q_char, tb.q_char) or len(ta.q_char) > 0
    # first substep of every character: identical query boxes in both arrangements
    firsts_a = [i for i in range(len(ta.q_char)) if ta.q_sub[i] == 0]
    firsts_b = [i for i in range(len(tb.q_char)) if tb.q_sub[i] == 0]
    assert len(firsts_a) == len(firsts_b) == 120
    differ = 0
    for ia, ib in zip(firsts_a, firsts_b):
        la = ta.mesh_ids[(ta.q_mesh_off[ia - 1] if ia else 0):ta.q_mesh_off[ia]]
        lb = tb.mesh_ids[(tb.q_mesh_off[ib - 1] if ib else 0):tb.q_mesh_off[ib]]
        assert list(lb) == sorted(la)
        differ += list(la) != list(lb)
    assert differ > 0                      # the tree's own order is not the id order, or this test shows nothing


@needs_l0
def test_l1_update_model_equals_reference():
    """l1_update_model (physics_system.cpp:161-202): a model nudged inside the leaf margin and then moved for real, in
    lock step with the reference's own updateModelColliders -- states, traces and tree order stay bit-identical"""
    scene = H.make_scene(32, 2, 2, 40, seed=104)
    l0 = H.build_oracle(scene, "l0", variant="traced")
    l1 = H.build_oracle(scene, "l1", order="literal")
    for f in range(16):
        if f in (3, 6, 11):
            t = {3: [0.01, 0.02, 0.0, 1.0, 0.0, 0.0, 0.0, 1.0, 1.0, 1.0],
                 6: [0.6, -0.1, 0.3, 0.9961947, 0.0, 0.0871557, 0.0, 1.0, 1.1, 1.0],
                 11: [0.6, -0.12, 0.31, 0.9961947, 0.0, 0.0871557, 0.0, 1.0, 1.1, 1.0]}[f]
            l0.update_model(0, t)
            l1.update_model(0, t)
        t0, t1 = l0.step(DT, trace=True), l1.step(DT, trace=True)
        assert H.trace_diff(t0, t1, Trace.FIELDS) == [], "frame %d" % f
        assert H.state_diff(l0.get_state(), l1.get_state()) == [], "frame %d" % f
        assert H.same(l0.dump_order()[0], l1.dump_order()[0]), "tree order diverged at frame %d" % f


@pytest.mark.skipif(not L0World.available("plain"), reason="oracle/_ref not built")
def test_l1_remove_and_re_add_equals_the_reference():
    """removeCollider + addModelCollider (physics_system.cpp:108-149, 44-106): the restatement goes through the same tree
    removals and insertions as the reference's own code; one-mesh models removed last-first (the envelope in which the
    reference is well defined, see oracle/l0_driver.cpp l0_remove_model)"""
    def patch(x0, z0, y, n=6):
        xyz, mf, mc = H.scenes.terrain(n, n, n)
        xyz = xyz.copy()
        xyz[:, 1] = np.float32(y); xyz[:, 0] += np.float32(x0); xyz[:, 2] += np.float32(z0)
        return xyz, mf, mc
    l0, l1 = L0World("plain"), L1World("literal")
    ids = []
    for pch in (patch(0, 0, 0.0), patch(6, 0, 0.25), patch(0, 6, -0.25)):
        assert l0.add_model(*pch) == len(ids)
        ids.append(l1.add_model(*pch))
    n = 50
    rs = np.random.RandomState(11)
    pos = np.stack([rs.uniform(0.8, 11.2, n), rs.uniform(0.8, 1.6, n), rs.uniform(0.8, 11.2, n)], axis=1).astype(np.float32)
    vel = rs.uniform(-0.05, 0.05, (n, 3)).astype(np.float32)
    for w in (l0, l1):
        for _ in range(n):
            w.add_capsule()
        w.set_state(pos=pos, vel=vel)

    def frames(k, what):
        for f in range(k):
            l0.step(1.0 / 30.0)
            l1.step(1.0 / 30.0)
            assert H.state_diff(l0.get_state(), l1.get_state()) == [], "%s, frame %d" % (what, f)
            assert np.array_equal(l0.dump_order()[1], l1.dump_order()[1]), what      # capsule leaves in the same DFS order
    frames(5, "three models")
    l0.remove_model(2); l1.remove_model(ids[2])
    frames(5, "last removed")
    lower = patch(0, 6, -0.75)
    assert l0.add_model(*lower) == 2
    ids[2] = l1.add_model(*lower)
    frames(5, "re-added")
    l0.remove_model(2); l1.remove_model(ids[2]); l0.remove_model(1); l1.remove_model(ids[1])
    frames(4, "two removed")
    assert l0.add_model(*patch(6, 0, -0.5)) == 1
    ids[1] = l1.add_model(*patch(6, 0, -0.5))
    frames(6, "one re-added")


def test_l1_jacobi_grid_and_threads_equal_the_plain_loop(monkeypatch):
    """the restatement's JACOBI neighbour search uses a uniform grid and host threads from 2048 capsules on (so that config
    5 can be checked at full size); both are pure acceleration: states, counters and the ordered capsule-candidate lists
    of the trace equal the single-threaded loop over all pairs -- crowded scene, NaN and far-away capsules included"""
    scene = H.make_scene(48, 2, 2, 2600, seed=21)
    scene["pos"][7] = np.nan
    scene["pos"][11, 0] = np.float32(3.0e8)
    scene["vel"][13, 2] = np.inf
    monkeypatch.setenv("PRT_L1_NO_GRID", "1")
    plain = H.build_oracle(scene, "l1", order="canonical", dyn="jacobi", threads=1)
    st, tr = [], []
    for f in range(4):
        tr.append(plain.step(1.0 / 30.0, trace=(f == 2)))
        st.append(plain.get_state())
    monkeypatch.delenv("PRT_L1_NO_GRID")
    fast = H.build_oracle(scene, "l1", order="canonical", dyn="jacobi", threads=4)
    for f in range(4):
        t = fast.step(1.0 / 30.0, trace=(f == 2))
        assert H.state_diff(st[f], fast.get_state()) == [], "frame %d" % f
        if f == 2:
            assert H.trace_diff(tr[f], t, keys=H.TRACE_KEYS + ("q_caps_off", "caps_ids")) == []
    assert plain.counters() == fast.counters() and plain.counters()["cc_tests"] > 1000
