"""The pooled kernel (k_step_pool: 32 characters per warp, their candidate triangles tested by all lanes, speculative
ordered commit) against the CPU oracle, bit for bit.  Small scenes normally run on the warp-per-character kernel, so
these tests switch that one off (PRTC_OPT_WARP_KERNEL = 0): every untraced launch below is the pooled kernel, with 4, 8,
16 or 32 owners per warp depending on the batch size."""
import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu

DT = 1.0 / 30.0


def _pooled_ctx(scene, **kw):
    from prototype2_b200 import capi
    opts = kw.pop("opts", {})
    ctx, tf, ph = H.build_gpu(scene, **kw)
    ctx.set_option(capi.OPT_WARP_KERNEL, 0)
    for k, v in opts.items():
        ctx.set_option(getattr(capi, k), v)
    return ctx, tf, ph


def _run(scene, frames, oracle_kw=None, gpu_kw=None, counters=("substeps", "tri_tests", "contacts")):
    orc = H.build_oracle(scene, "l1", **dict(dict(order="canonical"), **(oracle_kw or {})))
    ctx, tf, ph = _pooled_ctx(scene, **(gpu_kw or {}))
    for f in range(frames):
        orc.step(DT)
        ctx.update_characters(DT, tf, ph)
        assert H.state_diff(orc.get_state(), H.gpu_state(tf, ph)) == [], "frame %d" % f
    co, cg = orc.counters(), ctx.counters()
    for k in counters:
        assert co[k] == cg[k], "counter %s: oracle %d gpu %d" % (k, co[k], cg[k])
    assert cg["exact_tests"] > 0
    return orc, ctx, tf, ph


@pytest.mark.parametrize("n", [37, 1000, 9000, 40000, 90000])
def test_pooled_kernel_all_owner_counts(n):
    """4 / 8 / 16 / 32 owners per warp (batch size decides), ragged last warp"""
    scene = H.make_scene(96, 2, 2, n, seed=100 + n % 7)
    _run(scene, 4)


@pytest.mark.parametrize("opts", [{"OPT_TIGHT_REJECT": 0}, {"OPT_QUICK_REJECT": 0}, {"OPT_QUICK_REJECT": 0, "OPT_TIGHT_REJECT": 0},
                                  {"OPT_FRAME_CACHE": 0}, {"OPT_CROSS_FRAME_CACHE": 0}, {"OPT_POOL_KERNEL": 0}])
def test_pooled_kernel_options_do_not_change_results(opts):
    scene = H.make_scene(96, 2, 2, 6000, seed=8)
    scene["vel"] = scene["vel"] * np.float32(1.7)           # some boxes escape the frame region
    extra = ("node_tests",) if opts.get("OPT_FRAME_CACHE") == 0 else ()
    _run(scene, 5, gpu_kw=dict(opts=opts), counters=("substeps", "tri_tests", "contacts") + extra)


@pytest.mark.parametrize("abs_mode", ["int", "float"])
def test_pooled_kernel_c2_shape_both_abs_modes(abs_mode):
    from prototype2_b200 import capi
    scene = H.make_scene(224, 1, 1, 5000, seed=3)
    _run(scene, 6, dict(abs_mode=abs_mode), dict(abs_mode=capi.ABS_INT if abs_mode == "int" else capi.ABS_FLOAT))


def test_pooled_kernel_rotated_capsules_16_triangle_leaves_three_substeps():
    scene = H.make_scene(120, 4, 2, 7000, seed=21)
    scene["rot"] = H.random_rotations(7000, seed=5)
    _run(scene, 5, dict(substeps=3), dict(substeps=3))


def test_pooled_kernel_crowd_in_a_corner_long_candidate_lists():
    """everybody over the same few cells, fast: windows fill the job list, several pushes per query, candidate leaves
    overflow one pass of the ordered walk (frame cache off for half of the frames)"""
    from prototype2_b200 import capi
    scene = H.make_scene(40, 1, 1, 4000, seed=31)
    rs = np.random.RandomState(4)
    scene["pos"][:, 0] = 10.0 + rs.uniform(0, 4, 4000).astype(np.float32)
    scene["pos"][:, 2] = 10.0 + rs.uniform(0, 4, 4000).astype(np.float32)
    scene["pos"][:, 1] = H.scenes.height(scene["pos"][:, 0], scene["pos"][:, 2]) + rs.uniform(-0.3, 0.5, 4000).astype(np.float32)
    scene["vel"] = (scene["vel"] * np.float32(3.0)).astype(np.float32)
    scene["capsule"] = (1.2, 0.9, (0.0, 0.9, 0.0))          # big capsule over 2-triangle leaves: > 12 candidate leaves
    orc = H.build_oracle(scene, "l1", order="canonical")
    ctx, tf, ph = _pooled_ctx(scene)
    for f in range(6):
        ctx.set_option(capi.OPT_FRAME_CACHE, f % 2)
        orc.step(DT)
        ctx.update_characters(DT, tf, ph)
        assert H.state_diff(orc.get_state(), H.gpu_state(tf, ph)) == [], "frame %d" % f
    co, cg = orc.counters(), ctx.counters()
    for k in ("substeps", "tri_tests", "contacts"):
        assert co[k] == cg[k], k


def test_pooled_kernel_movement_tunnelling_and_non_finite_characters():
    scene = H.make_scene(64, 2, 2, 5000, seed=41)
    rs = np.random.RandomState(6)
    scene["move"][:, 0] = rs.uniform(-0.2, 0.2, 5000).astype(np.float32)
    scene["move"][:, 2] = rs.uniform(-0.2, 0.2, 5000).astype(np.float32)
    scene["vel"][::7, 1] = np.float32(-2.5)                  # faster than a radius per frame: tunnels like the reference (T6)
    scene["pos"][::11, 1] += np.float32(30.0)                # far above the terrain: `if (meshColIDs.empty()) break;`
    scene["pos"][5] = np.nan
    scene["vel"][9, 1] = np.inf
    scene["pos"][13, 0] = np.float32(3.0e38)
    _run(scene, 5)


def test_pooled_kernel_degenerate_and_empty_meshes():
    """zero-area triangles are skipped but counted; an empty mesh is a candidate leaf without triangles"""
    scene = H.make_scene(48, 2, 2, 4000, seed=51)
    xyz = scene["xyz"].reshape(-1, 9).copy()
    xyz[::5, 3:6] = xyz[::5, 0:3]                            # every 5th triangle: b == a
    scene["xyz"] = xyz.reshape(-1)
    mc = scene["mesh_count"].copy()
    mc[::9] = 0                                              # every 9th mesh collider has no triangles at all
    scene["mesh_count"] = mc
    _run(scene, 4)


def test_pooled_kernel_heterogeneous_capsules():
    from prototype2_b200 import capi
    scene = H.make_scene(64, 2, 2, 3000, seed=61)
    rs = np.random.RandomState(7)
    caps = [(float(rs.uniform(0.2, 1.2)), float(rs.uniform(0.2, 0.8)), (float(rs.uniform(-0.1, 0.1)), float(rs.uniform(0.2, 0.8)), float(rs.uniform(-0.1, 0.1))))
            for _ in range(3000)]
    orc = H.L1World("canonical")
    orc.add_model(scene["xyz"], scene["mesh_first"], scene["mesh_count"], None)
    ctx = capi.PhysicsContext()
    ctx.add_model_collider(scene["xyz"], scene["mesh_first"], scene["mesh_count"], None)
    for h, r, off in caps:
        orc.add_capsule(h, r, off)
        ctx.add_capsule_collider(h, r, off)
    ctx.set_option(capi.OPT_WARP_KERNEL, 0)
    orc.set_state(pos=scene["pos"], rot=scene["rot"], vel=scene["vel"], move=scene["move"])
    tf = capi.make_transforms(scene["pos"])
    ph = capi.make_physics(scene["vel"], collider=np.arange(3000, dtype=np.uint32))
    for f in range(6):
        orc.step(DT)
        ctx.update_characters(DT, tf, ph)
        assert H.state_diff(orc.get_state(), H.gpu_state(tf, ph)) == [], "frame %d" % f


@pytest.mark.parametrize("n,terrain", [(700, 48), (12000, 96)])
def test_pooled_kernel_jacobi_dynamic_pass(n, terrain):
    """the dynamic-vs-dynamic pass inside the pooled kernel (big batches; the lock-step kernel keeps the traced launches)"""
    from prototype2_b200 import capi
    scene = H.make_scene(terrain, 2, 2, n, seed=55)
    orc = H.build_oracle(scene, "l1", order="canonical", dyn="jacobi")
    ctx, tf, ph = _pooled_ctx(scene, dyn_mode=capi.DYN_JACOBI)
    for f in range(6):
        orc.step(DT)
        ctx.update_characters(DT, tf, ph)
        assert H.state_diff(orc.get_state(), H.gpu_state(tf, ph)) == [], "frame %d" % f
    co, cg = orc.counters(), ctx.counters()
    for k in ("substeps", "tri_tests", "contacts", "cc_tests"):
        assert co[k] == cg[k], k
    assert cg["cc_tests"] > 0


def test_pooled_kernel_jacobi_pile_more_neighbours_than_the_buffer():
    from prototype2_b200 import capi
    scene = H.make_scene(32, 2, 2, 90, seed=56)
    rs = np.random.RandomState(2)
    scene["pos"][:, 0] = 12.0 + rs.uniform(0, 6, 90).astype(np.float32)
    scene["pos"][:, 2] = 12.0 + rs.uniform(0, 6, 90).astype(np.float32)
    scene["pos"][:, 1] = H.scenes.height(scene["pos"][:, 0], scene["pos"][:, 2]) + rs.uniform(0, 2, 90).astype(np.float32)
    orc = H.build_oracle(scene, "l1", order="canonical", dyn="jacobi")
    ctx, tf, ph = _pooled_ctx(scene, dyn_mode=capi.DYN_JACOBI)
    for f in range(5):
        orc.step(DT)
        ctx.update_characters(DT, tf, ph)
        assert H.state_diff(orc.get_state(), H.gpu_state(tf, ph)) == [], "frame %d" % f
    assert ctx.counters()["cc_tests"] == orc.counters()["cc_tests"]


def test_pooled_kernel_by_id_order():
    from prototype2_b200 import capi
    scene = H.make_scene(96, 2, 2, 8000, seed=71)
    _run(scene, 4, dict(order="by_id"), dict(order_mode=capi.ORDER_BY_ID))


def test_records_edited_between_frames():
    """the per-slot leaf lists are kept across frames for the region a character was last seen in; the caller may edit the
    records in between: a third of a walking crowd is teleported (and turned around) after frame 2 and again after frame 4,
    so their slots' regions are wrong and the refresh / the per-substep region check must catch it"""
    scene = H.make_scene(96, 2, 2, 9000, seed=21)
    rs = np.random.RandomState(5)
    ang = rs.uniform(-np.pi, np.pi, scene["n"])
    scene["move"] = np.stack([0.08 * np.cos(ang), np.zeros(scene["n"]), 0.08 * np.sin(ang)], axis=1).astype(np.float32)
    orc = H.build_oracle(scene, "l1", order="canonical")
    ctx, tf, ph = _pooled_ctx(scene)
    for f in range(7):
        if f in (3, 5):
            st = orc.get_state()
            who = rs.choice(scene["n"], scene["n"] // 3, replace=False)
            dst = np.roll(who, 17)
            pos, vel = st["pos"].copy(), st["vel"].copy()
            pos[who] = st["pos"][dst] + np.float32(0.3)
            vel[who] = -st["vel"][dst]
            orc.set_state(pos=pos, vel=vel)
            tf["position"][:] = pos
            ph["velocity"][:] = vel
        orc.step(DT)
        ctx.update_characters(DT, tf, ph)
        assert H.state_diff(orc.get_state(), H.gpu_state(tf, ph)) == [], "frame %d" % f
    co, cg = orc.counters(), ctx.counters()
    for k in ("substeps", "tri_tests", "contacts"):
        assert co[k] == cg[k], k
