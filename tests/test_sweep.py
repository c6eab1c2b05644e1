"""TIER B -- PARITY UNPINNED: the swept ellipsoid-vs-triangle collide-and-slide kernel (prtc_sweep_ellipsoids).

The reference snapshot has no caller for this algorithm any more, only its helper functions
(src/util/physics_util.cpp:51-178).  The checker (oracle/l0_sweep_driver.cpp -> oracle/_ref/libprt_l0_sweep.so) runs
the REFERENCE's helper functions, compiled unmodified, under our restatement of Nettle's loop, brute force over every
triangle; the kernel must reproduce it bit for bit: hit triangle per iteration, time of impact, slide normal, final
position and remaining motion."""
import numpy as np
import pytest

import helpers as H
import oraclelib as O
from prototype2_b200 import capi

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not O.sweep_reference_available(), reason="oracle/_ref/libprt_l0_sweep.so not built")]


def _world(n_terrain, bx, bz, transform=None):
    xyz, mf, mc = H.scenes.terrain(n_terrain, bx, bz)
    ctx = capi.PhysicsContext()
    ctx.add_model_collider(xyz, mf, mc, transform)
    ctx.commit()
    return ctx, ctx.world_vertices()[: len(xyz)]


def _crowd(rs, n, lo, hi, speed=1.5):
    c = np.stack([rs.uniform(lo, hi, n), rs.uniform(0.3, 3.0, n), rs.uniform(lo, hi, n)], 1).astype(np.float32)
    r = np.stack([rs.uniform(0.3, 0.8, n), rs.uniform(0.4, 1.2, n), rs.uniform(0.3, 0.8, n)], 1).astype(np.float32)
    v = np.stack([rs.uniform(-speed, speed, n), rs.uniform(-2.0, 0.2, n), rs.uniform(-speed, speed, n)], 1).astype(np.float32)
    return c, r, v


def _same(got, want):
    assert np.array_equal(got["hit_tri"], want["hit_tri"]), "hit triangles"
    for k in ("hit_toi", "hit_normal", "pos", "vel"):
        assert H.same(got[k], want[k]), k


def test_sweep_matches_the_checker_on_a_terrain():
    ctx, world = _world(40, 2, 2)
    rs = np.random.RandomState(1)
    c, r, v = _crowd(rs, 600, 3, 37)
    got = ctx.sweep_ellipsoids(c, r, v, iterations=3)
    want = O.sweep_reference(world, c, r, v, iterations=3)
    _same(got, want)
    hits = (want["hit_tri"] >= 0)
    assert hits[:, 0].mean() > 0.5 and hits[:, 1].sum() > 20          # plenty of first impacts, some second ones after the slide
    assert np.all((want["hit_toi"] >= 0) & (want["hit_toi"] <= 1))
    assert ctx.counters()["tri_tests"] < want["tri_tests"] / 20        # the tree did prune


def test_sweep_under_a_model_transform_and_five_iterations():
    t = [1.0, -0.5, 2.0, 0.9914449, 0.0, 0.1305262, 0.0, 1.5, 0.8, 1.25]
    ctx, world = _world(24, 4, 2, t)
    rs = np.random.RandomState(2)
    c, r, v = _crowd(rs, 300, 4, 28, speed=3.0)
    got = ctx.sweep_ellipsoids(c, r, v, iterations=5, very_close=0.001)
    want = O.sweep_reference(world, c, r, v, iterations=5, very_close=0.001)
    _same(got, want)


def test_sweep_spheres_embedded_starts_and_no_motion():
    ctx, world = _world(16, 1, 1)
    rs = np.random.RandomState(3)
    c, r, v = _crowd(rs, 200, 2, 14)
    r[:] = 0.5                                                      # spheres
    c[:50, 1] = H.scenes.height(c[:50, 0], c[:50, 2]) + 0.3         # start cutting the surface (plane inside the sphere)
    v[50:60] = 0.0                                                  # no motion: no iteration runs
    v[60:70] = [0.0, 2.0, 0.0]                                      # moving away from every front face
    got = ctx.sweep_ellipsoids(c, r, v, iterations=3)
    want = O.sweep_reference(world, c, r, v, iterations=3)
    _same(got, want)
    assert np.all(want["hit_tri"][50:70] == -1) and H.same(want["pos"][50:60], c[50:60])
    assert (want["hit_toi"][:50, 0] == 0).sum() > 10                # embedded starts report an impact at time 0


def test_sweep_candidate_overflow_takes_the_ordered_passes():
    ctx, world = _world(48, 1, 1)                                   # 2304 leaves; a fat, fast ellipsoid overlaps > 128 of them
    c = np.float32([[24, 6, 24], [10, 5, 30], [24, 3, 24]])
    r = np.float32([[9, 3, 9], [6, 2, 7], [0.5, 0.5, 0.5]])
    v = np.float32([[3, -6, 2], [-2, -5, 1], [20, -1, 15]])
    got = ctx.sweep_ellipsoids(c, r, v, iterations=4)
    want = O.sweep_reference(world, c, r, v, iterations=4)
    _same(got, want)
    assert (want["hit_tri"][:2, 0] >= 0).all()


def test_sweep_argument_checks():
    ctx, _ = _world(8, 2, 2)
    with pytest.raises(capi.PrtcError):
        ctx.sweep_ellipsoids([[0, 1, 0]], [[0.5, 0.0, 0.5]], [[0, -1, 0]])          # zero radius component
    lit = capi.PhysicsContext(order_mode=capi.ORDER_LITERAL)
    xyz, mf, mc = H.scenes.terrain(8, 2, 2)
    lit.add_model_collider(xyz, mf, mc)
    with pytest.raises(capi.PrtcError):
        lit.sweep_ellipsoids([[0, 1, 0]], [[0.5, 0.5, 0.5]], [[0, -1, 0]])
