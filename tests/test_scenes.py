import numpy as np

from prototype2_b200 import scenes


def test_terrain_shapes_match_the_named_configs():
    xyz, mf, mc = scenes.terrain(224, 1, 1)
    assert len(xyz) == 3 * 100352 and len(mf) == 50176 and (mc == 6).all()          # C2
    xyz, mf, mc = scenes.terrain(64, 4, 2)
    assert (mc == 48).all() and len(mf) == 16 * 32                                    # C3 shape: 16-triangle leaves
    assert mf[0] == 0 and np.array_equal(mf[1:], np.cumsum(mc)[:-1])


def test_terrain_faces_up_and_is_watertight_in_order():
    xyz, mf, mc = scenes.terrain(20, 2, 2)
    t = xyz.reshape(-1, 3, 3)
    n = np.cross(t[:, 1] - t[:, 0], t[:, 2] - t[:, 0])
    assert (n[:, 1] > 0).all()                                                        # back faces would be culled
    assert np.allclose(t[:, :, 1], scenes.height(t[:, :, 0], t[:, :, 2]))


def test_ragged_edge_blocks():
    xyz, mf, mc = scenes.terrain(7, 2, 3)
    assert mc.sum() == 6 * 49 and mc.min() == 6 and mc.max() == 36


def test_capsule_placement_is_deterministic_and_in_range():
    p1, v1 = scenes.capsules(1000, 224)
    p2, v2 = scenes.capsules(1000, 224)
    assert np.array_equal(p1, p2) and np.array_equal(v1, v2)
    assert p1[:, 0].min() >= 2 and p1[:, 0].max() <= 222 and np.abs(v1).max() <= 0.2 + 1e-6
    pg, _ = scenes.capsules_grid(500, 120)
    d = pg[:, None, [0, 2]] - pg[None, :, [0, 2]]
    sep = np.abs(d).max(axis=2) + np.eye(500) * 10
    assert sep.min() >= 1.9                                                           # fat capsule boxes never overlap


def test_morton_order_is_a_permutation():
    p, _ = scenes.capsules(4096, 500)
    o = scenes.morton_order(p, 500)
    assert np.array_equal(np.sort(o), np.arange(4096))
    d_sorted = np.linalg.norm(np.diff(p[o][:, [0, 2]], axis=0), axis=1).mean()
    d_raw = np.linalg.norm(np.diff(p[:, [0, 2]], axis=0), axis=1).mean()
    assert d_sorted < d_raw / 5
