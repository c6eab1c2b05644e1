"""Host logic of the product (no GPU): the tree builder of prototype2_b200/csrc/bvh_builder.cpp must reproduce the
reference's tree -- checked through the DFS leaf order -- and the flattened right-first preorder layout with skip
links must enumerate exactly the leaves the reference's stack traversal finds, in the same order."""
import numpy as np
import pytest

import helpers as H
from oraclelib import L0World, L1World
from prototype2_b200 import capi, scenes


def mesh_boxes(xyz, mf, mc):
    v = xyz.reshape(-1, 3)
    lo = np.minimum.reduceat(v, mf, axis=0)
    hi = np.maximum.reduceat(v, mf, axis=0)
    return np.concatenate([lo, hi], axis=1).astype(np.float32)


def walk_flat(nodes, box):
    """numpy restatement of the device traversal (kernels.cu): front-to-back walk with skip links"""
    links = nodes[:, 3].view(np.int32) if nodes.dtype == np.float32 else nodes[:, 3]
    counts = nodes[:, 7].view(np.int32)
    out, i, n, tested = [], 0, len(nodes), 0
    while i < n:
        lo, hi = nodes[i, 0:3], nodes[i, 4:7]
        tested += 1
        ov = not (lo[0] > box[3] or hi[0] < box[0] or lo[1] > box[4] or hi[1] < box[1] or lo[2] > box[5] or hi[2] < box[2])
        if counts[i] < 0:
            i = i + 1 if ov else int(links[i])
        else:
            if ov:
                out.append(int(links[i]))
            i += 1
    return np.array(out, np.uint32), tested


@pytest.mark.parametrize("n,bx,bz", [(32, 1, 1), (64, 2, 2), (96, 4, 2), (33, 3, 5)])
def test_leaf_order_equals_oracle(n, bx, bz):
    xyz, mf, mc = scenes.terrain(n, bx, bz)
    order, nodes = capi.host_build_tree(mesh_boxes(xyz, mf, mc))
    l1 = L1World("canonical")
    l1.add_model(xyz, mf, mc)
    assert np.array_equal(order, l1.dump_order()[0])
    assert len(nodes) == 2 * len(mf) - 1


@pytest.mark.skipif(not L0World.available("plain"), reason="oracle/_ref not built")
def test_leaf_order_equals_reference_random_boxes():
    rs = np.random.RandomState(5)
    n = 20000
    c = rs.uniform(0, 200, (n, 3)).astype(np.float32)
    e = rs.uniform(0.1, 3.0, (n, 3)).astype(np.float32)
    boxes = np.concatenate([c - e, c + e], axis=1).astype(np.float32)
    order, nodes = capi.host_build_tree(boxes)
    # feed the same boxes to the reference as degenerate-free single-triangle meshes spanning each box
    tri = np.stack([boxes[:, 0:3], np.stack([boxes[:, 3], boxes[:, 1], boxes[:, 2]], 1), boxes[:, 3:6]], axis=1).reshape(-1, 3)
    l0 = L0World("plain")
    l0.add_model(tri, np.arange(n, dtype=np.uint32) * 3, np.full(n, 3, np.uint32))
    assert np.array_equal(order, l0.dump_order()[0])


def test_flat_walk_equals_stack_traversal():
    xyz, mf, mc = scenes.terrain(48, 2, 1)
    boxes = mesh_boxes(xyz, mf, mc)
    order, nodes = capi.host_build_tree(boxes)
    l1 = L1World("canonical")
    l1.add_model(xyz, mf, mc)
    rs = np.random.RandomState(2)
    fat = boxes.copy()
    fat[:, :3] -= np.float32(0.05)
    fat[:, 3:] += np.float32(0.05)
    for _ in range(200):
        c = rs.uniform(0, 48, 3).astype(np.float32)
        c[1] = rs.uniform(-1, 1)
        e = rs.uniform(0.2, 3.0, 3).astype(np.float32)
        q = np.concatenate([c - e, c + e]).astype(np.float32)
        got, tested = walk_flat(nodes, q)
        want, _ = l1.query(0xFFFFFFFF, q)
        assert np.array_equal(got, want)                       # same leaves, same ORDER
        brute = np.nonzero(~((fat[:, 0] > q[3]) | (fat[:, 3] < q[0]) | (fat[:, 1] > q[4]) | (fat[:, 4] < q[1]) |
                             (fat[:, 2] > q[5]) | (fat[:, 5] < q[2])))[0]
        assert np.array_equal(np.sort(got), brute)             # candidate SET == brute-force overlap of the fat boxes
    # everything-box enumerates every leaf in DFS order; nothing-box tests exactly one node
    assert np.array_equal(walk_flat(nodes, np.array([-1e9] * 3 + [1e9] * 3, np.float32))[0], order)
    assert walk_flat(nodes, np.array([1e8] * 3 + [1e9] * 3, np.float32))[1] == 1


def test_skip_links_are_consistent():
    xyz, mf, mc = scenes.terrain(40, 2, 2)
    _, nodes = capi.host_build_tree(mesh_boxes(xyz, mf, mc))
    links = nodes[:, 3].view(np.int32)
    counts = nodes[:, 7].view(np.int32)
    n = len(nodes)
    internal = counts < 0
    assert (links[internal] > np.nonzero(internal)[0] + 2).all() and (links[internal] <= n).all()
    # an internal node's box encloses every node of its subtree [i+1, link)
    for i in np.nonzero(internal)[0][:200]:
        sub = nodes[i + 1:links[i]]
        assert (sub[:, 0:3] >= nodes[i, 0:3]).all() and (sub[:, 4:7] <= nodes[i, 4:7]).all()
    assert counts[~internal].min() == 1
    # internal nodes carry -(second child): the node that follows the first child's subtree
    for i in np.nonzero(internal)[0]:
        first = i + 1
        after_first = links[first] if counts[first] < 0 else first + 1
        assert -counts[i] == after_first


def test_empty_and_single():
    order, nodes = capi.host_build_tree(np.zeros((0, 6), np.float32))
    assert len(order) == 0 and len(nodes) == 0
    order, nodes = capi.host_build_tree(np.array([[0, 0, 0, 1, 1, 1]], np.float32))
    assert list(order) == [0] and len(nodes) == 1 and nodes[0, 7].view(np.int32) == 1
    assert np.allclose(nodes[0, 0:3], -0.05) and np.allclose(nodes[0, 4:7], 1.05)
