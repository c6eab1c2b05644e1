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


def walk_staged(top, nodes, box, max_doors=6):
    """numpy restatement of k_xc_refresh's walk with the top-of-tree table (k_step_stream.cu): phase 1 runs over the
    table alone and collects the overlapping entries at the cut ("doors", at most 6, else the plain walk is used),
    phase 2 walks the ranges of the full array behind the doors in order; returns the emitted NODE indices, the two
    test counts and whether the walk fell back"""
    tl, tc = top[:, 3].view(np.int32), top[:, 7].view(np.int32)
    nl, nc = nodes[:, 3].view(np.int32), nodes[:, 7].view(np.int32)

    def ov(rec):
        lo, hi = rec[0:3], rec[4:7]
        return not (lo[0] > box[3] or hi[0] < box[0] or lo[1] > box[4] or hi[1] < box[1] or lo[2] > box[5] or hi[2] < box[2])
    doors, k, t_top, t_full, fell_back = [], 0, 0, 0, False
    while k < len(top):
        t_top += 1
        o = ov(top[k])
        if tc[k] == -1:
            k = k + 1 if o else int(tl[k])
        else:
            if o:
                if len(doors) == max_doors:
                    fell_back = True
                    break
                doors.append(k)
            k += 1
    ranges = [(0, len(nodes))] if fell_back else [
        (int(tl[k]), int(tl[k]) + 1) if tc[k] == 0 else (int(tl[k]) + 1, -int(tc[k])) for k in doors]
    out = []
    for node, node_end in ranges:
        while node != node_end:
            t_full += 1
            o = ov(nodes[node])
            if nc[node] < 0:
                node = node + 1 if o else int(nl[node])
            else:
                if o:
                    out.append(node)
                node += 1
            assert node <= node_end
    return out, t_top, t_full, fell_back


def walk_nodes(nodes, box):
    """plain walk returning node indices of the overlapping leaves"""
    nl, nc = nodes[:, 3].view(np.int32), nodes[:, 7].view(np.int32)
    out, i, tested = [], 0, 0
    while i < len(nodes):
        lo, hi = nodes[i, 0:3], nodes[i, 4:7]
        tested += 1
        o = not (lo[0] > box[3] or hi[0] < box[0] or lo[1] > box[4] or hi[1] < box[1] or lo[2] > box[5] or hi[2] < box[2])
        if nc[i] < 0:
            i = i + 1 if o else int(nl[i])
        else:
            if o:
                out.append(i)
            i += 1
    return out, tested


@pytest.mark.parametrize("cap", [16, 64, 1024])
def test_top_table_walk_equals_plain_walk(cap):
    """PRTC_OPT_TOP_STAGING: the walk over the staged top levels + the full array below the cut emits the same leaf
    sequence as the plain walk, and tests each node of the plain walk once (leaves above the cut twice)"""
    rng = np.random.default_rng(11)
    xyz, mf, mc = scenes.terrain(96, 2, 2)
    boxes = mesh_boxes(xyz, mf, mc)
    extra = rng.uniform(0, 96, (3000, 3)).astype(np.float32)                 # an unbalanced part: clustered random boxes
    extra[:, 1] = rng.uniform(-2, 6, 3000)
    boxes = np.concatenate([boxes, np.concatenate([extra, extra + rng.uniform(0.1, 4, (3000, 3)).astype(np.float32)], axis=1)])
    _, nodes = capi.host_build_tree(boxes)
    top = capi.host_top_table(nodes, cap)
    if len(nodes) < 4 * cap:
        assert len(top) == 0
        return
    assert 0 < len(top) <= cap
    tc = top[:, 7].view(np.int32)
    assert (tc == -1).any() and (tc < -1).any()
    # entries keep the boxes of the nodes they stand for
    g = top[:, 3].view(np.int32)
    for k in np.nonzero(tc != -1)[0][:200]:
        assert np.array_equal(top[k, [0, 1, 2, 4, 5, 6]], nodes[g[k], [0, 1, 2, 4, 5, 6]])
    n_fallback = 0
    for q in range(300):
        c = rng.uniform(0, 96, 3)
        c[1] = rng.uniform(-2, 3)
        e = rng.uniform(0.2, 6.0, 3) if q % 10 else np.array([200.0, 200.0, 200.0])   # every 10th box covers everything
        box = np.concatenate([c - e, c + e]).astype(np.float32)
        plain, tested = walk_nodes(nodes, box)
        staged, t_top, t_full, fell_back = walk_staged(top, nodes, box)
        assert staged == plain
        n_fallback += fell_back
        if fell_back:
            continue
        leaf_hits_above_cut = sum(1 for k in range(len(top)) if tc[k] == 0 and not (
            top[k, 0] > box[3] or top[k, 4] < box[0] or top[k, 1] > box[4] or top[k, 5] < box[1] or top[k, 2] > box[5] or top[k, 6] < box[2]))
        assert t_top + t_full == tested + leaf_hits_above_cut
    assert 30 <= n_fallback < 300          # the all-covering boxes open more than 6 doors; how many of the others do depends on how deep the cut is


def test_top_table_small_tree_is_not_staged():
    _, nodes = capi.host_build_tree(np.array([[0, 0, 0, 1, 1, 1], [2, 0, 0, 3, 1, 1], [0, 2, 0, 1, 3, 1]], np.float32))
    assert len(capi.host_top_table(nodes, 1024)) == 0
    assert len(capi.host_top_table(np.zeros((0, 8), np.float32), 1024)) == 0


def _reference_balanced(boxes):
    """the left-complete tree over the boxes by plain recursion (first half first): what the closed-form index arithmetic
    of csrc/balanced_tree.cuh must reproduce"""
    n = len(boxes)
    nodes = np.zeros((2 * n - 1, 8), np.float32)
    ints = nodes.view(np.int32)
    pos = [0]

    def build(a, b, size):                        # block [a, a + size) clipped to n, size a power of two
        while size > 1 and a + size // 2 >= b:    # second half empty: collapses into the first half
            size //= 2
        idx = pos[0]
        pos[0] += 1
        if size == 1:
            nodes[idx, 0:3], nodes[idx, 4:7] = boxes[a, 0:3], boxes[a, 3:6]
            ints[idx, 3], ints[idx, 7] = a, 1
            return idx
        c1 = build(a, a + size // 2, size // 2)
        c2 = build(a + size // 2, b, size // 2)
        nodes[idx, 0:3] = np.minimum(nodes[c1, 0:3], nodes[c2, 0:3])
        nodes[idx, 4:7] = np.maximum(nodes[c1, 4:7], nodes[c2, 4:7])
        ints[idx, 3], ints[idx, 7] = pos[0], -c2
        return idx
    size = 1
    while size < n:
        size *= 2
    build(0, n, size)
    assert pos[0] == 2 * n - 1
    return nodes


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 7, 8, 9, 31, 33, 100, 1000, 1025])
def test_balanced_tree_layout_equals_recursive_construction(n):
    """PRTC_ORDER_BY_ID: closed-form preorder indices, skip links, second-child indices and boxes == a recursive build"""
    rng = np.random.default_rng(n)
    lo = rng.uniform(-50, 50, (n, 3)).astype(np.float32)
    boxes = np.concatenate([lo, lo + rng.uniform(0.1, 5, (n, 3)).astype(np.float32)], axis=1)
    got = capi.host_balanced_tree(boxes)
    want = _reference_balanced(boxes)
    assert got.shape == want.shape
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_balanced_tree_walk_emits_overlapping_boxes_in_ascending_id():
    rng = np.random.default_rng(5)
    n = 3001
    lo = rng.uniform(0, 100, (n, 3)).astype(np.float32)
    boxes = np.concatenate([lo, lo + rng.uniform(0.1, 3, (n, 3)).astype(np.float32)], axis=1)
    boxes[17, :] = np.nan                                   # an all-NaN box is a candidate of every query (aabb.cpp:3-10)
    boxes[1200, 1] = np.nan                                 # one NaN coordinate: that compare never rejects
    nodes = capi.host_balanced_tree(boxes)
    for q in range(200):
        c = rng.uniform(0, 100, 3)
        e = rng.uniform(0.5, 8, 3)
        box = np.concatenate([c - e, c + e]).astype(np.float32)
        got, _ = walk_flat(nodes, box)
        brute = [i for i in range(n) if not (boxes[i, 0] > box[3] or boxes[i, 3] < box[0] or boxes[i, 1] > box[4] or
                                             boxes[i, 4] < box[1] or boxes[i, 2] > box[5] or boxes[i, 5] < box[2])]
        assert list(got) == brute
        assert 17 in brute


def test_balanced_tree_empty():
    assert len(capi.host_balanced_tree(np.zeros((0, 6), np.float32))) == 0


def test_parallel_flatten_equals_serial_flatten(monkeypatch):
    """a commit of a big level lays the reference-built tree out with several host threads (top levels by one thread, the
    subtrees below counted and written in parallel): node for node and leaf for leaf the serial right-first preorder"""
    rs = np.random.RandomState(42)
    for n in (9000, 60000):
        lo = rs.uniform(0, 300, (n, 3)).astype(np.float32)
        boxes = np.concatenate([lo, lo + rs.uniform(0.1, 3.0, (n, 3)).astype(np.float32)], axis=1)
        monkeypatch.delenv("PRTC_HOST_FLATTEN_THREADS", raising=False)
        order_a, nodes_a = capi.host_build_tree(boxes)
        for threads in ("2", "7", "16"):
            monkeypatch.setenv("PRTC_HOST_FLATTEN_THREADS", threads)
            order_b, nodes_b = capi.host_build_tree(boxes)
            assert np.array_equal(order_a, order_b)
            assert np.array_equal(nodes_a.view(np.uint32), nodes_b.view(np.uint32))
