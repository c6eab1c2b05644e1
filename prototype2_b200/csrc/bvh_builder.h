// bvh_builder.h -- host-side dynamic AABB tree with the reference's exact insertion / removal / rebalancing
// rules, plus flattening into the device layout.
//
// Why the reference's own build algorithm is part of the product: the narrowphase is an order-dependent
// Gauss-Seidel push-out (reference collision_system.cpp:29-155 mutates the position inside the triangle
// loop), so results depend on the ORDER in which DynamicAABBTree::query emits candidates
// (aabb_tree.cpp:34-68: explicit stack, left pushed before right => right-first preorder).  Reproducing the
// reference's results therefore requires reproducing its tree: branch-and-bound best-sibling insertion with
// its peculiar heap (aabb_tree.cpp:298-339, src/container/priority_queue.h:19-68), AVL-like single rotations
// (aabb_tree.cpp:258-296) and LIFO node reuse (aabb_tree.cpp:222-241).
#ifndef PRT_BVH_BUILDER_H
#define PRT_BVH_BUILDER_H

#include <cstdint>
#include <vector>

#include "prt_math.cuh"

namespace prt {

enum LeafKind : uint8_t { LEAF_NONE = 0, LEAF_MESH = 2, LEAF_CAPSULE = 3 };   // collider_tag.h:6-12 shapes

// Device node, 32 bytes, two float4:
//   lo.xyz, as_float(link)      internal: link = index of the node following this subtree ("skip")
//   hi.xyz, as_float(count)     internal: count = -(index of the second child) < 0 (the first child is index+1);
//                               mesh leaf: link = first triangle, count = #triangles (>= 0);
//                               capsule leaf (unified LITERAL tree only): link = capsule id, count = 0x40000000
// Nodes are stored in right-first preorder, so "descend" is index+1 and a stackless traversal that walks the
// array front to back emits leaves in exactly the reference's candidate order.
struct FlatNode {
    float lo[3]; int32_t link;
    float hi[3]; int32_t count;
};

struct FlatTree {
    std::vector<FlatNode> nodes;
    std::vector<uint32_t> node_leaf_id;     // per node: collider id of the leaf, 0xFFFFFFFF for internal nodes
    std::vector<uint8_t> node_leaf_kind;    // per node: LeafKind
    std::vector<uint32_t> leaf_order;       // collider ids of MESH leaves in traversal order
    std::vector<uint32_t> capsule_order;    // collider ids of CAPSULE leaves in traversal order (unified LITERAL tree)
    uint32_t max_depth = 0;
};

class RefTree {
public:
    explicit RefTree(float margin = 0.05f) : margin_(margin) {}
    void set_margin(float m) { margin_ = m; }

    // DynamicAABBTree::insertLeaf (aabb_tree.cpp:134-180); returns the node index
    int32_t insert_leaf(uint32_t id, LeafKind kind, Box const & aabb);
    // DynamicAABBTree::remove(int32_t) (aabb_tree.cpp:182-220)
    void remove(int32_t index);
    // DynamicAABBTree::update (aabb_tree.cpp:102-111): re-insert leaves whose stored fat box no longer contains aabb
    size_t update(int32_t * tree_indices, Box const * aabbs, size_t n);   // returns how many leaves were re-inserted
    Box const & node_box(int32_t index) const { return nodes_[size_t(index)].box; }
    int32_t size() const { return size_; }
    bool empty() const { return root_ == NIL; }

    // right-first preorder flattening.  tri_first/tri_count: per MESH collider id, its triangle range in the
    // (leaf-ordered) device triangle array -- filled by the caller through `assign`.
    template <class LeafRange>
    void flatten(FlatTree & out, LeafRange && leaf_range) const;

private:
    static constexpr int32_t NIL = -1;
    struct Node {                           // aabb_tree.h:112-139
        Box box;
        int32_t height;
        int32_t parent;                     // doubles as `next` of the free list
        int32_t left, right;
        uint32_t id;
        LeafKind kind;
    };
    struct Cost { int32_t index; float direct, inherited; };     // aabb_tree.h:141-147

    int32_t allocate();
    void release(int32_t index);
    int32_t best_sibling(int32_t leaf);
    void sync_hierarchy(int32_t index);
    void balance(int32_t index);
    static Node blank();
    static bool less(Cost const & l, Cost const & r) { return l.direct + l.inherited < r.direct + r.inherited; }
    void heap_push(Cost const & t);
    void heap_pop();

    float margin_;
    int32_t root_ = NIL, free_head_ = NIL, size_ = 0;
    std::vector<Node> nodes_;
    std::vector<Cost> heap_;                // scratch of best_sibling
};

template <class LeafRange>
void RefTree::flatten(FlatTree & out, LeafRange && leaf_range) const {
    out.nodes.clear(); out.node_leaf_id.clear(); out.node_leaf_kind.clear(); out.leaf_order.clear(); out.capsule_order.clear();
    out.max_depth = 0;
    if (root_ == NIL) return;
    out.nodes.reserve(size_t(size_));
    struct Item { int32_t node; int32_t flat_parent; uint32_t depth; };
    std::vector<Item> stack;
    stack.push_back({root_, -1, 0});
    // An internal node's skip link = flat index of the first node emitted after its whole subtree.  With an
    // explicit stack holding S items when the node was popped, its subtree is finished the first time a pop
    // leaves fewer than S items behind.
    std::vector<std::pair<int32_t, uint32_t>> pending;   // (flat index, S)
    while (!stack.empty()) {
        Item it = stack.back();
        stack.pop_back();
        while (!pending.empty() && pending.back().second > stack.size()) {
            out.nodes[size_t(pending.back().first)].link = int32_t(out.nodes.size());
            pending.pop_back();
        }
        Node const & n = nodes_[size_t(it.node)];
        FlatNode f;
        f.lo[0] = n.box.lo.x; f.lo[1] = n.box.lo.y; f.lo[2] = n.box.lo.z;
        f.hi[0] = n.box.hi.x; f.hi[1] = n.box.hi.y; f.hi[2] = n.box.hi.z;
        int32_t flat = int32_t(out.nodes.size());
        if (it.depth > out.max_depth) out.max_depth = it.depth;
        if (it.flat_parent >= 0) out.nodes[size_t(it.flat_parent)].count = -flat;   // we are our parent's second child
        if (n.right == NIL) {               // Node::isLeaf (aabb_tree.h:120)
            uint32_t first = 0, count = 0;
            if (n.kind == LEAF_MESH) { leaf_range(n.id, first, count); out.leaf_order.push_back(n.id); }
            else if (n.kind == LEAF_CAPSULE) { first = n.id; count = 0x40000000u; out.capsule_order.push_back(n.id); }   // PRT_CAPSULE_LEAF
            f.link = int32_t(first);
            f.count = int32_t(count);
            out.nodes.push_back(f);
            out.node_leaf_id.push_back(n.id);
            out.node_leaf_kind.push_back(uint8_t(n.kind));
        } else {
            f.link = -1;
            f.count = -1;
            out.nodes.push_back(f);
            out.node_leaf_id.push_back(0xFFFFFFFFu);
            out.node_leaf_kind.push_back(uint8_t(LEAF_NONE));
            // the subtree of `flat` is finished when the stack shrinks back to its current size
            pending.push_back({flat, uint32_t(stack.size())});
            stack.push_back({n.left, flat, it.depth + 1});      // aabb_tree.cpp:63-64: left pushed first,
            stack.push_back({n.right, -1, it.depth + 1});       // right popped first (=> it is flat+1)
        }
    }
    while (!pending.empty()) {
        out.nodes[size_t(pending.back().first)].link = int32_t(out.nodes.size());
        pending.pop_back();
    }
}

// Top of the flattened tree as a compact table for staging in shared memory (k_xc_refresh): every node of depth <= D,
// D the largest depth whose levels together fit `cap` entries, in the same right-first preorder -- so a walk that runs
// over the table and drops into the full array below it visits nodes in the order of the plain walk and emits the same
// leaf sequence.  Entry = FlatNode with the two integer words re-used:
//   internal, depth <  D:  link = TABLE index to continue at when the box is missed (table skip), count = -1
//   internal, depth == D:  link = its index g in the full array, count = -(skip of g) <= -3: on overlap the walk goes on
//                          in the full array over [g + 1, skip)
//   leaf:                  link = its index g in the full array, count = 0: on overlap the walk visits [g, g + 1)
// Returns an empty table when the whole tree is not bigger than a few tables (nothing to gain).
inline void build_top_table(std::vector<FlatNode> const & nodes, uint32_t cap, std::vector<FlatNode> & top) {
    top.clear();
    const size_t n = nodes.size();
    if (n < size_t(cap) * 4) return;
    std::vector<uint8_t> depth(n, 0);
    std::vector<uint32_t> per_depth;
    for (size_t i = 0; i < n; ++i) {
        const uint32_t d = depth[i];
        if (d >= per_depth.size()) per_depth.resize(d + 1, 0);
        ++per_depth[d];
        if (nodes[i].count < 0) {
            const uint8_t dc = uint8_t(d < 255 ? d + 1 : 255);
            depth[i + 1] = dc;
            depth[size_t(-nodes[i].count)] = dc;
        }
    }
    uint32_t D = 0, kept = per_depth[0];
    while (D + 1 < per_depth.size() && D + 1 < 255 && kept + per_depth[D + 1] <= cap) { ++D; kept += per_depth[D]; }
    if (D == 0) return;
    std::vector<int32_t> table_index(n + 1, -1);
    int32_t k = 0;
    for (size_t i = 0; i < n; ++i) if (depth[i] <= D) table_index[i] = k++;
    table_index[n] = k;
    top.reserve(size_t(k));
    for (size_t i = 0; i < n; ++i) {
        if (depth[i] > D) continue;
        FlatNode f = nodes[i];
        if (nodes[i].count < 0) {
            if (depth[i] < D) { f.link = table_index[size_t(nodes[i].link)]; f.count = -1; }   // the node after a subtree is never deeper than its root
            else { f.link = int32_t(i); f.count = -nodes[i].link; }
        } else {
            f.link = int32_t(i); f.count = 0;
        }
        top.push_back(f);
    }
}

} // namespace prt
#endif
