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

#include <atomic>
#include <cstdint>
#include <thread>
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
    // the same arrays, written by several host threads (mesh-only trees; `leaf_range` must be safe to call concurrently):
    // the top levels are laid out by one thread, the subtrees below them are counted and then written in parallel
    template <class LeafRange>
    void flatten_parallel(FlatTree & out, LeafRange && leaf_range, unsigned threads) const;

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

template <class LeafRange>
void RefTree::flatten_parallel(FlatTree & out, LeafRange && leaf_range, unsigned threads) const {
    constexpr uint32_t CUT = 9;                                   // subtrees rooted at this depth (or leaves above it) are the parallel tasks
    if (root_ == NIL || size_ < 8192 || threads < 2) { flatten(out, leaf_range); return; }
    struct Ref { bool task; uint32_t idx; };
    struct Top { int32_t node; Ref r, l; uint32_t size, leaves; size_t flat, leaf_base; };
    struct Task { int32_t root; uint32_t size, leaves, depth; size_t flat, leaf_base; };
    std::vector<Top> tops;
    std::vector<Task> tasks;
    // 1. the top part, right-first preorder (aabb_tree.cpp:63-64: left pushed first, right popped first)
    struct Frame { int32_t node; uint32_t depth; int parent; bool second; };
    std::vector<Frame> stack;
    stack.push_back({root_, 0, -1, false});
    Ref root_ref{false, 0};
    while (!stack.empty()) {
        const Frame f = stack.back();
        stack.pop_back();
        Node const & n = nodes_[size_t(f.node)];
        Ref me;
        if (n.right == NIL || f.depth >= CUT) { me = Ref{true, uint32_t(tasks.size())}; tasks.push_back(Task{f.node, 0, 0, 0, 0, 0}); }
        else {
            me = Ref{false, uint32_t(tops.size())};
            tops.push_back(Top{f.node, Ref{true, 0}, Ref{true, 0}, 0, 0, 0, 0});
            stack.push_back({n.left, f.depth + 1, int(me.idx), true});
            stack.push_back({n.right, f.depth + 1, int(me.idx), false});
        }
        if (f.parent < 0) root_ref = me;
        else if (f.second) tops[size_t(f.parent)].l = me;
        else tops[size_t(f.parent)].r = me;
    }
    auto run = [&](auto && body) {
        std::atomic<size_t> next{0};
        auto worker = [&]() { for (;;) { const size_t k = next.fetch_add(1); if (k >= tasks.size()) break; body(tasks[k]); } };
        const unsigned nt = unsigned(std::min<size_t>(threads, tasks.size()));
        std::vector<std::thread> pool;
        for (unsigned t = 1; t < nt; ++t) pool.emplace_back(worker);
        worker();
        for (auto & th : pool) th.join();
    };
    // 2. sizes of the subtrees, in parallel
    run([&](Task & t) {
        std::vector<std::pair<int32_t, uint32_t>> st;
        st.push_back({t.root, 0u});
        uint32_t size = 0, leaves = 0, depth = 0;
        while (!st.empty()) {
            const auto it = st.back();
            st.pop_back();
            Node const & n = nodes_[size_t(it.first)];
            ++size;
            if (it.second > depth) depth = it.second;
            if (n.right == NIL) { if (n.kind == LEAF_MESH) ++leaves; }
            else { st.push_back({n.left, it.second + 1}); st.push_back({n.right, it.second + 1}); }
        }
        t.size = size; t.leaves = leaves; t.depth = depth;
    });
    // 3. sizes of the top nodes (children follow their parents in the list), then every node's place
    auto size_of = [&](Ref r) { return r.task ? tasks[r.idx].size : tops[r.idx].size; };
    auto leaves_of = [&](Ref r) { return r.task ? tasks[r.idx].leaves : tops[r.idx].leaves; };
    for (size_t k = tops.size(); k-- > 0;) {
        tops[k].size = 1u + size_of(tops[k].r) + size_of(tops[k].l);
        tops[k].leaves = leaves_of(tops[k].r) + leaves_of(tops[k].l);
    }
    const size_t nn = size_of(root_ref), nl = leaves_of(root_ref);
    out.nodes.resize(nn); out.node_leaf_id.resize(nn); out.node_leaf_kind.resize(nn);
    out.leaf_order.resize(nl); out.capsule_order.clear();
    struct Place { Ref r; size_t flat, leaf_base; };
    std::vector<Place> todo;
    todo.push_back({root_ref, 0, 0});
    uint32_t max_depth = 0;
    while (!todo.empty()) {
        const Place pl = todo.back();
        todo.pop_back();
        if (pl.r.task) { tasks[pl.r.idx].flat = pl.flat; tasks[pl.r.idx].leaf_base = pl.leaf_base; continue; }
        Top & t = tops[pl.r.idx];
        t.flat = pl.flat; t.leaf_base = pl.leaf_base;
        Node const & n = nodes_[size_t(t.node)];
        FlatNode f;
        f.lo[0] = n.box.lo.x; f.lo[1] = n.box.lo.y; f.lo[2] = n.box.lo.z;
        f.hi[0] = n.box.hi.x; f.hi[1] = n.box.hi.y; f.hi[2] = n.box.hi.z;
        f.link = int32_t(pl.flat + t.size);                        // skip: the node after this subtree
        f.count = -int32_t(pl.flat + 1 + size_of(t.r));            // second child (the left one)
        out.nodes[pl.flat] = f;
        out.node_leaf_id[pl.flat] = 0xFFFFFFFFu;
        out.node_leaf_kind[pl.flat] = uint8_t(LEAF_NONE);
        todo.push_back({t.r, pl.flat + 1, pl.leaf_base});
        todo.push_back({t.l, pl.flat + 1 + size_of(t.r), pl.leaf_base + leaves_of(t.r)});
    }
    for (Task const & t : tasks) if (t.depth + CUT > max_depth) max_depth = t.depth + CUT;
    out.max_depth = max_depth;
    // 4. the subtrees, in parallel: the serial algorithm with the subtree's base added
    run([&](Task & t) {
        struct Item { int32_t node; int64_t flat_parent; };
        std::vector<Item> st;
        std::vector<std::pair<size_t, uint32_t>> pending;
        st.push_back({t.root, -1});
        size_t at = t.flat, leaf_at = t.leaf_base;
        while (!st.empty()) {
            const Item it = st.back();
            st.pop_back();
            while (!pending.empty() && pending.back().second > st.size()) { out.nodes[pending.back().first].link = int32_t(at); pending.pop_back(); }
            Node const & n = nodes_[size_t(it.node)];
            FlatNode f;
            f.lo[0] = n.box.lo.x; f.lo[1] = n.box.lo.y; f.lo[2] = n.box.lo.z;
            f.hi[0] = n.box.hi.x; f.hi[1] = n.box.hi.y; f.hi[2] = n.box.hi.z;
            if (it.flat_parent >= 0) out.nodes[size_t(it.flat_parent)].count = -int32_t(at);
            if (n.right == NIL) {
                uint32_t first = 0, count = 0;
                if (n.kind == LEAF_MESH) { leaf_range(n.id, first, count); out.leaf_order[leaf_at++] = n.id; }
                f.link = int32_t(first); f.count = int32_t(count);
                out.node_leaf_id[at] = n.id;
                out.node_leaf_kind[at] = uint8_t(n.kind);
            } else {
                f.link = -1; f.count = -1;
                out.node_leaf_id[at] = 0xFFFFFFFFu;
                out.node_leaf_kind[at] = uint8_t(LEAF_NONE);
                pending.push_back({at, uint32_t(st.size())});
                st.push_back({n.left, int64_t(at)});
                st.push_back({n.right, -1});
            }
            out.nodes[at] = f;
            ++at;
        }
        while (!pending.empty()) { out.nodes[pending.back().first].link = int32_t(at); pending.pop_back(); }
    });
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
inline void build_top_table(std::vector<FlatNode> const & nodes, uint32_t cap, std::vector<FlatNode> & top, std::vector<uint32_t> * src = nullptr) {
    top.clear();
    if (src) src->clear();
    const size_t n = nodes.size();
    if (n < size_t(cap) * 4) return;
    // level by level from the root until the next level no longer fits: only the table's own nodes are touched
    std::vector<uint32_t> level{0u}, next;
    uint32_t D = 0, kept = 1;
    for (;;) {
        next.clear();
        for (uint32_t i : level)
            if (nodes[i].count < 0) { next.push_back(i + 1); next.push_back(uint32_t(-nodes[i].count)); }
        if (next.empty() || kept + next.size() > cap || D + 1 >= 255) break;
        ++D; kept += uint32_t(next.size());
        level.swap(next);
    }
    if (D == 0) return;
    // the kept nodes in preorder (= ascending index) with their depths
    struct It { uint32_t node, depth; };
    std::vector<It> order, st;
    st.push_back({0u, 0u});
    while (!st.empty()) {
        const It it = st.back();
        st.pop_back();
        order.push_back(it);
        if (nodes[it.node].count < 0 && it.depth < D) {
            st.push_back({uint32_t(-nodes[it.node].count), it.depth + 1});      // second child later
            st.push_back({it.node + 1, it.depth + 1});
        }
    }
    // table index of the first kept node at or after a full-array index (for the table skip links)
    auto table_index_of = [&](uint32_t full) {
        size_t lo = 0, hi = order.size();
        while (lo < hi) { const size_t mid = (lo + hi) / 2; if (order[mid].node < full) lo = mid + 1; else hi = mid; }
        return int32_t(lo);
    };
    top.reserve(order.size());
    for (It const & it : order) {
        FlatNode f = nodes[it.node];
        if (nodes[it.node].count < 0) {
            if (it.depth < D) { f.link = table_index_of(uint32_t(nodes[it.node].link)); f.count = -1; }   // the node after a subtree is never deeper than its root
            else { f.link = int32_t(it.node); f.count = -nodes[it.node].link; }
        } else {
            f.link = int32_t(it.node); f.count = 0;
        }
        top.push_back(f);
        if (src) src->push_back(it.node);
    }
}

} // namespace prt
#endif
