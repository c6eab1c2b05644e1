// bvh_builder.cpp -- see bvh_builder.h.  Host code; all float arithmetic is plain IEEE binary32
// (+, -, *, compare), compiled with -ffp-contract=off.
#include "bvh_builder.h"

#include <algorithm>

namespace prt {

RefTree::Node RefTree::blank() {           // `m_nodes[index] = {}` / `push_back({})`, aabb_tree.cpp:233,237
    Node n;
    n.box.lo = v3(0.0f); n.box.hi = v3(0.0f);
    n.height = 0; n.parent = NIL; n.left = NIL; n.right = NIL; n.id = 0; n.kind = LEAF_NONE;
    return n;
}

int32_t RefTree::allocate() {              // aabb_tree.cpp:229-241
    int32_t index;
    if (free_head_ == NIL) {
        index = int32_t(nodes_.size());
        nodes_.push_back(blank());
    } else {
        index = free_head_;
        free_head_ = nodes_[size_t(index)].parent;
        nodes_[size_t(index)] = blank();
    }
    ++size_;
    return index;
}

void RefTree::release(int32_t index) {     // aabb_tree.cpp:222-227
    nodes_[size_t(index)].parent = free_head_;
    nodes_[size_t(index)].height = -1;
    free_head_ = index;
    --size_;
}

// prt::priority_queue::push (priority_queue.h:19-32): sift-up that keeps climbing after a failed compare and
// always compares the NEW element against the ancestor, swapping whatever sits below it.
void RefTree::heap_push(Cost const & t) {
    size_t index = heap_.size();
    heap_.push_back(t);
    while (index != 0) {
        size_t parent = (index / 2 + (index % 2 != 0)) - 1;
        Cost elem = heap_[index];
        if (less(t, heap_[parent])) {
            heap_[index] = heap_[parent];
            heap_[parent] = elem;
        }
        index = parent;
    }
}

// prt::priority_queue::pop (priority_queue.h:34-68): sift-down that always runs to a leaf, taking the right
// child unless the left one is strictly smaller.
void RefTree::heap_pop() {
    heap_.front() = heap_.back();
    heap_.pop_back();
    size_t index = 0;
    for (;;) {
        size_t l = index * 2 + 1, r = l + 1, child;
        bool has_l = l < heap_.size(), has_r = r < heap_.size();
        if (has_l && has_r) child = less(heap_[l], heap_[r]) ? l : r;
        else if (has_l) child = l;
        else if (has_r) child = r;
        else return;
        Cost c = heap_[child];
        if (less(c, heap_[index])) {
            heap_[child] = heap_[index];
            heap_[index] = c;
        }
        index = child;
    }
}

int32_t RefTree::best_sibling(int32_t leaf_index) {        // aabb_tree.cpp:298-339
    Box const leaf = nodes_[size_t(leaf_index)].box;
    float const leaf_area = box_area(leaf);
    int32_t best = root_;
    float best_cost = box_area(box_union(leaf, nodes_[size_t(root_)].box));
    heap_.clear();
    heap_push(Cost{root_, best_cost, 0.0f});
    while (!heap_.empty()) {
        Cost nc = heap_.front();
        heap_pop();
        float current = nc.direct + nc.inherited;
        if (current < best_cost) {
            best_cost = current;
            best = nc.index;
        }
        Node const & curr = nodes_[size_t(nc.index)];
        if (curr.right != NIL) {
            float inherited = nc.inherited + box_area(box_union(leaf, curr.box)) - box_area(curr.box);
            float c_low = leaf_area + inherited;
            if (c_low < best_cost) {
                int32_t l = curr.left, r = curr.right;
                Cost cl{l, box_area(box_union(leaf, nodes_[size_t(l)].box)), inherited};
                Cost cr{r, box_area(box_union(leaf, nodes_[size_t(r)].box)), inherited};
                heap_push(cl);
                heap_push(cr);
            }
        }
    }
    return best;
}

void RefTree::balance(int32_t index) {                      // aabb_tree.cpp:258-296
    if (nodes_[size_t(index)].height < 2) return;
    int32_t il = nodes_[size_t(index)].left, ir = nodes_[size_t(index)].right;
    int32_t diff = nodes_[size_t(il)].height - nodes_[size_t(ir)].height;
    auto h = [&](int32_t i) { return nodes_[size_t(i)].height; };
    if (diff > 1) {            // left is higher: its taller child moves up to be `index`'s right child
        Node & left = nodes_[size_t(il)];
        bool take_left = h(left.left) > h(left.right);
        int32_t a = take_left ? left.left : left.right;
        nodes_[size_t(ir)].parent = il;
        nodes_[size_t(a)].parent = index;
        nodes_[size_t(index)].right = a;
        (take_left ? left.left : left.right) = ir;
        left.box = box_union(nodes_[size_t(left.left)].box, nodes_[size_t(left.right)].box);
        left.height = std::max(h(left.left), h(left.right)) + 1;
    } else if (diff < -1) {    // right is higher
        Node & right = nodes_[size_t(ir)];
        bool take_left = h(right.left) > h(right.right);
        int32_t a = take_left ? right.left : right.right;
        nodes_[size_t(il)].parent = ir;
        nodes_[size_t(a)].parent = index;
        nodes_[size_t(index)].left = a;
        (take_left ? right.left : right.right) = il;
        right.box = box_union(nodes_[size_t(right.left)].box, nodes_[size_t(right.right)].box);
        right.height = std::max(h(right.left), h(right.right)) + 1;
    }
    nodes_[size_t(index)].height = std::max(h(nodes_[size_t(index)].left), h(nodes_[size_t(index)].right)) + 1;
}

void RefTree::sync_hierarchy(int32_t index) {               // aabb_tree.cpp:243-256
    while (index != NIL) {
        balance(index);
        Node & n = nodes_[size_t(index)];
        n.height = 1 + std::max(nodes_[size_t(n.left)].height, nodes_[size_t(n.right)].height);
        n.box = box_union(nodes_[size_t(n.left)].box, nodes_[size_t(n.right)].box);
        index = n.parent;
    }
}

int32_t RefTree::insert_leaf(uint32_t id, LeafKind kind, Box const & aabb) {   // aabb_tree.cpp:134-180
    int32_t leaf = allocate();
    nodes_[size_t(leaf)].id = id;
    nodes_[size_t(leaf)].kind = kind;
    nodes_[size_t(leaf)].box.lo = aabb.lo - margin_;
    nodes_[size_t(leaf)].box.hi = aabb.hi + margin_;
    if (root_ == NIL) {
        root_ = leaf;
        return leaf;
    }
    int32_t sibling = best_sibling(leaf);
    int32_t old_parent = nodes_[size_t(sibling)].parent;
    int32_t new_parent = allocate();
    nodes_[size_t(new_parent)].parent = old_parent;
    nodes_[size_t(new_parent)].box = box_union(nodes_[size_t(leaf)].box, nodes_[size_t(sibling)].box);
    nodes_[size_t(new_parent)].height = 1 + nodes_[size_t(sibling)].height;
    if (old_parent != NIL) {
        if (nodes_[size_t(old_parent)].left == sibling) nodes_[size_t(old_parent)].left = new_parent;
        else nodes_[size_t(old_parent)].right = new_parent;
    } else {
        root_ = new_parent;
    }
    nodes_[size_t(new_parent)].left = sibling;
    nodes_[size_t(new_parent)].right = leaf;
    nodes_[size_t(sibling)].parent = new_parent;
    nodes_[size_t(leaf)].parent = new_parent;
    sync_hierarchy(new_parent);
    return leaf;
}

void RefTree::remove(int32_t index) {                       // aabb_tree.cpp:182-220
    if (index == root_) {
        root_ = NIL;
        release(index);
        return;
    }
    int32_t parent = nodes_[size_t(index)].parent;
    int32_t sibling = nodes_[size_t(parent)].left == index ? nodes_[size_t(parent)].right : nodes_[size_t(parent)].left;
    int32_t grand = nodes_[size_t(parent)].parent;
    if (grand != NIL) {
        if (nodes_[size_t(grand)].left == parent) nodes_[size_t(grand)].left = sibling;
        else nodes_[size_t(grand)].right = sibling;
        nodes_[size_t(sibling)].parent = grand;
        release(parent);
        sync_hierarchy(grand);
    } else {
        root_ = sibling;
        nodes_[size_t(sibling)].parent = NIL;
        release(parent);
    }
    release(index);
}

size_t RefTree::update(int32_t * tree_indices, Box const * aabbs, size_t n) {    // aabb_tree.cpp:102-111
    size_t reinserted = 0;
    for (size_t i = 0; i < n; ++i) {
        Node const & node = nodes_[size_t(tree_indices[i])];
        if (!box_contains(node.box, aabbs[i])) {
            uint32_t id = node.id;
            LeafKind kind = node.kind;
            remove(tree_indices[i]);
            tree_indices[i] = insert_leaf(id, kind, aabbs[i]);
            ++reinserted;
        }
    }
    return reinserted;
}

} // namespace prt
