// balanced_tree.cuh -- index arithmetic of the tree PRTC_ORDER_BY_ID builds on the device (k_tree_build.cu), shared with
// the host-only utility prtc_host_balanced_tree (prtc.cu) so that the layout can be checked without a GPU.
//
// The tree is the left-complete binary tree over the L live mesh leaves taken in ascending collider id: with
// 2^k >= L > 2^(k-1), block (level, j) covers the leaves [j * 2^level, min((j + 1) * 2^level, L)); a block whose second
// half is empty is not a node of its own (it collapses into its first half), every other block with level >= 1 is an
// internal node with the two half blocks as children, blocks of level 0 are the leaves.  Stored in preorder with the
// FIRST half first, the flat array has the layout every kernel already walks (bvh_builder.h FlatNode: "descend" is
// index + 1, `link` skips the subtree) and a front-to-back walk emits the leaves in ascending id -- the candidate order
// of this mode.  Everything about a node follows from (level, j) alone:
//   preorder index  = 2 * (first leaf) + (number of real ancestors the path from the root leaves through the FIRST half)
//                     -- left of the path hang complete subtrees over the leaves before the block (2 * leaves - 1 nodes
//                     each, one per turn into a second half), above it sit the ancestors;
//   subtree size    = 2 * (leaves of the block) - 1;   second child = index + 2^level  (the first half is full).
#ifndef PRT_BALANCED_TREE_CUH
#define PRT_BALANCED_TREE_CUH
#include <cstdint>
#include "prt_math.cuh"

namespace prt {

PRT_HD uint32_t bt_levels(uint32_t n_leaves) {
    uint32_t k = 0;
    while ((1ull << k) < n_leaves) ++k;
    return k;
}
PRT_HD bool bt_is_node(uint32_t level, uint32_t j, uint32_t n_leaves) {
    const unsigned long long first = (unsigned long long)j << level;
    if (first >= n_leaves) return false;
    return level == 0 || first + (1ull << (level - 1)) < n_leaves;
}
PRT_HD uint32_t bt_leaves_of(uint32_t level, uint32_t j, uint32_t n_leaves) {
    const unsigned long long first = (unsigned long long)j << level, end = first + (1ull << level);
    return uint32_t((end < n_leaves ? end : n_leaves) - first);
}
PRT_HD uint32_t bt_index(uint32_t level, uint32_t j, uint32_t k, uint32_t n_leaves) {
    uint32_t first_half_turns = 0;
    for (uint32_t up = level + 1; up <= k; ++up) {
        if (((j >> (up - 1 - level)) & 1u) == 0u) {                 // the path enters the FIRST half of the ancestor at level `up`
            const unsigned long long anc_first = (unsigned long long)(j >> (up - level)) << up;
            if (anc_first + (1ull << (up - 1)) < n_leaves) ++first_half_turns;     // ... which is a real node (its second half exists)
        }
    }
    return uint32_t(2ull * ((unsigned long long)j << level)) + first_half_turns;
}

} // namespace prt
#endif
