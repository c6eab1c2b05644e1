// k_tree_build.cu -- PRTC_ORDER_BY_ID: the static tree is built on the device.
//
// The reference builds its tree by sequential insertion (aabb_tree.cpp:134-180, 298-328): 2.2 s on the host for C4's
// 1.25 M leaves, and inherently serial because the shape -- hence the candidate ORDER of every query -- depends on the
// insertion history.  The candidate SET does not (it is "every leaf whose fat box overlaps the query box", SURVEY
// finding 3), so a mode that takes the set in ascending collider id is free to use any tree whose leaves come out in
// that order.  This one is the left-complete tree over the id sequence (balanced_tree.cuh): no sorting, no
// Morton codes, node indices in closed form, boxes by one bottom-up pass per level.
#include "balanced_tree.cuh"
#include "kernels.cuh"

namespace prt {

namespace {

// level 0: the host hands the live leaves over as finished leaf records (fat box, first triangle, triangle count) in
// ascending id; each goes to its preorder position
__global__ void k_bt_leaves(const float4 * __restrict__ leaves, uint32_t n_leaves, uint32_t k, float4 * __restrict__ nodes) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_leaves) return;
    const uint32_t idx = bt_index(0, j, k, n_leaves);
    nodes[2 * size_t(idx)] = leaves[2 * size_t(j)];
    nodes[2 * size_t(idx) + 1] = leaves[2 * size_t(j) + 1];
}

// a box that holds a NaN overlaps every query box (AABB::intersect, aabb.cpp:3-10, answers true when a compare is
// unordered), so the union must keep the NaN or the leaf below would be pruned: not fminf/fmaxf
__device__ __forceinline__ float lo_of(float a, float b) { return (a != a || b != b) ? __int_as_float(0x7fc00000) : fminf(a, b); }
__device__ __forceinline__ float hi_of(float a, float b) { return (a != a || b != b) ? __int_as_float(0x7fc00000) : fmaxf(a, b); }

// level >= 1: box = union of the two children (written by the passes below), link = index after the subtree,
// count = -(index of the second child)
__global__ void k_bt_level(uint32_t level, uint32_t n_blocks, uint32_t n_leaves, uint32_t k, float4 * __restrict__ nodes) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_blocks || !bt_is_node(level, j, n_leaves)) return;
    const uint32_t idx = bt_index(level, j, k, n_leaves);
    const uint32_t c1 = idx + 1, c2 = idx + (1u << level);
    const float4 lo1 = nodes[2 * size_t(c1)], hi1 = nodes[2 * size_t(c1) + 1];
    const float4 lo2 = nodes[2 * size_t(c2)], hi2 = nodes[2 * size_t(c2) + 1];
    float4 lo, hi;
    lo.x = lo_of(lo1.x, lo2.x); lo.y = lo_of(lo1.y, lo2.y); lo.z = lo_of(lo1.z, lo2.z);
    hi.x = hi_of(hi1.x, hi2.x); hi.y = hi_of(hi1.y, hi2.y); hi.z = hi_of(hi1.z, hi2.z);
    lo.w = __int_as_float(int(idx + 2u * bt_leaves_of(level, j, n_leaves) - 1u));
    hi.w = __int_as_float(-int(c2));
    nodes[2 * size_t(idx)] = lo;
    nodes[2 * size_t(idx) + 1] = hi;
}

} // namespace

// leaves: n_leaves leaf records (2 float4 each) in ascending id; nodes: room for 2 * n_leaves - 1 records
void launch_tree_build(const float4 * leaves, uint32_t n_leaves, float4 * nodes, cudaStream_t stream) {
    if (n_leaves == 0) return;
    const uint32_t k = bt_levels(n_leaves);
    k_bt_leaves<<<(n_leaves + 255) / 256, 256, 0, stream>>>(leaves, n_leaves, k, nodes);
    for (uint32_t level = 1; level <= k; ++level) {
        const uint32_t n_blocks = uint32_t((uint64_t(n_leaves) + (1ull << level) - 1) >> level);
        k_bt_level<<<(n_blocks + 255) / 256, 256, 0, stream>>>(level, n_blocks, n_leaves, k, nodes);
    }
}

} // namespace prt
