// step_common.cuh -- device helpers and compile-time constants shared by the step kernels
// (k_step.cu, k_step_stream.cu, k_step_warp.cu).
#ifndef PRT_STEP_COMMON_CUH
#define PRT_STEP_COMMON_CUH

#include "kernels.cuh"

namespace prt {

// cp.async (LDGSTS): 16-byte global -> shared copies that need no destination registers; groups complete in order
__device__ __forceinline__ void cp_async16(void * smem_dst, const void * gmem_src) {
    const uint32_t d = uint32_t(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" :: "r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }

__device__ __forceinline__ bool node_overlaps(float4 lo, float4 hi, Box const & q) {   // AABB::intersect aabb.cpp:3-10
    return !(lo.x > q.hi.x || hi.x < q.lo.x || lo.y > q.hi.y || hi.y < q.lo.y || lo.z > q.hi.z || hi.z < q.lo.z);
}

// uniform xz grid of the JACOBI pass: cell coordinate of a world coordinate, clamped to the grid (clamping is safe,
// the box-box test stays exact)
__device__ __forceinline__ int dyn_cell(float v, float origin, float inv_cell, uint32_t g) {
    const float f = floorf((v - origin) * inv_cell);
    int c = (f >= float(g)) ? int(g) - 1 : (f < 0.0f ? 0 : int(f));
    if (!(f == f)) c = 0;                                          // NaN boxes land in cell 0; they overlap everything anyway
    return c;
}

constexpr int DYN_CAP = 16;      // capsule neighbours buffered per lane and pass of the JACOBI pass

// JACOBI pass: the (at most DYN_CAP) smallest record indices > `last` of the characters other than `self` whose STORED fat
// box overlaps `b`, ascending, into col[0], col[32], col[64], ... (one column of a [DYN_CAP][32] shared array); `more` says
// that there are further ones (call again with last = the largest returned).  Every record lives in the cell of its stored
// box's lower corner, so the cells to look at are those of `b` and `reach` cells before them.
struct DynGrid {
    const float4 * dyn_records; const uint32_t * dyn_cell_start, * dyn_items, * dyn_reach;
    uint32_t dyn_gx, dyn_gz; float dyn_origin_x, dyn_origin_z, dyn_inv_cell;
};
__device__ __forceinline__ DynGrid dyn_grid_of(StepParams const & q) {
    DynGrid g;
    g.dyn_records = q.dyn_records; g.dyn_cell_start = q.dyn_cell_start; g.dyn_items = q.dyn_items; g.dyn_reach = q.dyn_reach;
    g.dyn_gx = q.dyn_gx; g.dyn_gz = q.dyn_gz; g.dyn_origin_x = q.dyn_origin_x; g.dyn_origin_z = q.dyn_origin_z; g.dyn_inv_cell = q.dyn_inv_cell;
    return g;
}
// (not inlined: it is called from two places of the pooled kernel, whose instruction footprint matters)
static __device__ __noinline__ uint32_t dyn_collect(DynGrid p, Box b, uint32_t self, long long last, uint32_t * col, bool & more) {
    const int rx = int(__ldg(p.dyn_reach)), rz = int(__ldg(p.dyn_reach + 1));
    int cx0 = dyn_cell(b.lo.x, p.dyn_origin_x, p.dyn_inv_cell, p.dyn_gx) - rx, cz0 = dyn_cell(b.lo.z, p.dyn_origin_z, p.dyn_inv_cell, p.dyn_gz) - rz;
    const int cx1 = dyn_cell(b.hi.x, p.dyn_origin_x, p.dyn_inv_cell, p.dyn_gx), cz1 = dyn_cell(b.hi.z, p.dyn_origin_z, p.dyn_inv_cell, p.dyn_gz);
    if (cx0 < 0) cx0 = 0;
    if (cz0 < 0) cz0 = 0;
    uint32_t ncand = 0;
    more = false;
    for (int cz = cz0; cz <= cz1; ++cz) {
        // the cells of one grid row are contiguous in the item array
        const uint32_t beg = __ldg(p.dyn_cell_start + uint32_t(cz) * p.dyn_gx + uint32_t(cx0));
        const uint32_t end = __ldg(p.dyn_cell_start + uint32_t(cz) * p.dyn_gx + uint32_t(cx1) + 1);
        // four items per trip: their indices, then their stored boxes, are independent loads in flight (one dependent load
        // after the other made this search a good part of a batch's latency)
        for (uint32_t k4 = beg; k4 < end; k4 += 4) {
            uint32_t jj[4];
            float4 blo[4], bhi[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) jj[u] = __ldg(p.dyn_items + (k4 + u < end ? k4 + u : k4));
#pragma unroll
            for (int u = 0; u < 4; ++u) { blo[u] = __ldg(p.dyn_records + 4 * size_t(jj[u]) + 2); bhi[u] = __ldg(p.dyn_records + 4 * size_t(jj[u]) + 3); }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint32_t j = jj[u];
                if (k4 + u >= end || j == self || (long long)j <= last) continue;
                if (!node_overlaps(blo[u], bhi[u], b)) continue;
                if (ncand == uint32_t(DYN_CAP)) {                 // keep the DYN_CAP smallest, sorted
                    more = true;
                    if (j > col[32 * (DYN_CAP - 1)]) continue;
                    --ncand;
                }
                uint32_t s = ncand;
                while (s > 0 && col[32 * (s - 1)] > j) { col[32 * s] = col[32 * (s - 1)]; --s; }
                col[32 * s] = j;
                ++ncand;
            }
        }
    }
    return ncand;
}

// Conservative per-triangle pre-cull.  A contact needs `inside || intersects` (collision_system.cpp:101):
//   intersects -> a point of the triangle lies within `radius` of `center`, and `center` lies on the segment AB;
//   inside     -> point0 = center - n*dist lies in the triangle with 0 <= dist < dmax (the range cull :71-72), so
//                 center lies within dmax*|n_axis| of the triangle's extent along every axis.
// Hence if the triangle's AABB, grown per axis by max(radius, dmax*|n_axis|) + eps, misses the AABB of AB, the exact
// test cannot report a contact at this pose and skipping it changes nothing.  eps (StepParams::qr_eps, >= 64 ulp of
// the largest world coordinate) absorbs the rounding of the exact test's intermediate points.  Any NaN makes every
// compare false => "not rejected", i.e. the exact path decides.
__device__ __forceinline__ bool quick_reject(V3 smin, V3 smax, float radius, float dmax, float eps,
                                             float4 t0, float4 t1, float4 t2) {
    const float ex = fmaxf(radius, dmax * fabsf(t0.w)) + eps;
    const float ey = fmaxf(radius, dmax * fabsf(t1.w)) + eps;
    const float ez = fmaxf(radius, dmax * fabsf(t2.w)) + eps;
    const float lox = fminf(fminf(t0.x, t1.x), t2.x), hix = fmaxf(fmaxf(t0.x, t1.x), t2.x);
    const float loy = fminf(fminf(t0.y, t1.y), t2.y), hiy = fmaxf(fmaxf(t0.y, t1.y), t2.y);
    const float loz = fminf(fminf(t0.z, t1.z), t2.z), hiz = fmaxf(fmaxf(t0.z, t1.z), t2.z);
    return (lox - ex > smax.x) || (hix + ex < smin.x) || (loy - ey > smax.y) || (hiy + ey < smin.y) ||
           (loz - ez > smax.z) || (hiz + ez < smin.z);
}

// Cold per-lane state of k_step_stream lives in shared memory ([field][lane], conflict free)
enum Park { P_VEL = 0, P_STEPV = 3, P_GN = 6, P_SA1 = 9, P_SA2 = 12, P_SB1 = 15, P_SB2 = 18,
            P_PREVA = 21, P_PREVB = 24, P_QR_EPS = 27, P_COUNT = 28 };
struct Parked {
    float (*s)[32];
    uint32_t tid;
    __device__ __forceinline__ V3 get(int f) const { return v3(s[f][tid], s[f + 1][tid], s[f + 2][tid]); }
    __device__ __forceinline__ void put(int f, V3 v) const { s[f][tid] = v.x; s[f + 1][tid] = v.y; s[f + 2][tid] = v.z; }
    __device__ __forceinline__ CapsuleFrame frame(float radius) const {
        CapsuleFrame fr;
        fr.sa1 = get(P_SA1); fr.sa2 = get(P_SA2); fr.sb1 = get(P_SB1); fr.sb2 = get(P_SB2);
        fr.radius = radius;
        return fr;
    }
};

constexpr int STEP_BLOCK = 32;      // one warp per block: a finished warp frees its slot at once (2.79 -> 2.71 ms at C4 vs 128)
constexpr int MODE_STATIC = 0, MODE_LITERAL = 1, MODE_JACOBI = 2;
constexpr int XC_CAP = 12;       // leaves per cached region slot (k_xc_refresh / k_step_stream)
constexpr int TRI_RING = 2;      // triangles in flight per lane (cp.async ring in shared memory)
constexpr int LEAF_CAP = 8;      // candidate leaves buffered per lane and pass (more => another pass)

// per-kernel launchers (each kernel lives in its own translation unit); launch_step() in kernels.cu dispatches
void launch_k_step(StepParams const & p, bool traced, int mode, uint32_t grid, cudaStream_t stream);
void launch_k_step_warp(StepParams const & p, int mode, cudaStream_t stream);
void launch_k_xc_refresh(StepParams const & p, cudaStream_t stream);
void launch_k_step_stream(StepParams const & p, uint32_t blocks, bool gated, cudaStream_t stream);
void launch_k_step_pool(StepParams const & p, uint32_t blocks, bool gated, int mode, cudaStream_t stream);

} // namespace prt
#endif
