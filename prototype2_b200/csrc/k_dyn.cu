// k_dyn.cu -- the JACOBI dynamic pass around the step kernels: per-character records and the uniform grid over the stored boxes
// Compile with -fmad=false: results must be bit-identical to the CPU reference (see prt_math.cuh).
//
//   k_dyn_prep    K1 of the pass: per own character the phase-1 predicted box (physics_system.cpp:263-273), the stored fat box
//                 hysteresis of DynamicAABBTree::update / insertLeaf (aabb_tree.cpp:102-111,140-141) and the start-of-frame
//                 segment; record = 4 float4: (A.xyz, radius) (B.xyz, 0) (stored.lo, 0) (stored.hi, 0)
//   k_dyn_grid    counting sort of ALL characters' stored boxes into a uniform xz grid: every box lives in exactly ONE cell --
//                 the cell of its lower corner -- so the item array has exactly n entries (nothing to size from a device
//                 count, no host read-back); how many cells a box can reach beyond its own is tracked per frame as a
//                 device-side maximum (`reach`), and a query simply starts that many cells earlier (step_common.cuh dyn_collect)
//   k_dyn_scan    exclusive prefix sum of the cell counts, one block (the grid has at most 2^20 cells; 81 k at C5)
// Everything runs on the context's stream without the host looking at any of it.
#include <algorithm>

#include "step_common.cuh"

namespace prt {

namespace {

__global__ void k_dyn_prep(const prtc_transform * __restrict__ tf, const prtc_char_physics * __restrict__ ph,
                           const CapsuleDev * __restrict__ capsules, uint32_t n_capsules, uint32_t n, uint32_t first,
                           const float4 * prev_records, float4 * records, DynPeers peers, float dt, float gravity, float margin, uint32_t init) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
    uint32_t cid = ph[i].collider;
    if (cid >= n_capsules) cid = 0;               // unknown capsule id: never dereferenced (the step kernel reports it, counters[6])
    const CapsuleDev cap = capsules[cid];
    Quat q; q.x = tf[i].rotation[0]; q.y = tf[i].rotation[1]; q.z = tf[i].rotation[2]; q.w = tf[i].rotation[3];
    const Rot3 R = rotation_of(q);
    const CapsuleFrame fr = capsule_frame(cap, R);
    const V3 pos = v3(tf[i].position[0], tf[i].position[1], tf[i].position[2]);
    V3 A, B;
    segment_ends(fr, pos, A, B);
    Box e = capsule_box(A, B, cap.radius);
    const V3 vel = v3(ph[i].velocity[0], ph[i].velocity[1], ph[i].velocity[2]);
    const V3 d = vel + (v3(0.0f, -1.0f, 0.0f) * gravity) * dt;
    const V3 pos2 = ((R.c0 * d.x + R.c1 * d.y) + R.c2 * d.z) + pos;
    V3 A2, B2;
    segment_ends(fr, pos, pos2, A2, B2);
    e = box_union(e, capsule_box(A2, B2, cap.radius));
    float4 * rec = records + 4 * size_t(first + i);
    Box stored;
    if (init) {                                                    // the leaf is created at the identity pose (physics_system.cpp:36-39)
        Quat qi; qi.x = qi.y = qi.z = 0.0f; qi.w = 1.0f;
        V3 iA, iB;
        segment_ends(capsule_frame(cap, rotation_of(qi)), v3(0.0f), iA, iB);
        const Box ib = capsule_box(iA, iB, cap.radius);
        stored.lo = ib.lo - margin; stored.hi = ib.hi + margin;
    } else {                                                       // last frame's stored box (the other buffer when the records are double-buffered)
        const float4 lo = prev_records[4 * size_t(first + i) + 2], hi = prev_records[4 * size_t(first + i) + 3];
        stored.lo = v3(lo.x, lo.y, lo.z); stored.hi = v3(hi.x, hi.y, hi.z);
    }
    if (!box_contains(stored, e)) { stored.lo = e.lo - margin; stored.hi = e.hi + margin; }
    const float4 r0 = make_float4(A.x, A.y, A.z, cap.radius), r1 = make_float4(B.x, B.y, B.z, 0.0f);
    const float4 r2 = make_float4(stored.lo.x, stored.lo.y, stored.lo.z, 0.0f), r3 = make_float4(stored.hi.x, stored.hi.y, stored.hi.z, 0.0f);
    rec[0] = r0; rec[1] = r1; rec[2] = r2; rec[3] = r3;
    // several devices in one process (prtc_multi): the record goes straight into the other devices' arrays too -- peer stores
    // over NVLink from the kernel that computes it, instead of a collective behind it
    for (uint32_t k = 0; k < (peers.push_separately ? 0u : peers.n); ++k) {
        float4 * dst = peers.records[k] + 4 * size_t(first + i);
        dst[0] = r0; dst[1] = r1; dst[2] = r2; dst[3] = r3;
    }
    }
    if (peers.n_flags && !peers.push_separately) {
        // arrival: every block's stores are out (system-wide fence) before it counts itself; the last one tells everybody
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0 && atomicAdd(peers.block_count, 1u) + 1u == gridDim.x) {
            *peers.block_count = 0u;
            __threadfence_system();
            for (uint32_t k = 0; k < peers.n_flags; ++k) *reinterpret_cast<volatile uint32_t *>(peers.flags[k]) = peers.flag_value;
            __threadfence_system();
        }
    }
}

// The exchange across processes: one thread per (destination, float4 of this rank's slice) -- eight times the threads of
// k_dyn_prep, every one with a single 16-byte peer store in flight -- then the arrival words, written by the last block.
__global__ void k_dyn_push(const float4 * __restrict__ records, uint32_t first, uint32_t n, DynPeers peers) {
    const size_t per = 4 * size_t(n);
    const size_t t = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t < per * peers.n) {
        const uint32_t k = uint32_t(t / per);
        const size_t e = 4 * size_t(first) + t % per;
        peers.records[k][e] = records[e];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0 && atomicAdd(peers.block_count, 1u) + 1u == gridDim.x) {
        *peers.block_count = 0u;
        __threadfence_system();
        for (uint32_t k = 0; k < peers.n_flags; ++k) *reinterpret_cast<volatile uint32_t *>(peers.flags[k]) = peers.flag_value;
        __threadfence_system();
    }
}

__global__ void k_dyn_wait(const uint32_t * flags, uint32_t n_ranks, uint32_t value, unsigned long long * counters) {
    const uint32_t r = threadIdx.x;
    if (r >= n_ranks) return;
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    // (frame numbers only grow; a rank that is AHEAD has a larger number)
    while (int32_t(*reinterpret_cast<const volatile uint32_t *>(flags + r) - value) < 0) {
        __nanosleep(100);
        unsigned long long t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > 4000000000ull) { if (counters) atomicAdd(counters + 7, 1ull); break; }     // 4 s: a peer is gone; never hang the GPU
    }
    __threadfence_system();
}

// pass 0 counts (and tracks the reach), pass 1 fills; `cell_count` is zero before either pass
__global__ void k_dyn_grid(const float4 * __restrict__ records, uint32_t n, float ox, float oz, float inv_cell, uint32_t gx, uint32_t gz,
                           uint32_t * __restrict__ cell_count, const uint32_t * __restrict__ cell_start, uint32_t * __restrict__ items,
                           uint32_t * __restrict__ reach, int fill) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const float4 lo = records[4 * size_t(j) + 2], hi = records[4 * size_t(j) + 3];
    const int x0 = dyn_cell(lo.x, ox, inv_cell, gx), z0 = dyn_cell(lo.z, oz, inv_cell, gz);
    const uint32_t cell = uint32_t(z0) * gx + uint32_t(x0);
    const uint32_t slot = atomicAdd(cell_count + cell, 1u);
    if (fill) { items[cell_start[cell] + slot] = j; return; }
    // how far beyond its own cell the box reaches; a box with a non-finite corner overlaps every query box
    // (AABB::intersect is written with negated compares, aabb.cpp:3-10), so it makes every query look at every cell
    uint32_t rx, rz;
    const float s = (lo.x + lo.z) + (hi.x + hi.z);
    if (!(s - s == 0.0f)) { rx = gx; rz = gz; }
    else {
        rx = uint32_t(dyn_cell(hi.x, ox, inv_cell, gx) - x0);
        rz = uint32_t(dyn_cell(hi.z, oz, inv_cell, gz) - z0);
    }
    if (rx > reach[0]) atomicMax(reach + 0, rx);
    if (rz > reach[1]) atomicMax(reach + 1, rz);
}

// exclusive prefix sum of `count[0 .. n)` into `start[0 .. n]` (start[n] = total), one block of 1024 threads: every thread
// sums a contiguous run, the block scans the 1024 run sums, every thread writes its run
__global__ void __launch_bounds__(1024) k_dyn_scan(const uint32_t * __restrict__ count, uint32_t * __restrict__ start, uint32_t n) {
    __shared__ uint32_t s_warp[32];
    const uint32_t t = threadIdx.x;
    const uint32_t per = (n + 1023u) / 1024u;
    const uint32_t beg = t * per < n ? t * per : n, end = beg + per < n ? beg + per : n;
    uint32_t sum = 0;
    for (uint32_t k = beg; k < end; ++k) sum += count[k];
    uint32_t incl = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d); if (int(t & 31u) >= d) incl += v; }
    if ((t & 31u) == 31u) s_warp[t >> 5] = incl;
    __syncthreads();
    if (t < 32u) {
        uint32_t w = s_warp[t];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, w, d); if (int(t) >= d) w += v; }
        s_warp[t] = w;
    }
    __syncthreads();
    uint32_t run = incl - sum + ((t >> 5) ? s_warp[(t >> 5) - 1] : 0u);
    for (uint32_t k = beg; k < end; ++k) { start[k] = run; run += count[k]; }
    if (t == 1023u) start[n] = s_warp[31];
}

} // namespace

void launch_dyn_prep(const prtc_transform * tf, const prtc_char_physics * ph, const CapsuleDev * capsules, uint32_t n_capsules, uint32_t n,
                     uint32_t first, const float4 * prev_records, float4 * records, DynPeers const & peers, float dt, float gravity, float margin,
                     bool init, cudaStream_t stream) {
    if (n == 0 && !peers.n_flags) return;
    const uint32_t blocks = n ? (n + 127) / 128 : 1u;             // (a rank without characters still has to arrive)
    k_dyn_prep<<<blocks, 128, 0, stream>>>(tf, ph, capsules, n_capsules, n, first, prev_records, records, peers, dt, gravity, margin, init ? 1u : 0u);
    if (peers.push_separately) {
        const size_t threads = std::max<size_t>(1, 4 * size_t(n) * peers.n);
        k_dyn_push<<<unsigned((threads + 255) / 256), 256, 0, stream>>>(records, first, n, peers);
    }
}

void launch_dyn_wait(const uint32_t * flags, uint32_t n_ranks, uint32_t value, unsigned long long * counters, cudaStream_t stream) {
    k_dyn_wait<<<1, 32, 0, stream>>>(flags, n_ranks, value, counters);
}

int launch_dyn_grid_build(const float4 * records, uint32_t n, float ox, float oz, float inv_cell, uint32_t gx, uint32_t gz,
                          uint32_t * cell_count, uint32_t * cell_start, uint32_t * items, uint32_t * reach, cudaStream_t stream) {
    const size_t cells = size_t(gx) * gz;
    cudaMemsetAsync(cell_count, 0, cells * sizeof(uint32_t), stream);
    cudaMemsetAsync(reach, 0, 2 * sizeof(uint32_t), stream);
    if (n) k_dyn_grid<<<(n + 127) / 128, 128, 0, stream>>>(records, n, ox, oz, inv_cell, gx, gz, cell_count, nullptr, nullptr, reach, 0);
    k_dyn_scan<<<1, 1024, 0, stream>>>(cell_count, cell_start, uint32_t(cells));
    cudaMemsetAsync(cell_count, 0, cells * sizeof(uint32_t), stream);
    if (n) k_dyn_grid<<<(n + 127) / 128, 128, 0, stream>>>(records, n, ox, oz, inv_cell, gx, gz, cell_count, cell_start, items, reach, 1);
    return n ? 3 : 1;
}

} // namespace prt
