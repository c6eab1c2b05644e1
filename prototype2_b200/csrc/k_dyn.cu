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
                           const float4 * prev_records, float4 * records, DynPeers peers, float dt, float gravity, float margin, uint32_t init,
                           uint32_t * reach) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0 && reach) { reach[0] = 0u; reach[1] = 0u; }         // this frame's grid build starts from no reach
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

// The exchange across processes.  A warp moves one 512-byte piece (32 x float4) of this rank's slice to ONE destination;
// consecutive pieces go to consecutive destinations, so at any moment the stores of a rank are spread over all its peers
// (and no peer is the target of everybody at once); a few hundred fat blocks walk the pieces with four stores per thread
// in flight and ONE system-wide fence per block (a fence waits for the acknowledgements of the block's stores: one
// thread block per 256 stores spent its life in that wait).  The last block writes the arrival words.
__global__ void __launch_bounds__(256) k_dyn_push(const float4 * __restrict__ records, uint32_t first, uint32_t n, DynPeers peers) {
    const uint32_t per = 4u * n;                                   // float4 of the slice
    const uint32_t pieces_per_dst = (per + 31u) / 32u;
    const uint32_t pieces = pieces_per_dst * peers.n;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
    const float4 * src = records + 4 * size_t(first);
    uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    for (; w + 3u * warps < pieces; w += 4u * warps) {
        float4 v[4]; uint32_t e[4], k[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const uint32_t piece = w + uint32_t(u) * warps;
            k[u] = piece % peers.n; e[u] = (piece / peers.n) * 32u + lane;
            if (e[u] < per) v[u] = src[e[u]];
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) if (e[u] < per) peers.records[k[u]][4 * size_t(first) + e[u]] = v[u];
    }
    for (; w < pieces; w += warps) {
        const uint32_t k = w % peers.n, e = (w / peers.n) * 32u + lane;
        if (e < per) peers.records[k][4 * size_t(first) + e] = src[e];
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

// Counting sort of the stored boxes by cell, three launches and no memset:
//   k_dyn_count   cell_count[cell]++ (zero on entry: k_dyn_scan leaves it so) and the reach
//   k_dyn_scan    exclusive prefix sum -> cell_start and a copy in `cursor`; zeroes cell_count for the next frame
//   k_dyn_fill    items[cursor[cell]++] = character
__global__ void k_dyn_count(const float4 * __restrict__ records, uint32_t n, float ox, float oz, float inv_cell, uint32_t gx, uint32_t gz,
                            uint32_t * __restrict__ cell_count, uint32_t * __restrict__ reach) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const float4 lo = records[4 * size_t(j) + 2], hi = records[4 * size_t(j) + 3];
    const int x0 = dyn_cell(lo.x, ox, inv_cell, gx), z0 = dyn_cell(lo.z, oz, inv_cell, gz);
    atomicAdd(cell_count + (uint32_t(z0) * gx + uint32_t(x0)), 1u);
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

__global__ void k_dyn_fill(const float4 * __restrict__ records, uint32_t n, float ox, float oz, float inv_cell, uint32_t gx, uint32_t gz,
                           uint32_t * __restrict__ cursor, uint32_t * __restrict__ items) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const float4 lo = records[4 * size_t(j) + 2];
    const int x0 = dyn_cell(lo.x, ox, inv_cell, gx), z0 = dyn_cell(lo.z, oz, inv_cell, gz);
    items[atomicAdd(cursor + (uint32_t(z0) * gx + uint32_t(x0)), 1u)] = j;
}

// Exclusive prefix sum of `count[0 .. n)` into `start[0 .. n]` (start[n] = total) and `cursor[0 .. n)`, count := 0.
// Single pass over several blocks (2048 cells each) with a decoupled look-back: a block publishes the sum of its tile,
// then warp 0 adds up the published words of its predecessors, 32 at a time, until it meets one that already carries an
// inclusive prefix.  Words are (launch number * 2 + is-prefix) << 32 | value, written and read in one 64-bit access, so
// nothing is reset between launches; tile numbers come from a ticket counter (a block only ever waits for blocks that
// are already running), `ticket_base` = the tickets all earlier launches took.
constexpr uint32_t SCAN_THREADS = 512, SCAN_TILE = 4 * SCAN_THREADS;

__global__ void __launch_bounds__(SCAN_THREADS) k_dyn_scan(uint32_t * __restrict__ count, uint32_t * __restrict__ start, uint32_t * __restrict__ cursor,
                                                           uint32_t n, unsigned long long * state, uint32_t * ticket, uint32_t ticket_base,
                                                           uint32_t launch_no) {
    __shared__ uint32_t s_warp[SCAN_THREADS / 32];
    __shared__ uint32_t s_tile, s_excl;
    const uint32_t t = threadIdx.x, lane = t & 31u, warp = t >> 5;
    if (t == 0) s_tile = atomicAdd(ticket, 1u) - ticket_base;
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint32_t base = tile * SCAN_TILE + 4u * t;
    uint32_t v[4] = {0u, 0u, 0u, 0u};
    if (base + 3u < n) { const uint4 q = *reinterpret_cast<const uint4 *>(count + base); v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w; }
    else { for (uint32_t k = 0; k < 4u; ++k) if (base + k < n) v[k] = count[base + k]; }
    const uint32_t sum = (v[0] + v[1]) + (v[2] + v[3]);
    uint32_t incl = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t u = __shfl_up_sync(0xffffffffu, incl, d); if (int(lane) >= d) incl += u; }
    if (lane == 31u) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = lane < SCAN_THREADS / 32 ? s_warp[lane] : 0u;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t u = __shfl_up_sync(0xffffffffu, w, d); if (int(lane) >= d) w += u; }
        if (lane < SCAN_THREADS / 32) s_warp[lane] = w;                        // inclusive over the warps
        const uint32_t total = __shfl_sync(0xffffffffu, w, SCAN_THREADS / 32 - 1);
        const unsigned long long tag = (unsigned long long)(launch_no) << 1;
        volatile unsigned long long * st = state;
        if (lane == 0) st[tile] = ((tag | (tile == 0 ? 1ull : 0ull)) << 32) | total;
        uint32_t excl = 0;
        if (tile > 0) {
            int end = int(tile) - 1;                                            // nearest predecessor not yet added
            for (;;) {
                const int idx = end - int(lane);
                unsigned long long word = (tag | 1ull) << 32;                  // before tile 0: prefix 0
                if (idx >= 0) { do { word = st[idx]; } while ((word >> 33) != (unsigned long long)(launch_no)); }
                const uint32_t is_prefix = uint32_t(word >> 32) & 1u;
                const uint32_t have = __ballot_sync(0xffffffffu, is_prefix != 0u);
                const uint32_t upto = have ? uint32_t(__ffs(int(have)) - 1) : 31u;   // add lanes 0 .. upto
                uint32_t part = lane <= upto ? uint32_t(word) : 0u;
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) part += __shfl_xor_sync(0xffffffffu, part, d);
                excl += part;
                if (have) break;
                end -= 32;
            }
            if (lane == 0) st[tile] = ((tag | 1ull) << 32) | (unsigned long long)(excl + total);
        }
        if (lane == 0) {
            s_excl = excl;
            if (tile == n / SCAN_TILE) start[n] = excl + total;                  // (cells at and past n count as empty)
        }
    }
    __syncthreads();
    uint32_t run = s_excl + (incl - sum) + (warp ? s_warp[warp - 1] : 0u);
    if (base + 3u < n) {
        uint4 q; q.x = run; q.y = run + v[0]; q.z = q.y + v[1]; q.w = q.z + v[2];
        *reinterpret_cast<uint4 *>(start + base) = q;
        *reinterpret_cast<uint4 *>(cursor + base) = q;
        *reinterpret_cast<uint4 *>(count + base) = make_uint4(0u, 0u, 0u, 0u);
    } else {
        for (uint32_t k = 0; k < 4u; ++k) if (base + k < n) { start[base + k] = run; cursor[base + k] = run; count[base + k] = 0u; run += v[k]; }
    }
}

} // namespace

void launch_dyn_prep(const prtc_transform * tf, const prtc_char_physics * ph, const CapsuleDev * capsules, uint32_t n_capsules, uint32_t n,
                     uint32_t first, const float4 * prev_records, float4 * records, DynPeers const & peers, float dt, float gravity, float margin,
                     bool init, uint32_t * reach, cudaStream_t stream) {
    const uint32_t blocks = n ? (n + 127) / 128 : 1u;             // (a rank without characters still has to arrive, and to reset the reach)
    k_dyn_prep<<<blocks, 128, 0, stream>>>(tf, ph, capsules, n_capsules, n, first, prev_records, records, peers, dt, gravity, margin, init ? 1u : 0u, reach);
}

void launch_dyn_push(const float4 * records, uint32_t first, uint32_t n, DynPeers const & peers, cudaStream_t stream) {
    const uint32_t pieces = ((4u * n + 31u) / 32u) * peers.n;
    const uint32_t want = (pieces + 31u) / 32u;                    // blocks if every warp moved four pieces
    const uint32_t blocks = std::max(1u, std::min(want, 148u * 4u));
    k_dyn_push<<<blocks, 256, 0, stream>>>(records, first, n, peers);
}

void launch_dyn_wait(const uint32_t * flags, uint32_t n_ranks, uint32_t value, unsigned long long * counters, cudaStream_t stream) {
    k_dyn_wait<<<1, 32, 0, stream>>>(flags, n_ranks, value, counters);
}

int launch_dyn_grid_build(const float4 * records, uint32_t n, float ox, float oz, float inv_cell, uint32_t gx, uint32_t gz,
                          uint32_t * cell_count, uint32_t * cell_start, uint32_t * cursor, uint32_t * items, uint32_t * reach,
                          unsigned long long * scan_state, uint32_t & scan_tickets, uint32_t & scan_launches, cudaStream_t stream) {
    const uint32_t cells = gx * gz;
    if (n) k_dyn_count<<<(n + 127) / 128, 128, 0, stream>>>(records, n, ox, oz, inv_cell, gx, gz, cell_count, reach);
    const uint32_t tiles = cells / SCAN_TILE + 1u;                 // (the tile of index `cells` writes the total)
    scan_launches = (scan_launches + 1u) & 0x3fffffffu;
    if (scan_launches == 0u) scan_launches = 1u;
    k_dyn_scan<<<tiles, SCAN_THREADS, 0, stream>>>(cell_count, cell_start, cursor, cells, scan_state, reinterpret_cast<uint32_t *>(scan_state + DYN_SCAN_WORDS),
                                                   scan_tickets, scan_launches);
    scan_tickets += tiles;
    if (n) k_dyn_fill<<<(n + 127) / 128, 128, 0, stream>>>(records, n, ox, oz, inv_cell, gx, gz, cursor, items);
    return n ? 3 : 1;
}

} // namespace prt
