// k_step_stream.cu -- the persistent, lane-decoupled static-pass kernel and the refresh of its per-slot leaf cache
// Compile with -fmad=false: results must be bit-identical to the CPU reference (see prt_math.cuh).
#include "step_common.cuh"

namespace prt {

namespace {

// k_xc_refresh: keeps the per-slot frame cache of k_step_stream valid for the coming frame.  One thread per
// character: the region its frame can need (box of the current pose united with the pose after the pre-stepped
// velocity, +0.25 for push-outs); if the slot's stored region (built for this tree version) already contains it, done --
// resting or slow characters never walk the tree again; else one stackless walk with the region grown by xc_margin
// collects the overlapping leaves (emission order) and the slot is rewritten.  The list is a pure function of
// (tree, region): it does not matter which character occupied the slot before, and k_step_stream re-checks
// "query box inside region" at every substep, so a stale or too-small region can only cost time, never correctness.
//
// Top-level staging (north_star: "shared-memory or TMA staging of top-level nodes").  Every walk starts with the same
// few hundred nodes, so a block that has at least one slot to refill brings the top-of-tree table (all nodes of depth
// <= D, <= TOP_CAP entries, same preorder -- bvh_builder.h build_top_table) into shared memory with ONE bulk copy of the
// TMA unit (cp.async.bulk global -> shared, completion counted on an mbarrier; SASS: UBLKCP + SYNCS), issued by one thread
// while the others are still computing their regions.  The walk then alternates between two cursors: `k` runs over the
// table in shared memory; an overlapping entry of depth D (or a leaf) opens the range [node, node_end) of the full array,
// which is walked exactly as before and hands back to the table when it is exhausted.  Same node visit order, same
// leaf sequence.  A block whose slots are all still valid (the steady state) never issues the copy.
constexpr int XC_BLOCK = 512;

__device__ __forceinline__ void mbar_init(uint64_t * bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(uint32_t(__cvta_generic_to_shared(bar))), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void bulk_load(void * smem_dst, const void * gmem_src, uint32_t bytes, uint64_t * bar) {
    const uint32_t b = uint32_t(__cvta_generic_to_shared(bar));
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(b), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(uint32_t(__cvta_generic_to_shared(smem_dst))), "l"(gmem_src), "r"(bytes), "r"(b) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t * bar, uint32_t parity) {
    const uint32_t b = uint32_t(__cvta_generic_to_shared(bar));
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done) : "r"(b), "r"(parity) : "memory");
    }
}

__global__ void __launch_bounds__(XC_BLOCK) k_xc_refresh(StepParams p, int reuse) {
    __shared__ __align__(128) float4 s_top[2 * TOP_CAP];
    __shared__ __align__(8) uint64_t s_bar;
    const bool staged = p.top != nullptr;                         // uniform over the grid
    if (staged && threadIdx.x == 0) mbar_init(&s_bar, 1);
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    bool walk = i < p.n;
    Box R; R.lo = v3(0.0f); R.hi = v3(0.0f);
    float ph_vx = 0.0f, ph_vz = 0.0f, ph_mx = 0.0f, ph_mz = 0.0f;
    const size_t s = size_t(p.xc_first) + i;
    if (walk) {
        const prtc_transform * tf = p.transforms + i;
        const prtc_char_physics * ph = p.physics + i;
        const V3 pos = v3(tf->position[0], tf->position[1], tf->position[2]);
        Quat q; q.x = tf->rotation[0]; q.y = tf->rotation[1]; q.z = tf->rotation[2]; q.w = tf->rotation[3];
        V3 vel = v3(ph->velocity[0], ph->velocity[1], ph->velocity[2]);
        uint32_t cid = ph->collider;
        if (cid >= p.n_capsules) cid = 0;
        const CapsuleDev cap = p.capsules[cid];
        const CapsuleFrame fr = capsule_frame(cap, rotation_of(q));
        if (ph->grounded) vel = vel + ((-p.stick * p.gravity) * p.dt) * v3(ph->ground_normal[0], ph->ground_normal[1], ph->ground_normal[2]);
        ph_vx = ph->velocity[0]; ph_vz = ph->velocity[2]; ph_mx = ph->movement[0]; ph_mz = ph->movement[2];
        vel.x = vel.x + ph->movement[0];
        vel.z = vel.z + ph->movement[2];
        V3 a0, b0, eA, eB;
        segment_ends(fr, pos, a0, b0);
        segment_ends(fr, pos + vel, eA, eB);
        R = box_union(capsule_box(a0, b0, cap.radius), capsule_box(eA, eB, cap.radius));
        R.lo = R.lo - v3(0.25f); R.hi = R.hi + v3(0.25f);
        if (reuse && p.xc_version[s] == p.tree_version) {
            if (p.xc_region[0 * size_t(p.xc_stride) + s] <= R.lo.x && p.xc_region[1 * size_t(p.xc_stride) + s] <= R.lo.y &&
                p.xc_region[2 * size_t(p.xc_stride) + s] <= R.lo.z && p.xc_region[3 * size_t(p.xc_stride) + s] >= R.hi.x &&
                p.xc_region[4 * size_t(p.xc_stride) + s] >= R.hi.y && p.xc_region[5 * size_t(p.xc_stride) + s] >= R.hi.z)
                walk = false;
        }
    }
    uint32_t n_top = 0;
    if (staged) {
        // the barrier is initialised (thread 0, above) before anyone can wait on it; the vote doubles as that fence
        if (!__syncthreads_or(walk ? 1 : 0)) return;
        if (threadIdx.x == 0) bulk_load(s_top, p.top, p.n_top * 32u, &s_bar);
        n_top = p.n_top;
    }
    if (!walk) return;
    if (reuse) {
        // A fresh list should last: grow the region by xc_margin and stretch it along the way the character is going to
        // coast -- friction leaves velocity.xz * r after this frame, * r^2 after the next, ... (r = 1 / (1 + dt friction),
        // physics_system.cpp:307), r / (1 - r) of it in total, plus two more frames of the movement input.  Any region
        // is correct (the list is a function of the region); this one keeps a crowd that was set in motion from refilling
        // its slots in waves a few frames later.
        const float r = fdiv(1.0f, 1.0f + p.dt * p.friction);
        const float coast = r < 0.999f ? fdiv(r, 1.0f - r) : 0.0f;
        const float dx = coast * ph_vx + 2.0f * ph_mx, dz = coast * ph_vz + 2.0f * ph_mz;
        if (dx == dx && dz == dz && fabsf(dx) < 4.0f && fabsf(dz) < 4.0f) {
            R.lo.x = R.lo.x + fminf(dx, 0.0f); R.hi.x = R.hi.x + fmaxf(dx, 0.0f);
            R.lo.z = R.lo.z + fminf(dz, 0.0f); R.hi.z = R.hi.z + fmaxf(dz, 0.0f);
        }
        R.lo = R.lo - v3(p.xc_margin); R.hi = R.hi + v3(p.xc_margin);
    }
    if (staged) mbar_wait(&s_bar, 0);
    // Phase 1 (staged only): the table walk alone, all lanes in the same loop.  The overlapping entries at the cut (and
    // leaves above it) are the doors into the full array; they come up in emission order and are kept as 10-bit table
    // indices packed into one 64-bit register (<= 6; a region that opens more than that -- it covers a good part of the
    // world -- falls back to the plain walk).  Phase 2: the ranges behind the doors, in order, walked as before.
    // (Interleaving the two cursors in one loop measured no faster than the plain walk: long-scoreboard stalls per issue
    // fell from 34 to 19, but lanes in "table" and "array" steps serialised, 16.7 -> 13.1 threads per instruction.)
    unsigned long long doors = 0ull;
    uint32_t n_doors = 0, door = 0;
    uint32_t node = 0, node_end = staged ? 0u : p.n_nodes, ncache = 0;
    if (staged) {
        uint32_t k = 0;
        while (k < n_top) {
            const float4 lo = s_top[2 * k];
            const float4 hi = s_top[2 * k + 1];
            const bool ov = node_overlaps(lo, hi, R);
            const int cnt = __float_as_int(hi.w);
            if (cnt == -1) {
                k = ov ? k + 1 : uint32_t(__float_as_int(lo.w));
            } else {
                if (ov) {
                    if (n_doors == 6u) { n_doors = 0; node_end = p.n_nodes; break; }      // too many doors: plain walk
                    doors |= (unsigned long long)k << (10u * n_doors);
                    ++n_doors;
                }
                ++k;
            }
        }
    }
    for (;;) {
        if (node == node_end) {
            if (door == n_doors) break;
            const uint32_t k = uint32_t(doors >> (10u * door)) & 1023u;
            ++door;
            const uint32_t g = uint32_t(__float_as_int(s_top[2 * k].w));
            const int cnt = __float_as_int(s_top[2 * k + 1].w);
            node = cnt == 0 ? g : g + 1;                 // a leaf above the cut is visited itself; below an overlapping
            node_end = cnt == 0 ? g + 1 : uint32_t(-cnt); // root at the cut the walk goes on with its first child
            continue;
        }
        const float4 lo = __ldg(p.nodes + 2 * node);
        const float4 hi = __ldg(p.nodes + 2 * node + 1);
        const bool ov = node_overlaps(lo, hi, R);
        const int cnt = __float_as_int(hi.w);
        if (cnt < 0) {
            node = ov ? node + 1 : uint32_t(__float_as_int(lo.w));
        } else {
            if (ov) {
                if (ncache == XC_CAP) { ncache = 0xFFFFFFFFu; break; }
                p.xc_leaf[size_t(ncache) * p.xc_stride + s] = node;
                ++ncache;
            }
            node = node + 1;
        }
    }
    if (ncache == 0xFFFFFFFFu) { p.xc_version[s] = 0u; return; }   // too many leaves for a slot: this character walks per substep
    p.xc_region[0 * size_t(p.xc_stride) + s] = R.lo.x; p.xc_region[1 * size_t(p.xc_stride) + s] = R.lo.y; p.xc_region[2 * size_t(p.xc_stride) + s] = R.lo.z;
    p.xc_region[3 * size_t(p.xc_stride) + s] = R.hi.x; p.xc_region[4 * size_t(p.xc_stride) + s] = R.hi.y; p.xc_region[5 * size_t(p.xc_stride) + s] = R.hi.z;
    p.xc_count[s] = ncache;
    p.xc_version[s] = p.tree_version;
}

// ------------------------------------------------------------------------------------------------------------
// k_step_stream: the static pass (MODE_STATIC, untraced) as a PERSISTENT, lane-decoupled kernel.
//
// k_step keeps one character per lane and walks all lanes through "substep -> filter -> exact test" in lock step,
// so a warp's cost per substep is set by its slowest lane (ncu: 14 of 32 lanes active per instruction).  Here every
// lane is a small state machine that moves through its character's substeps on its own:
//     LOAD     fetch the next character from a global counter, pre-step, one tree walk for the frame cache
//     SUBSTEP  advance, query box, candidate leaves (cache or ordered walk), start the triangle stream
//     STREAM   filter loop over the cp.async triangle ring until a triangle needs the exact test
// and the only converged section is the expensive one, the exact capsule-triangle test, which runs whenever any lane
// has a pending triangle.  LOAD and SUBSTEP are batched (they run when enough lanes wait for them or nobody has
// anything else to do) so that their divergent code is shared by several lanes.  Arithmetic, order of triangles per
// character and all results are the same as k_step's -- only the interleaving between characters differs.
// ------------------------------------------------------------------------------------------------------------
constexpr int ST_LOAD = 0, ST_SUBSTEP = 1, ST_STREAM = 2, ST_DONE = 3;
constexpr int LOAD_BATCH = 16;       // lanes that must wait for new work before the (expensive) LOAD stage runs
constexpr int SUBSTEP_BATCH = 8;     // same for the SUBSTEP stage

//
// GATED: the batch is still arriving over PCIe while the kernel runs (prtc_step_characters with host buffers).  The
// copy-in stream publishes how many characters have landed in `*gate`; LOAD waits for its own indices.  Every finished
// character is counted (per 1024, then per 65536) so the copy-out stream can start on a chunk the moment its last
// character is written, while later chunks are still being computed.
template <bool GATED>
__global__ void __launch_bounds__(32, 20) k_step_stream(StepParams p) {
    constexpr unsigned FULL = 0xffffffffu;
    __shared__ uint32_t s_leaf[LEAF_CAP][32];              // candidate leaves of the substep (leaf_entry words), emission order
    __shared__ uint32_t s_cache[XC_CAP][32];
    __shared__ float s_region[6][32];
    __shared__ float4 s_tri[TRI_RING][3][32];
    __shared__ float s_park[P_COUNT][32];
    const uint32_t tid = threadIdx.x;
    const Parked park{s_park, tid};
    const unsigned lt_mask = (1u << tid) - 1u;
    uint32_t c_tri = 0, c_node = 0;
    __shared__ uint32_t s_cnt[4][32];                      // rarely bumped counters + ncache live here, not in registers
    s_cnt[0][tid] = 0; s_cnt[1][tid] = 0; s_cnt[2][tid] = 0; s_cnt[3][tid] = 0;
    __shared__ uint32_t s_poll, s_arrived;                 // GATED: arrival count as last polled by lane 0
    if (tid == 0) { s_poll = 0; s_arrived = 0; }
    __syncwarp(FULL);
#define C_SUB s_cnt[0][tid]
#define C_CONTACT s_cnt[1][tid]
#define C_EXACT s_cnt[2][tid]
#define NCACHE s_cnt[3][tid]

    int state = ST_LOAD;
    uint32_t i = 0xFFFFFFFFu;
    V3 pos = v3(0.0f);
    uint32_t grounded = 0, stepsLeft = 0;
    float radius = 0.0f;
    bool cache_ok = false, found = true;
    uint32_t node = 0, nleaf = 0;
    Box qb; qb.lo = v3(0.0f); qb.hi = v3(0.0f);
    Pose ps = make_pose(park.frame(0.0f), pos);
    uint32_t f_leaf = 0, f_cur = 0, f_end = 0, n_fetched = 0, n_consumed = 0;
    bool pending = false;
    float4 t0 = make_float4(0, 0, 0, 0), t1 = t0, t2 = t0;
    const bool use_qr = p.qr_enable != 0u;
    const bool use_cache = !(p.tune & 1u);
    const int load_batch = ((p.tune >> 8) & 0xFFu) ? int((p.tune >> 8) & 0xFFu) : LOAD_BATCH;
    const int substep_batch = ((p.tune >> 16) & 0xFFu) ? int((p.tune >> 16) & 0xFFu) : SUBSTEP_BATCH;
    for (int f = 0; f < P_COUNT; ++f) s_park[f][tid] = 0.0f;

    // A candidate leaf is buffered as its triangle range when that fits one word (first < 2^24, count < 128: every
    // synthetic shape), so the triangle stream does not wait for the node again; else as bit 31 | node index.
    auto leaf_entry = [](uint32_t node_index, float4 lo, float4 hi) -> uint32_t {
        const uint32_t first = uint32_t(__float_as_int(lo.w)), count = uint32_t(__float_as_int(hi.w));
        return (first < (1u << 24) && count < 128u) ? (first | (count << 24)) : (0x80000000u | node_index);
    };
    auto fetch_one = [&]() {
        while (f_cur == f_end) {
            if (f_leaf == nleaf) return;
            const uint32_t e = s_leaf[f_leaf][tid];
            ++f_leaf;
            if (e & 0x80000000u) {
                const uint32_t leaf = e & 0x7FFFFFFFu;
                f_cur = uint32_t(__float_as_int(__ldg(p.nodes + 2 * leaf).w));
                f_end = f_cur + uint32_t(__float_as_int(__ldg(p.nodes + 2 * leaf + 1).w));
            } else {
                f_cur = e & 0x00FFFFFFu;
                f_end = f_cur + (e >> 24);
            }
        }
        const uint32_t slot = n_fetched % TRI_RING;
        const float4 * src = p.tris + 3 * size_t(f_cur);
        cp_async16(&s_tri[slot][0][tid], src);
        cp_async16(&s_tri[slot][1][tid], src + 1);
        cp_async16(&s_tri[slot][2][tid], src + 2);
        cp_async_commit();
        ++f_cur;
        ++n_fetched;
    };
    // ordered stackless walk (DynamicAABBTree::query order) from `node`, buffering up to LEAF_CAP leaves
    auto ordered_walk = [&]() {
        nleaf = 0;
        while (node < p.n_nodes && nleaf < LEAF_CAP) {
            const float4 lo = __ldg(p.nodes + 2 * node);
            const float4 hi = __ldg(p.nodes + 2 * node + 1);
            const bool ov = node_overlaps(lo, hi, qb);
            const int link = __float_as_int(lo.w);
            const int cnt = __float_as_int(hi.w);
            ++c_node;
            if (cnt < 0) {
                node = ov ? node + 1 : uint32_t(link);
            } else {
                if (ov) { s_leaf[nleaf][tid] = leaf_entry(node, lo, hi); ++nleaf; }
                node = node + 1;
            }
        }
    };
    auto start_stream = [&]() {
        f_leaf = 0; f_cur = 0; f_end = 0; n_fetched = 0; n_consumed = 0;
#pragma unroll
        for (int r = 0; r < TRI_RING; ++r) fetch_one();
    };

    for (;;) {
        const unsigned m_load = __ballot_sync(FULL, state == ST_LOAD);
        const unsigned m_done = __ballot_sync(FULL, state == ST_DONE);
        if (m_done == FULL) break;
        const unsigned m_sub = __ballot_sync(FULL, state == ST_SUBSTEP);
        const unsigned m_stream = __ballot_sync(FULL, state == ST_STREAM);

        // ---------------- LOAD: new characters + their frame cache ----------------
        // Lanes without a character claim the next index from the global counter.  GATED: an index whose records are
        // still on the wire stays claimed while its lane waits in LOAD (the lanes that have work go on); the arrival
        // count is polled sparsely as long as other lanes are busy.  Chunk boundaries are multiples of 32 characters =
        // whole 128-byte lines of both record arrays, so no line with not-yet-arrived bytes was touched before.
        if (m_load && (__popc(m_load) >= load_batch || (m_sub | m_stream) == 0u)) {
            const bool others_busy = (m_sub | m_stream) != 0u;
            const unsigned m_claim = GATED ? __ballot_sync(FULL, state == ST_LOAD && i == 0xFFFFFFFFu) : m_load;
            if (m_claim) {
                unsigned base = 0;
                if (tid == 0) base = atomicAdd(p.work_counter, (unsigned)__popc(m_claim));
                base = __shfl_sync(FULL, base, 0);
                if ((m_claim >> tid) & 1u) {
                    i = base + __popc(m_claim & lt_mask);
                    if (i >= p.n) state = ST_DONE;
                }
            }
            bool take = state == ST_LOAD;
            bool run = true;
            if (GATED) {
                if (tid == 0 && (!others_busy || (++s_poll & 7u) == 0u)) s_arrived = *reinterpret_cast<const volatile uint32_t *>(p.gate);
                __syncwarp(FULL);
                take = take && i < s_arrived;
                const unsigned m_take = __ballot_sync(FULL, take);
                const unsigned m_wait = __ballot_sync(FULL, state == ST_LOAD);
                run = m_take != 0u && (__popc(m_take) >= load_batch || !others_busy || m_take == m_wait);
                if (run) __threadfence();
                else if (!others_busy) __nanosleep(300);
            }
            if (run) {
                bool loaded_now = false;
                if (take) {
                    const prtc_transform * tf = p.transforms + i;
                    const prtc_char_physics * ph = p.physics + i;
                    uint32_t cid = ph->collider;
                    if (cid >= p.n_capsules) {                    // unknown capsule id: never dereferenced; reported, result undefined
                        if (p.counters) atomicAdd(p.counters + 6, 1ull);
                        if (p.host_flags) p.host_flags[1] = 1u;
                        cid = 0;
                    }
                    pos = v3(tf->position[0], tf->position[1], tf->position[2]);
                    Quat q; q.x = tf->rotation[0]; q.y = tf->rotation[1]; q.z = tf->rotation[2]; q.w = tf->rotation[3];
                    V3 vel = v3(ph->velocity[0], ph->velocity[1], ph->velocity[2]);
                    const V3 gn = v3(ph->ground_normal[0], ph->ground_normal[1], ph->ground_normal[2]);
                    const float mvx = ph->movement[0], mvz = ph->movement[2];
                    const CapsuleDev cap = p.capsules[cid];
                    const CapsuleFrame fr = capsule_frame(cap, rotation_of(q));
                    radius = cap.radius;
                    s_park[P_QR_EPS][tid] = p.qr_eps + 1e-5f * (fabsf(cap.radius) + fabsf(cap.height) + fabsf(cap.ox) + fabsf(cap.oy) + fabsf(cap.oz));
                    // ---- phase 1 (physics_system.cpp:277-285) ----
                    if (ph->grounded) vel = vel + ((-p.stick * p.gravity) * p.dt) * gn;
                    vel.x = vel.x + mvx;
                    vel.z = vel.z + mvz;
                    grounded = 0;
                    park.put(P_VEL, vel);
                    park.put(P_STEPV, vel / float(p.substeps));
                    park.put(P_GN, gn);
                    park.put(P_SA1, fr.sa1); park.put(P_SA2, fr.sa2); park.put(P_SB1, fr.sb1); park.put(P_SB2, fr.sb2);
                    V3 a0, b0;
                    segment_ends(fr, pos, a0, b0);                // prevTform                      :344
                    park.put(P_PREVA, a0); park.put(P_PREVB, b0);
                    stepsLeft = p.substeps;
                    found = true;
                    cache_ok = false;
                    NCACHE = 0;
                    state = ST_SUBSTEP;
                    loaded_now = true;
                }
                // The frame cache (candidate leaves valid for a region around the character, kept per character slot in
                // device memory and refreshed by k_xc_refresh right before this kernel) is only read here.
                if (loaded_now) {
                    cache_ok = false;
                    if (use_cache) {
                        const size_t s = size_t(p.xc_first) + i;
                        if (p.xc_version[s] == p.tree_version) {
                            const uint32_t nc = p.xc_count[s];
                            NCACHE = nc;
                            for (uint32_t j = 0; j < nc; ++j) s_cache[j][tid] = p.xc_leaf[size_t(j) * p.xc_stride + s];
#pragma unroll
                            for (int f = 0; f < 6; ++f) s_region[f][tid] = p.xc_region[size_t(f) * p.xc_stride + s];
                            cache_ok = true;
                        }
                    }
                }
                continue;
            }
        }

        // ---------------- SUBSTEP: finish the frame or start the next substep ----------------
        if (m_sub && (__popc(m_sub) >= substep_batch || m_stream == 0u)) {
            if (state == ST_SUBSTEP) {
                if (!found || stepsLeft == 0) {
                    // `if (meshColIDs.empty()) break;` (:391-393) / loop end; remaining motion (:437); phase 3 (:298-317)
                    prtc_transform * tf = p.transforms + i;
                    prtc_char_physics * ph = p.physics + i;
                    pos = pos + (float(stepsLeft) * park.get(P_VEL)) / float(p.substeps);
                    V3 vel = v3(ph->velocity[0], ph->velocity[1], ph->velocity[2]);   // prevVelocities[i]: still untouched in memory (:277,304)
                    const float frictionRatio = fdiv(1.0f, 1.0f + (p.dt * p.friction));
                    vel.x = vel.x * frictionRatio;
                    vel.z = vel.z * frictionRatio;
                    if (grounded) vel.y = gmax((0.0f * p.gravity) * p.dt, vel.y);
                    else vel.y = vel.y + ((-1.0f * p.gravity) * p.dt);
                    const V3 gn = park.get(P_GN);
                    tf->position[0] = pos.x; tf->position[1] = pos.y; tf->position[2] = pos.z;
                    ph->velocity[0] = vel.x; ph->velocity[1] = vel.y; ph->velocity[2] = vel.z;
                    ph->ground_normal[0] = gn.x; ph->ground_normal[1] = gn.y; ph->ground_normal[2] = gn.z;
                    ph->grounded = grounded;
                    if (GATED) {
                        __threadfence();
                        const uint32_t sb = i >> INGEST_SUB_SHIFT;
                        const uint32_t left = p.n - (sb << INGEST_SUB_SHIFT);
                        const uint32_t sb_size = left < (1u << INGEST_SUB_SHIFT) ? left : (1u << INGEST_SUB_SHIFT);
                        if (atomicAdd(p.done_sub + sb, 1u) + 1u == sb_size) {
                            __threadfence_system();
                            atomicAdd(p.done_chunk + (sb >> INGEST_CHUNK_SHIFT), 1u);
                        }
                        i = 0xFFFFFFFFu;                          // no index claimed
                    }
                    state = ST_LOAD;
                } else {
                    --stepsLeft;
                    pos = pos + park.get(P_STEPV);                //                                 :350
                    const CapsuleFrame fr = park.frame(radius);
                    ps = make_pose(fr, pos);
                    qb = box_union(capsule_box(park.get(P_PREVA), park.get(P_PREVB), radius), capsule_box(ps.A, ps.B, radius));   // :355-356
                    park.put(P_PREVA, ps.A); park.put(P_PREVB, ps.B);     // prevTform = tform       :358
                    ++C_SUB;
                    found = false;
                    nleaf = 0;
                    if (cache_ok &&
                        s_region[0][tid] <= qb.lo.x && s_region[1][tid] <= qb.lo.y && s_region[2][tid] <= qb.lo.z &&
                        s_region[3][tid] >= qb.hi.x && s_region[4][tid] >= qb.hi.y && s_region[5][tid] >= qb.hi.z) {
                        bool fits = true;
                        const uint32_t ncache = NCACHE;
                        // two leaves per trip: four independent loads in flight instead of two behind every compare
                        for (uint32_t j = 0; j < ncache && fits; j += 2) {
                            const bool two = j + 1 < ncache;
                            const uint32_t leaf0 = s_cache[j][tid];
                            const uint32_t leaf1 = two ? s_cache[j + 1][tid] : leaf0;
                            const float4 lo0 = __ldg(p.nodes + 2 * leaf0), hi0 = __ldg(p.nodes + 2 * leaf0 + 1);
                            const float4 lo1 = __ldg(p.nodes + 2 * leaf1), hi1 = __ldg(p.nodes + 2 * leaf1 + 1);
                            ++c_node;
                            if (node_overlaps(lo0, hi0, qb)) {
                                if (nleaf == LEAF_CAP) { fits = false; break; }
                                s_leaf[nleaf][tid] = leaf_entry(leaf0, lo0, hi0);
                                ++nleaf;
                            }
                            if (two) {
                                ++c_node;
                                if (node_overlaps(lo1, hi1, qb)) {
                                    if (nleaf == LEAF_CAP) { fits = false; break; }
                                    s_leaf[nleaf][tid] = leaf_entry(leaf1, lo1, hi1);
                                    ++nleaf;
                                }
                            }
                        }
                        node = p.n_nodes;
                        if (!fits) { node = 0; ordered_walk(); }  // more candidates than one pass holds: walk in passes
                    } else {
                        node = 0;
                        ordered_walk();
                    }
                    if (nleaf) found = true;
                    start_stream();
                    state = ST_STREAM;
                }
            }
            continue;
        }

        // ---------------- STREAM: filter until a triangle needs the exact test ----------------
        if (state == ST_STREAM && !pending) {
            const V3 smin = gmin(ps.A, ps.B), smax = gmax(ps.A, ps.B);
            const float qr_eps = s_park[P_QR_EPS][tid];
            const float dmax = (p.abs_mode == 1u) ? radius : floorf(radius) + 1.0f;   // largest plane distance surviving the range cull
            for (;;) {
                while (n_consumed != n_fetched) {
                    if (n_fetched - n_consumed >= 2) cp_async_wait<1>(); else cp_async_wait<0>();
                    const uint32_t slot = n_consumed % TRI_RING;
                    t0 = s_tri[slot][0][tid];
                    t1 = s_tri[slot][1][tid];
                    t2 = s_tri[slot][2][tid];
                    ++n_consumed;
                    fetch_one();
                    ++c_tri;
                    if (__float_as_uint(t0.w) == PRT_DEGENERATE_BITS) continue;       // length(n) == 0   :53
                    if (use_qr && quick_reject(smin, smax, radius, dmax, qr_eps, t0, t1, t2)) continue;
                    pending = true;
                    break;
                }
                if (pending) break;
                if (node < p.n_nodes) {                           // more leaves than the buffer holds: next pass
                    ordered_walk();
                    if (nleaf) found = true;
                    start_stream();
                    if (n_fetched == 0) { state = ST_SUBSTEP; break; }
                } else {
                    state = ST_SUBSTEP;                           // this substep's candidates are exhausted
                    break;
                }
            }
        }

        // ---------------- exact tests, converged ----------------
        if (__any_sync(FULL, pending)) {
            if (pending) {
                pending = false;
                Contact ct;
                ++C_EXACT;
                if (capsule_triangle(ps, radius, p.abs_mode, v3(t0.x, t0.y, t0.z), v3(t1.x, t1.y, t1.z),
                                     v3(t2.x, t2.y, t2.z), v3(t0.w, t1.w, t2.w), ct)) {
                    pos = pos + ct.impulse;                       // handleCollision (collision_system.cpp:242-252)
                    ++C_CONTACT;
                    if (dot(ct.normal, v3(0.0f, 1.0f, 0.0f)) > p.ground_dot) {
                        grounded = 1;
                        park.put(P_GN, ct.normal);
                    }
                    if (ct.impulse.x != 0.0f || ct.impulse.y != 0.0f || ct.impulse.z != 0.0f)
                        ps = make_pose(park.frame(radius), pos);  // tform is rebuilt per triangle     :40
                }
            }
            __syncwarp(FULL);
        }
    }

    const uint32_t c_sub = __reduce_add_sync(FULL, C_SUB);
    c_tri = __reduce_add_sync(FULL, c_tri);
    c_node = __reduce_add_sync(FULL, c_node);
    const uint32_t c_contact = __reduce_add_sync(FULL, C_CONTACT);
    const uint32_t c_exact = __reduce_add_sync(FULL, C_EXACT);
    if (tid == 0 && p.counters) {
        atomicAdd(p.counters + 0, (unsigned long long)c_sub);
        atomicAdd(p.counters + 1, (unsigned long long)c_tri);
        atomicAdd(p.counters + 2, (unsigned long long)c_node);
        atomicAdd(p.counters + 3, (unsigned long long)c_contact);
        atomicAdd(p.counters + 5, (unsigned long long)c_exact);
    }
}

#undef C_SUB
#undef C_CONTACT
#undef C_EXACT
#undef NCACHE

} // namespace

void launch_k_xc_refresh(StepParams const & p, cudaStream_t stream) {
    k_xc_refresh<<<(p.n + XC_BLOCK - 1) / XC_BLOCK, XC_BLOCK, 0, stream>>>(p, p.xc_reuse ? 1 : 0);
}

void launch_k_step_stream(StepParams const & p, uint32_t blocks, bool gated, cudaStream_t stream) {
    cudaMemsetAsync(p.work_counter, 0, sizeof(unsigned), stream);
    if (gated) k_step_stream<true><<<blocks, 32, 0, stream>>>(p);
    else k_step_stream<false><<<blocks, 32, 0, stream>>>(p);
}

} // namespace prt
