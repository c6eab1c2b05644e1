// kernels.cu -- sm_100a kernels of the character-collision path.
//
//   k_step             K1+K2+K3+K4+K6 fused: one thread per character runs the whole frame of
//                      PhysicsSystem::updateCharacters for it (reference physics_system.cpp:245-318,330-438):
//                      velocity pre-step, `substeps` x { advance, swept query box, stackless BVH traversal in the
//                      reference's candidate order, Gauss-Seidel capsule-triangle push-out }, velocity finalize.
//   k_transform_refit  K8: world-space vertex cache and per-mesh AABB (physics_system.cpp:68-93,161-202)
//   k_tri_setup        leaf-ordered triangle records with precomputed unit normals
//   k_query_box        one box against the flattened tree (introspection)
//
// Compile with -fmad=false: results must be bit-identical to the CPU reference (see prt_math.cuh).
#include "kernels.cuh"

namespace prt {

namespace {

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
constexpr int DYN_CAP = 16;      // capsule neighbours buffered per lane and pass of the JACOBI pass
constexpr int XC_CAP = 12;       // leaves per cached region slot (k_xc_refresh / k_step_stream)
constexpr int TRI_RING = 2;      // triangles in flight per lane (cp.async ring in shared memory)
constexpr int LEAF_CAP = 8;      // candidate leaves buffered per lane and pass (more => another pass)

// One thread per character; every loop is WARP-UNIFORM (trip counts come from votes), bodies are predicated.
//   phase A  each lane walks the flattened tree for its query box and buffers the overlapping leaves, in the
//            reference's emission order, in shared memory;
//   phase B  each lane runs a cheap filter loop over its buffered triangles (load, degenerate skip, conservative
//            pre-cull) until one needs the exact test; the warp votes and all pending exact tests run together.
// History (profiles/): v1 let each lane run its own loops -> 2.55 of 32 lanes active per instruction; v2 made the
// loops vote-driven (13.5 lanes) but interleaved traversal and testing leaf by leaf.
template <bool TRACE, int MODE>
__global__ void __launch_bounds__(STEP_BLOCK, 512 / STEP_BLOCK) k_step(StepParams p) {
    constexpr bool LITERAL = MODE == MODE_LITERAL;
    constexpr bool JACOBI = MODE == MODE_JACOBI;
    constexpr unsigned FULL = 0xffffffffu;
    __shared__ uint2 s_leaf[LEAF_CAP][STEP_BLOCK];
    __shared__ uint32_t s_cache[XC_CAP][STEP_BLOCK];       // frame cache: leaves (node indices) overlapping s_region
    __shared__ float s_region[6][STEP_BLOCK];
    // per-lane ring of candidate triangles streamed in with cp.async while the exact tests run
    __shared__ float4 s_tri[TRI_RING][3][STEP_BLOCK];
    __shared__ uint32_t s_dyn[JACOBI ? DYN_CAP : 1][STEP_BLOCK];      // JACOBI: neighbour capsules, ascending
    const uint32_t tid = threadIdx.x;
    // small batches: only the first `lanes_per_warp` lanes of every warp carry a character, so that the batch is
    // spread over all resident warps and latency is hidden across warps instead of paid inside one
    const uint32_t lpw = p.lanes_per_warp ? p.lanes_per_warp : 32u;
    const uint32_t i = (blockIdx.x * (blockDim.x / 32u) + tid / 32u) * lpw + (tid & 31u);
    const bool valid = (tid & 31u) < lpw && i < p.n;
    uint32_t c_sub = 0, c_tri = 0, c_node = 0, c_contact = 0;

    prtc_transform * tf = p.transforms + (valid ? i : 0);
    prtc_char_physics * ph = p.physics + (valid ? i : 0);
    V3 pos = v3(0.0f), vel = v3(0.0f), gn = v3(0.0f), prevVel = v3(0.0f);
    uint32_t grounded = 0;
    CapsuleFrame fr;
    fr.sa1 = fr.sa2 = fr.sb1 = fr.sb2 = v3(0.0f);
    fr.radius = 0.0f;
    float radius = 0.0f, qr_eps = 0.0f;
    uint32_t my_capsule = 0xFFFFFFFFu, c_cc = 0, c_exact = 0;
    if (valid) {
        pos = v3(tf->position[0], tf->position[1], tf->position[2]);
        Quat q; q.x = tf->rotation[0]; q.y = tf->rotation[1]; q.z = tf->rotation[2]; q.w = tf->rotation[3];
        vel = v3(ph->velocity[0], ph->velocity[1], ph->velocity[2]);
        gn = v3(ph->ground_normal[0], ph->ground_normal[1], ph->ground_normal[2]);
        grounded = ph->grounded;
        const float mvx = ph->movement[0], mvz = ph->movement[2];
        uint32_t cid = ph->collider;
        if (cid >= p.n_capsules) {                                // unknown capsule id: never dereferenced; reported, result undefined
            if (p.counters) atomicAdd(p.counters + 6, 1ull);
            cid = 0;
        }
        const CapsuleDev cap = p.capsules[cid];
        my_capsule = cid;
        fr = capsule_frame(cap, rotation_of(q));
        radius = cap.radius;
        qr_eps = p.qr_eps + 1e-5f * (fabsf(cap.radius) + fabsf(cap.height) + fabsf(cap.ox) + fabsf(cap.oy) + fabsf(cap.oz));
        // ---- phase 1 (physics_system.cpp:277-285) ----
        prevVel = vel;
        if (grounded) vel = vel + ((-p.stick * p.gravity) * p.dt) * gn;
        vel.x = vel.x + mvx;
        vel.z = vel.z + mvz;
        grounded = 0;
    }
    // largest plane distance that survives the range cull (collision_system.cpp:72, see prtc_abs_mode)
    const float dmax = (p.abs_mode == 1u) ? radius : floorf(radius) + 1.0f;
    const bool use_qr = p.qr_enable != 0u;

    // ---- phase 2: collideCharacterWithWorld (physics_system.cpp:330-438) ----
    const float fsub = float(p.substeps);
    const V3 stepv = vel / fsub;
    V3 prevA, prevB;
    segment_ends(fr, pos, prevA, prevB);                          // prevTform                      :344
    uint32_t stepsLeft = valid ? p.substeps : 0u;
    bool alive = stepsLeft != 0;                                  // this lane is still inside the substep loop

    // ---- frame cache.  The reference walks the tree once per substep with a box that moves by velocity/substeps
    // each time.  One walk with a box R that covers the whole frame's motion (plus slack for push-outs) finds a
    // superset of every substep's candidates, in emission order; a substep whose query box lies inside R then only
    // re-tests those few leaf boxes (identical candidate list, identical order), and one whose box escapes R -- or a
    // lane whose cache overflowed -- simply walks the tree as before.
    bool cache_ok = false;
    uint32_t ncache = 0;
    if (!LITERAL && !TRACE && !(p.tune & 1u) && p.xc_version) {
        // per-slot lists kept valid by k_xc_refresh (see there); nothing to walk here
        if (alive) {
            const size_t s = size_t(p.xc_first) + i;
            if (p.xc_version[s] == p.tree_version) {
                ncache = p.xc_count[s];
                for (uint32_t j = 0; j < ncache; ++j) s_cache[j][tid] = p.xc_leaf[size_t(j) * p.xc_stride + s];
#pragma unroll
                for (int f = 0; f < 6; ++f) s_region[f][tid] = p.xc_region[size_t(f) * p.xc_stride + s];
                cache_ok = true;
            }
        }
    } else if (!LITERAL && !(p.tune & 1u)) {
        Box R;
        R.lo = v3(0.0f); R.hi = v3(0.0f);
        uint32_t node = p.n_nodes;
        if (alive) {
            V3 eA, eB;
            segment_ends(fr, pos + vel, eA, eB);
            R = box_union(capsule_box(prevA, prevB, radius), capsule_box(eA, eB, radius));
            const float slack = 0.25f;
            R.lo = R.lo - v3(slack); R.hi = R.hi + v3(slack);
            node = 0;
        }
        while (__any_sync(FULL, node < p.n_nodes)) {
            if (node < p.n_nodes) {
                const float4 lo = __ldg(p.nodes + 2 * node);
                const float4 hi = __ldg(p.nodes + 2 * node + 1);
                const bool ov = node_overlaps(lo, hi, R);
                const int cnt = __float_as_int(hi.w);
                ++c_node;
                if (cnt < 0) {
                    node = ov ? node + 1 : uint32_t(__float_as_int(lo.w));
                } else {
                    if (ov) {
                        if (ncache < LEAF_CAP) { s_cache[ncache][tid] = node; ++ncache; node = node + 1; }
                        else { node = 0xFFFFFFFFu; }              // too many leaves: this lane does not use the cache
                    } else node = node + 1;
                }
            }
        }
        cache_ok = alive && node != 0xFFFFFFFFu;
        s_region[0][tid] = R.lo.x; s_region[1][tid] = R.lo.y; s_region[2][tid] = R.lo.z;
        s_region[3][tid] = R.hi.x; s_region[4][tid] = R.hi.y; s_region[5][tid] = R.hi.z;
    }
    while (__any_sync(FULL, alive)) {
        Pose ps;
        Box qb;
        uint32_t slot = 0, t_mesh = 0, t_hits = 0, t_tris = 0;
        if (alive) {
            --stepsLeft;
            pos = pos + stepv;                                    //                                 :350
            ps = make_pose(fr, pos);
            qb = box_union(capsule_box(prevA, prevB, radius), capsule_box(ps.A, ps.B, radius));   // :355-356
            prevA = ps.A; prevB = ps.B;                           // prevTform = tform               :358
            ++c_sub;
            if (TRACE) {
                slot = i * p.substeps + (p.substeps - 1 - stepsLeft);
                TraceQuery * tq = p.trace.queries + slot;
                tq->aabb[0] = qb.lo.x; tq->aabb[1] = qb.lo.y; tq->aabb[2] = qb.lo.z;
                tq->aabb[3] = qb.hi.x; tq->aabb[4] = qb.hi.y; tq->aabb[5] = qb.hi.z;
            }
        } else {
            ps = make_pose(fr, pos);
            qb.lo = v3(0.0f); qb.hi = v3(0.0f);
        }

        if (JACOBI) {
            // ---- capsule-capsule pass, CANONICAL order: neighbours = every other capsule whose
            // STORED fat box (hysteresis of aabb_tree.cpp:102-111,140-141) overlaps the query box, visited in
            // ascending global index, each seen at its start-of-frame pose (the snapshot all ranks share).  The
            // stored boxes are binned in a uniform xz grid; a pair is reported from the one cell that holds the
            // lower corner of the two boxes' common cell range, so nothing is reported twice.
            const uint32_t self = p.dyn_first + i;
            int cx0 = 0, cx1 = -1, cz0 = 0, cz1 = -1;
            if (alive) {
                cx0 = dyn_cell(qb.lo.x, p.dyn_origin_x, p.dyn_inv_cell, p.dyn_gx); cx1 = dyn_cell(qb.hi.x, p.dyn_origin_x, p.dyn_inv_cell, p.dyn_gx);
                cz0 = dyn_cell(qb.lo.z, p.dyn_origin_z, p.dyn_inv_cell, p.dyn_gz); cz1 = dyn_cell(qb.hi.z, p.dyn_origin_z, p.dyn_inv_cell, p.dyn_gz);
            }
            uint32_t t_caps = 0;
            long long last = -1;                                  // largest neighbour index already processed
            while (true) {
                uint32_t ncand = 0;
                bool more = false;
                for (int cz = cz0; cz <= cz1; ++cz) {
                    for (int cx = cx0; cx <= cx1; ++cx) {
                        const uint32_t cell = uint32_t(cz) * p.dyn_gx + uint32_t(cx);
                        const uint32_t beg = __ldg(p.dyn_cell_start + cell), end = __ldg(p.dyn_cell_start + cell + 1);
                        for (uint32_t k = beg; k < end; ++k) {
                            const uint32_t j = __ldg(p.dyn_items + k);
                            if (j == self || (long long)j <= last) continue;
                            const float4 blo = __ldg(p.dyn_records + 4 * size_t(j) + 2);
                            const float4 bhi = __ldg(p.dyn_records + 4 * size_t(j) + 3);
                            if (!node_overlaps(blo, bhi, qb)) continue;
                            const int bx0 = dyn_cell(blo.x, p.dyn_origin_x, p.dyn_inv_cell, p.dyn_gx);
                            const int bz0 = dyn_cell(blo.z, p.dyn_origin_z, p.dyn_inv_cell, p.dyn_gz);
                            if (max(bx0, cx0) != cx || max(bz0, cz0) != cz) continue;      // reported from another cell
                            // keep the DYN_CAP smallest indices, sorted
                            if (ncand == DYN_CAP) {
                                more = true;
                                if (j > s_dyn[DYN_CAP - 1][tid]) continue;
                                --ncand;
                            }
                            uint32_t q = ncand;
                            while (q > 0 && s_dyn[q - 1][tid] > j) { s_dyn[q][tid] = s_dyn[q - 1][tid]; --q; }
                            s_dyn[q][tid] = j;
                            ++ncand;
                        }
                    }
                }
                __syncwarp(FULL);
                const uint32_t rounds = __reduce_max_sync(FULL, ncand);
                if (rounds == 0) break;
                for (uint32_t r = 0; r < rounds; ++r) {
                    if (r < ncand) {
                        const uint32_t j = s_dyn[r][tid];
                        if (TRACE) {
                            if (t_caps < p.trace.cap_cand) p.trace.caps_ids[size_t(slot) * p.trace.cap_cand + t_caps] = j;
                            ++t_caps;
                        }
                        const float4 sA = __ldg(p.dyn_records + 4 * size_t(j)), sB = __ldg(p.dyn_records + 4 * size_t(j) + 1);
                        V3 aA, aB;
                        segment_ends(fr, pos, aA, aB);
                        Contact ct;
                        ++c_cc;
                        if (capsule_capsule(aA, aB, radius, v3(sA.x, sA.y, sA.z), v3(sB.x, sB.y, sB.z), sA.w, ct)) {
                            pos = pos + ct.impulse;               // handleCollision (collision_system.cpp:242-252)
                            ++c_contact;
                            if (dot(ct.normal, v3(0.0f, 1.0f, 0.0f)) > p.ground_dot) {
                                grounded = 1;
                                gn = ct.normal;
                            }
                            if (TRACE) {
                                if (t_hits < p.trace.cap_hits) {
                                    TraceHit * h = p.trace.hits + size_t(slot) * p.trace.cap_hits + t_hits;
                                    h->poly = -1;
                                    h->impulse[0] = ct.impulse.x; h->impulse[1] = ct.impulse.y; h->impulse[2] = ct.impulse.z;
                                    h->normal[0] = ct.normal.x; h->normal[1] = ct.normal.y; h->normal[2] = ct.normal.z;
                                    h->pos[0] = pos.x; h->pos[1] = pos.y; h->pos[2] = pos.z;
                                }
                                ++t_hits;
                            }
                        }
                    }
                }
                if (ncand) last = (long long)s_dyn[ncand - 1][tid];
                if (!more) { cx1 = cx0 - 1; cz1 = cz0 - 1; }       // this lane is finished: empty cell range
            }
            if (TRACE && alive) p.trace.queries[slot].n_caps = t_caps;
            ps = make_pose(fr, pos);
        }
        if (LITERAL) {
            // ---- capsule-capsule pass (physics_system.cpp:369-389): every capsule leaf of the unified tree that the
            // query box overlaps, in emission order, BEFORE any triangle.  Character j is seen at its end-of-frame
            // pose of the previous fixpoint round when j < i and at its start-of-frame pose when j > i -- at the
            // fixpoint that is exactly the reference's sequential visibility (T18).
            uint32_t cnode = alive ? 0u : p.n_nodes;
            uint32_t t_caps = 0;
            while (true) {
                uint32_t other = 0xFFFFFFFFu;
                while (cnode < p.n_nodes) {
                    const float4 lo = __ldg(p.nodes + 2 * cnode);
                    const float4 hi = __ldg(p.nodes + 2 * cnode + 1);
                    const bool ov = node_overlaps(lo, hi, qb);
                    const int link = __float_as_int(lo.w);
                    const int cnt = __float_as_int(hi.w);
                    if (cnt < 0) {
                        cnode = ov ? cnode + 1 : uint32_t(link);
                    } else {
                        cnode = cnode + 1;
                        if (ov && (cnt & PRT_CAPSULE_LEAF) && uint32_t(link) != my_capsule) { other = uint32_t(link); break; }
                    }
                }
                if (!__any_sync(FULL, other != 0xFFFFFFFFu)) break;
                if (other != 0xFFFFFFFFu) {
                    if (TRACE) {
                        if (t_caps < p.trace.cap_cand) p.trace.caps_ids[size_t(slot) * p.trace.cap_cand + t_caps] = other;
                        ++t_caps;
                    }
                    const uint32_t j = p.cap2char[other];
                    const float4 * seg = (j < i ? p.seg_prev : p.seg_start) + 2 * size_t(j);
                    const float4 sA = __ldg(seg), sB = __ldg(seg + 1);
                    V3 aA, aB;
                    segment_ends(fr, pos, aA, aB);
                    Contact ct;
                    ++c_cc;
                    if (capsule_capsule(aA, aB, radius, v3(sA.x, sA.y, sA.z), v3(sB.x, sB.y, sB.z), sA.w, ct)) {
                        pos = pos + ct.impulse;                   // handleCollision (collision_system.cpp:242-252)
                        ++c_contact;
                        if (dot(ct.normal, v3(0.0f, 1.0f, 0.0f)) > p.ground_dot) {
                            grounded = 1;
                            gn = ct.normal;
                        }
                        if (TRACE) {
                            if (t_hits < p.trace.cap_hits) {
                                TraceHit * h = p.trace.hits + size_t(slot) * p.trace.cap_hits + t_hits;
                                h->poly = -1;
                                h->impulse[0] = ct.impulse.x; h->impulse[1] = ct.impulse.y; h->impulse[2] = ct.impulse.z;
                                h->normal[0] = ct.normal.x; h->normal[1] = ct.normal.y; h->normal[2] = ct.normal.z;
                                h->pos[0] = pos.x; h->pos[1] = pos.y; h->pos[2] = pos.z;
                            }
                            ++t_hits;
                        }
                    }
                }
            }
            if (TRACE && alive) p.trace.queries[slot].n_caps = t_caps;
            ps = make_pose(fr, pos);
            if (alive && p.last_box) {
                float * lb = p.last_box + 6 * size_t(i);
                lb[0] = qb.lo.x; lb[1] = qb.lo.y; lb[2] = qb.lo.z; lb[3] = qb.hi.x; lb[4] = qb.hi.y; lb[5] = qb.hi.z;
            }
        }
        bool found = false;
        uint32_t node = alive ? 0u : p.n_nodes;
        uint32_t poly = 0;
        uint32_t nleaf = 0;
        if (alive && cache_ok &&
            s_region[0][tid] <= qb.lo.x && s_region[1][tid] <= qb.lo.y && s_region[2][tid] <= qb.lo.z &&
            s_region[3][tid] >= qb.hi.x && s_region[4][tid] >= qb.hi.y && s_region[5][tid] >= qb.hi.z) {
            // query box inside the frame region: its candidates are the cached leaves that overlap it
            for (uint32_t j = 0; j < ncache; ++j) {
                const uint32_t leaf = s_cache[j][tid];
                const float4 lo = __ldg(p.nodes + 2 * leaf);
                const float4 hi = __ldg(p.nodes + 2 * leaf + 1);
                ++c_node;
                if (node_overlaps(lo, hi, qb)) {
                    if (nleaf == LEAF_CAP) { nleaf = 0xFFFFFFFFu; break; }    // more than one pass holds: ordered walk below
                    if (TRACE) {
                        if (t_mesh < p.trace.cap_cand) p.trace.mesh_ids[size_t(slot) * p.trace.cap_cand + t_mesh] = p.node_leaf_id[leaf];
                        ++t_mesh;
                    }
                    s_leaf[nleaf][tid] = make_uint2(uint32_t(__float_as_int(lo.w)), uint32_t(__float_as_int(hi.w)));
                    ++nleaf;
                }
            }
            node = p.n_nodes;
            if (nleaf == 0xFFFFFFFFu) { nleaf = 0; node = 0; if (TRACE) t_mesh = 0; }
        }
        while (true) {
            // ---- phase A: stackless traversal in right-first preorder == DynamicAABBTree::query's emission order
            // (aabb_tree.cpp:34-68).  The query box is fixed for the substep, so buffering candidates and testing
            // them afterwards is exactly the reference's "query, then collideCapsuleMesh" (physics_system.cpp:367,431).
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
                    if (ov && !(LITERAL && (cnt & PRT_CAPSULE_LEAF))) {
                        if (TRACE) {
                            if (t_mesh < p.trace.cap_cand) p.trace.mesh_ids[size_t(slot) * p.trace.cap_cand + t_mesh] = p.node_leaf_id[node];
                            ++t_mesh;
                        }
                        s_leaf[nleaf][tid] = make_uint2(uint32_t(link), uint32_t(cnt));
                        ++nleaf;
                    }
                    node = node + 1;
                }
            }
            if (!__any_sync(FULL, nleaf != 0)) break;             // every lane has exhausted its traversal
            if (nleaf) found = true;

            // ---- phase B: Gauss-Seidel push-out over the buffered triangles, in order (collision_system.cpp:29-155)
            // The lane's candidate triangles form one stream (leaf after leaf, in order).  A fetch cursor runs
            // TRI_RING triangles ahead of the consume position, so the 48-byte loads overlap the exact tests.
            uint32_t f_leaf = 0, f_cur = 0, f_end = 0, n_fetched = 0, n_consumed = 0;
            auto fetch_one = [&]() {
                while (f_cur == f_end) {
                    if (f_leaf == nleaf) return;
                    const uint2 e = s_leaf[f_leaf][tid];
                    ++f_leaf;
                    f_cur = e.x; f_end = e.x + e.y;
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
#pragma unroll
            for (int r = 0; r < TRI_RING; ++r) fetch_one();
            while (true) {
                bool pending = false;
                float4 t0, t1, t2;
                uint32_t my_poly = 0;
                const V3 smin = gmin(ps.A, ps.B), smax = gmax(ps.A, ps.B);
                while (n_consumed != n_fetched) {                 // filter loop: cheap, per lane
                    {   // the oldest outstanding group is the triangle to consume: allow (outstanding - 1) younger ones
                        const uint32_t outstanding = n_fetched - n_consumed;
                        if (TRI_RING >= 4 && outstanding >= 4) cp_async_wait<3>();
                        else if (TRI_RING >= 3 && outstanding == 3) cp_async_wait<2>();
                        else if (outstanding == 2) cp_async_wait<1>();
                        else cp_async_wait<0>();
                    }
                    const uint32_t slot = n_consumed % TRI_RING;
                    t0 = s_tri[slot][0][tid];
                    t1 = s_tri[slot][1][tid];
                    t2 = s_tri[slot][2][tid];
                    ++n_consumed;
                    fetch_one();                                  // refill the slot just read
                    ++c_tri;
                    if (TRACE) ++t_tris;
                    my_poly = poly++;
                    if (__float_as_uint(t0.w) == PRT_DEGENERATE_BITS) continue;       // length(n) == 0   :53
                    if (use_qr && quick_reject(smin, smax, radius, dmax, qr_eps, t0, t1, t2)) continue;
                    pending = true;
                    break;
                }
                if (!__any_sync(FULL, pending)) break;
                if (pending) {
                    Contact ct;
                    ++c_exact;
                    if (capsule_triangle(ps, radius, p.abs_mode, v3(t0.x, t0.y, t0.z), v3(t1.x, t1.y, t1.z),
                                         v3(t2.x, t2.y, t2.z), v3(t0.w, t1.w, t2.w), ct)) {
                        // handleCollision (collision_system.cpp:242-252)
                        pos = pos + ct.impulse;
                        ++c_contact;
                        if (dot(ct.normal, v3(0.0f, 1.0f, 0.0f)) > p.ground_dot) {
                            grounded = 1;
                            gn = ct.normal;
                        }
                        if (TRACE) {
                            if (t_hits < p.trace.cap_hits) {
                                TraceHit * h = p.trace.hits + size_t(slot) * p.trace.cap_hits + t_hits;
                                h->poly = int32_t(my_poly);
                                h->impulse[0] = ct.impulse.x; h->impulse[1] = ct.impulse.y; h->impulse[2] = ct.impulse.z;
                                h->normal[0] = ct.normal.x; h->normal[1] = ct.normal.y; h->normal[2] = ct.normal.z;
                                h->pos[0] = pos.x; h->pos[1] = pos.y; h->pos[2] = pos.z;
                            }
                            ++t_hits;
                        }
                        if (ct.impulse.x != 0.0f || ct.impulse.y != 0.0f || ct.impulse.z != 0.0f)
                            ps = make_pose(fr, pos);              // tform is rebuilt per triangle     :40
                    }
                }
                __syncwarp(FULL);                                 // keep the lanes in step round by round
            }
            nleaf = 0;
        }
        if (alive) {
            if (TRACE) {
                TraceQuery * tq = p.trace.queries + slot;
                tq->executed = 1; tq->n_mesh = t_mesh; if (!LITERAL && !JACOBI) tq->n_caps = 0; tq->n_hits = t_hits; tq->tri_tests = t_tris;
            }
            if (!found || stepsLeft == 0) alive = false;          // `if (meshColIDs.empty()) break;`  :391-393
        }
    }

    if (valid) {
        pos = pos + (float(stepsLeft) * vel) / fsub;              //                                   :437

        // ---- phase 3 (physics_system.cpp:298-317) ----
        vel = prevVel;
        const float frictionRatio = fdiv(1.0f, 1.0f + (p.dt * p.friction));
        vel.x = vel.x * frictionRatio;
        vel.z = vel.z * frictionRatio;
        if (grounded) vel.y = gmax((0.0f * p.gravity) * p.dt, vel.y);
        else vel.y = vel.y + ((-1.0f * p.gravity) * p.dt);

        tf->position[0] = pos.x; tf->position[1] = pos.y; tf->position[2] = pos.z;
        ph->velocity[0] = vel.x; ph->velocity[1] = vel.y; ph->velocity[2] = vel.z;
        ph->ground_normal[0] = gn.x; ph->ground_normal[1] = gn.y; ph->ground_normal[2] = gn.z;
        ph->grounded = grounded;
    }

    // work counters: one atomic per warp per counter
    c_sub = __reduce_add_sync(FULL, c_sub);
    c_tri = __reduce_add_sync(FULL, c_tri);
    c_node = __reduce_add_sync(FULL, c_node);
    c_contact = __reduce_add_sync(FULL, c_contact);
    c_exact = __reduce_add_sync(FULL, c_exact);
    if (LITERAL || JACOBI) c_cc = __reduce_add_sync(FULL, c_cc);
    if ((tid & 31u) == 0 && p.counters) {
        if (LITERAL || JACOBI) atomicAdd(p.counters + 4, (unsigned long long)c_cc);
        atomicAdd(p.counters + 5, (unsigned long long)c_exact);
        atomicAdd(p.counters + 0, (unsigned long long)c_sub);
        atomicAdd(p.counters + 1, (unsigned long long)c_tri);
        atomicAdd(p.counters + 2, (unsigned long long)c_node);
        atomicAdd(p.counters + 3, (unsigned long long)c_contact);
    }
}

// One thread per mesh collider: sequential left fold, exactly `min = glm::min(min, cache[i])`
// (physics_system.cpp:80-91), so NaN vertices behave as in the reference.
__global__ void k_transform_refit(const float * __restrict__ raw, float * __restrict__ world,
                                  const MeshDev * __restrict__ meshes, uint32_t first_mesh, uint32_t n_meshes,
                                  const ModelXform * __restrict__ xforms, float * __restrict__ aabbs) {
    uint32_t m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= n_meshes) return;
    m += first_mesh;
    const MeshDev md = meshes[m];
    const Affine xf = xforms[md.model].m;
    V3 mn = v3(3.402823466e+38f);
    V3 mx = v3(-3.402823466e+38f);
    for (uint32_t k = 0; k < md.n_vertices; ++k) {
        size_t o = 3 * size_t(md.first_vertex + k);
        V3 w = affine_point(xf, v3(raw[o], raw[o + 1], raw[o + 2]));
        world[o] = w.x; world[o + 1] = w.y; world[o + 2] = w.z;
        mn = gmin(mn, w);
        mx = gmax(mx, w);
    }
    float * a = aabbs + 6 * size_t(m);
    a[0] = mn.x; a[1] = mn.y; a[2] = mn.z; a[3] = mx.x; a[4] = mx.y; a[5] = mx.z;
}

__global__ void k_tri_setup(const float * __restrict__ world, const uint32_t * __restrict__ tri_src_vertex,
                            uint32_t n_tris, float4 * __restrict__ tris) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tris) return;
    size_t o = 3 * size_t(tri_src_vertex[t]);
    V3 a = v3(world[o], world[o + 1], world[o + 2]);
    V3 b = v3(world[o + 3], world[o + 4], world[o + 5]);
    V3 c = v3(world[o + 6], world[o + 7], world[o + 8]);
    V3 n = v3(0.0f);
    bool ok = triangle_normal(a, b, c, n);
    if (!ok) n = v3(__uint_as_float(PRT_DEGENERATE_BITS), 0.0f, 0.0f);
    tris[3 * size_t(t)] = make_float4(a.x, a.y, a.z, n.x);
    tris[3 * size_t(t) + 1] = make_float4(b.x, b.y, b.z, n.y);
    tris[3 * size_t(t) + 2] = make_float4(c.x, c.y, c.z, n.z);
}

__global__ void k_query_box(const float4 * __restrict__ nodes, uint32_t n_nodes, const uint32_t * __restrict__ node_leaf_id,
                            const float * __restrict__ aabb6, uint32_t * out_ids, uint32_t capacity, uint32_t * out_count) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    Box q;
    q.lo = v3(aabb6[0], aabb6[1], aabb6[2]);
    q.hi = v3(aabb6[3], aabb6[4], aabb6[5]);
    uint32_t node = 0, found = 0;
    while (node < n_nodes) {
        const float4 lo = nodes[2 * node];
        const float4 hi = nodes[2 * node + 1];
        const bool ov = node_overlaps(lo, hi, q);
        const int cnt = __float_as_int(hi.w);
        if (cnt < 0) {
            node = ov ? node + 1 : uint32_t(__float_as_int(lo.w));
        } else {
            if (ov) {
                if (found < capacity) out_ids[found] = node_leaf_id[node];
                ++found;
            }
            ++node;
        }
    }
    *out_count = found;
}

// ------------------------------------------------------------------------------------------------------------
// k_step_warp: ONE WARP PER CHARACTER, for batches smaller than the number of resident warps (engine-sized scenes:
// C1, C2), where what matters is the latency of one character's frame, not throughput.
//   broadphase   the warp walks the flattened tree level by level: a frontier of pending nodes lives in shared
//                memory, every lane tests one node per round (coalesced-ish 32-byte loads, 32 in flight), children of
//                overlapping nodes are compacted into the next frontier with ballots, overlapping leaves are
//                collected and then sorted by node index -- a leaf's index in the right-first preorder array IS its rank
//                in DynamicAABBTree::query's emission order (aabb_tree.cpp:34-68);
//   narrowphase  the candidate triangles form one ordered stream; 32 consecutive ones are tested SPECULATIVELY against
//                the current pose, one per lane; a ballot finds the first lane (= first triangle in reference order)
//                whose contact carries a non-zero push; every contact before it is committed in lane order (they do
//                not move the capsule, collision_system.cpp:132-141), that push is applied, and testing resumes with
//                the triangle after it -- exactly the sequential Gauss-Seidel loop of collision_system.cpp:29-155.
// LITERAL order runs the capsule-capsule pass (tree order) before the triangles.  Untraced launches only; results,
// counters and candidate order are those of k_step.
// ------------------------------------------------------------------------------------------------------------
constexpr int WF_CAP = 128;          // frontier / leaf-list capacity per warp (overflow => ordered passes)

template <int MODE>
__global__ void __launch_bounds__(32) k_step_warp(StepParams p) {
    constexpr bool LITERAL = MODE == MODE_LITERAL;
    constexpr unsigned FULL = 0xffffffffu;
    __shared__ uint32_t s_front[2][WF_CAP];
    __shared__ uint32_t s_found[WF_CAP];        // overlapping mesh leaves (node indices), unsorted
    __shared__ uint32_t s_mesh[WF_CAP];         // ... sorted = emission order
    __shared__ uint32_t s_start[WF_CAP + 1];    // exclusive prefix sum of the leaves' triangle counts
    __shared__ uint32_t s_tbeg[WF_CAP];
    __shared__ uint32_t s_caps[WF_CAP];         // LITERAL: overlapping capsule leaves (node indices), sorted
    const uint32_t lane = threadIdx.x;
    const unsigned lt = (1u << lane) - 1u;
    const uint32_t i = blockIdx.x;
    if (i >= p.n) return;
    uint32_t c_sub = 0, c_tri = 0, c_node = 0, c_contact = 0, c_cc = 0;

    // every lane holds the character's state (uniform loads, broadcast)
    prtc_transform * tf = p.transforms + i;
    prtc_char_physics * ph = p.physics + i;
    V3 pos = v3(tf->position[0], tf->position[1], tf->position[2]);
    Quat q; q.x = tf->rotation[0]; q.y = tf->rotation[1]; q.z = tf->rotation[2]; q.w = tf->rotation[3];
    V3 vel = v3(ph->velocity[0], ph->velocity[1], ph->velocity[2]);
    V3 gn = v3(ph->ground_normal[0], ph->ground_normal[1], ph->ground_normal[2]);
    uint32_t grounded = ph->grounded;
    uint32_t cid = ph->collider;
    if (cid >= p.n_capsules) {
        if (lane == 0 && p.counters) atomicAdd(p.counters + 6, 1ull);
        cid = 0;
    }
    const CapsuleDev cap = p.capsules[cid];
    const CapsuleFrame fr = capsule_frame(cap, rotation_of(q));
    const float radius = cap.radius;
    // ---- phase 1 (physics_system.cpp:277-285) ----
    const V3 prevVel = vel;
    if (grounded) vel = vel + ((-p.stick * p.gravity) * p.dt) * gn;
    vel.x = vel.x + ph->movement[0];
    vel.z = vel.z + ph->movement[2];
    grounded = 0;

    // ---- phase 2: collideCharacterWithWorld (physics_system.cpp:330-438) ----
    const float fsub = float(p.substeps);
    const V3 stepv = vel / fsub;
    V3 prevA, prevB;
    segment_ends(fr, pos, prevA, prevB);
    uint32_t stepsLeft = p.substeps;
    while (stepsLeft != 0) {
        --stepsLeft;
        pos = pos + stepv;
        Pose ps = make_pose(fr, pos);
        const Box qb = box_union(capsule_box(prevA, prevB, radius), capsule_box(ps.A, ps.B, radius));
        prevA = ps.A; prevB = ps.B;
        ++c_sub;
        if (LITERAL && p.last_box && lane == 0) {
            float * lb = p.last_box + 6 * size_t(i);
            lb[0] = qb.lo.x; lb[1] = qb.lo.y; lb[2] = qb.lo.z; lb[3] = qb.hi.x; lb[4] = qb.hi.y; lb[5] = qb.hi.z;
        }

        // ---- broadphase: level-by-level frontier walk ----
        uint32_t nmesh = 0, ncaps = 0, walk_tests = 0;
        bool overflow = false;
        {
            uint32_t cur = 0, nfront = p.n_nodes ? 1u : 0u;
            if (lane == 0) s_front[0][0] = 0u;
            __syncwarp(FULL);
            while (nfront != 0 && !overflow) {
                uint32_t nnext = 0;
                for (uint32_t base = 0; base < nfront; base += 32) {
                    const uint32_t k = base + lane;
                    bool internal = false, leaf_mesh = false, leaf_caps = false;
                    uint32_t idx = 0, second = 0;
                    if (k < nfront) {
                        idx = s_front[cur][k];
                        const float4 lo = __ldg(p.nodes + 2 * idx);
                        const float4 hi = __ldg(p.nodes + 2 * idx + 1);
                        ++walk_tests;
                        if (node_overlaps(lo, hi, qb)) {
                            const int cnt = __float_as_int(hi.w);
                            if (cnt < 0) { internal = true; second = uint32_t(-cnt); }
                            else if (!(cnt & PRT_CAPSULE_LEAF)) leaf_mesh = true;
                            else if (LITERAL && uint32_t(__float_as_int(lo.w)) != cid) leaf_caps = true;
                        }
                    }
                    const unsigned m_int = __ballot_sync(FULL, internal);
                    const unsigned m_mesh = __ballot_sync(FULL, leaf_mesh);
                    const unsigned m_caps = __ballot_sync(FULL, leaf_caps);
                    if (nnext + 2 * __popc(m_int) > WF_CAP || nmesh + __popc(m_mesh) > WF_CAP || ncaps + __popc(m_caps) > WF_CAP) { overflow = true; break; }
                    if (internal) {
                        const uint32_t o = nnext + 2 * __popc(m_int & lt);
                        s_front[cur ^ 1][o] = idx + 1;
                        s_front[cur ^ 1][o + 1] = second;
                    }
                    if (leaf_mesh) s_found[nmesh + __popc(m_mesh & lt)] = idx;
                    if (leaf_caps) s_caps[ncaps + __popc(m_caps & lt)] = idx;     // sorted below (few)
                    nnext += 2 * __popc(m_int);
                    nmesh += __popc(m_mesh);
                    ncaps += __popc(m_caps);
                }
                cur ^= 1;
                nfront = nnext;
                __syncwarp(FULL);
            }
        }
        // ordered passes (overflow only): the stackless right-first walk, executed uniformly by the whole warp
        uint32_t onode = p.n_nodes, cnode = p.n_nodes;            // resume positions of the mesh / capsule passes
        if (overflow) { onode = 0; cnode = 0; nmesh = 0; ncaps = 0; }
        else {
            c_node += __reduce_add_sync(FULL, walk_tests) ;
            // sort by node index (rank by counting; the lists are short)
            __syncwarp(FULL);
            for (uint32_t j = lane; j < nmesh; j += 32) {
                const uint32_t e = s_found[j];
                uint32_t r = 0;
                for (uint32_t t = 0; t < nmesh; ++t) r += s_found[t] < e;
                s_mesh[r] = e;
            }
            if (LITERAL) {
                uint32_t mine[WF_CAP / 32], rk[WF_CAP / 32];
#pragma unroll
                for (int s = 0; s < WF_CAP / 32; ++s) {
                    const uint32_t j = lane + 32 * s;
                    mine[s] = j < ncaps ? s_caps[j] : 0xFFFFFFFFu;
                    rk[s] = 0;
                    for (uint32_t t = 0; t < ncaps; ++t) rk[s] += s_caps[t] < mine[s];
                }
                __syncwarp(FULL);
#pragma unroll
                for (int s = 0; s < WF_CAP / 32; ++s) if (lane + 32 * s < ncaps) s_caps[rk[s]] = mine[s];
            }
            __syncwarp(FULL);
        }

        // ---- LITERAL: capsule-capsule pass in tree order, before any triangle (physics_system.cpp:369-389) ----
        if (LITERAL) {
            for (;;) {
                if (overflow) {                                   // next pass of capsule leaves, ordered walk
                    ncaps = 0;
                    while (cnode < p.n_nodes && ncaps < WF_CAP) {
                        const float4 lo = __ldg(p.nodes + 2 * cnode);
                        const float4 hi = __ldg(p.nodes + 2 * cnode + 1);
                        const bool ov = node_overlaps(lo, hi, qb);
                        const int cnt = __float_as_int(hi.w);
                        if (cnt < 0) cnode = ov ? cnode + 1 : uint32_t(__float_as_int(lo.w));
                        else {
                            if (ov && (cnt & PRT_CAPSULE_LEAF) && uint32_t(__float_as_int(lo.w)) != cid) { if (lane == 0) s_caps[ncaps] = cnode; ++ncaps; }
                            cnode = cnode + 1;
                        }
                    }
                    __syncwarp(FULL);
                }
                for (uint32_t j = 0; j < ncaps; ++j) {
                    const uint32_t other = uint32_t(__float_as_int(__ldg(p.nodes + 2 * s_caps[j]).w));
                    const uint32_t oc = p.cap2char[other];
                    const float4 * seg = (oc < i ? p.seg_prev : p.seg_start) + 2 * size_t(oc);
                    const float4 sA = __ldg(seg), sB = __ldg(seg + 1);
                    V3 aA, aB;
                    segment_ends(fr, pos, aA, aB);
                    Contact ct;
                    ++c_cc;
                    if (capsule_capsule(aA, aB, radius, v3(sA.x, sA.y, sA.z), v3(sB.x, sB.y, sB.z), sA.w, ct)) {
                        pos = pos + ct.impulse;
                        ++c_contact;
                        if (dot(ct.normal, v3(0.0f, 1.0f, 0.0f)) > p.ground_dot) { grounded = 1; gn = ct.normal; }
                    }
                }
                __syncwarp(FULL);
                if (!overflow || cnode >= p.n_nodes) break;
            }
            ps = make_pose(fr, pos);
        }

        // ---- narrowphase: speculative 32-wide Gauss-Seidel over the ordered triangle stream ----
        bool found = false;
        for (;;) {
            if (overflow) {                                       // next pass of mesh leaves, ordered walk
                nmesh = 0;
                while (onode < p.n_nodes && nmesh < WF_CAP) {
                    const float4 lo = __ldg(p.nodes + 2 * onode);
                    const float4 hi = __ldg(p.nodes + 2 * onode + 1);
                    const bool ov = node_overlaps(lo, hi, qb);
                    const int cnt = __float_as_int(hi.w);
                    if (lane == 0) ++c_node;
                    if (cnt < 0) onode = ov ? onode + 1 : uint32_t(__float_as_int(lo.w));
                    else {
                        if (ov && !(cnt & PRT_CAPSULE_LEAF)) { if (lane == 0) s_mesh[nmesh] = onode; ++nmesh; }
                        onode = onode + 1;
                    }
                }
                __syncwarp(FULL);
            }
            if (nmesh) found = true;
            // triangle ranges of the leaves and their exclusive prefix sum (serial in lane 0: the lists are short)
            for (uint32_t j = lane; j < nmesh; j += 32) s_tbeg[j] = uint32_t(__float_as_int(__ldg(p.nodes + 2 * s_mesh[j]).w));
            __syncwarp(FULL);
            if (lane == 0) {
                uint32_t acc = 0;
                for (uint32_t j = 0; j < nmesh; ++j) { s_start[j] = acc; acc += uint32_t(__float_as_int(__ldg(p.nodes + 2 * s_mesh[j] + 1).w)); }
                s_start[nmesh] = acc;
            }
            __syncwarp(FULL);
            const uint32_t total = s_start[nmesh];
            uint32_t base = 0;
            while (base < total) {
                const uint32_t seq = base + lane;
                bool contact = false, ground = false, nonzero = false;
                Contact ct; ct.impulse = v3(0.0f); ct.normal = v3(0.0f);
                if (seq < total) {
                    uint32_t lo_j = 0, hi_j = nmesh;              // leaf holding triangle #seq: last j with start[j] <= seq
                    while (hi_j - lo_j > 1) { const uint32_t mid = (lo_j + hi_j) >> 1; if (s_start[mid] <= seq) lo_j = mid; else hi_j = mid; }
                    const size_t t = size_t(s_tbeg[lo_j]) + (seq - s_start[lo_j]);
                    const float4 t0 = __ldg(p.tris + 3 * t), t1 = __ldg(p.tris + 3 * t + 1), t2 = __ldg(p.tris + 3 * t + 2);
                    if (__float_as_uint(t0.w) != PRT_DEGENERATE_BITS)
                        contact = capsule_triangle(ps, radius, p.abs_mode, v3(t0.x, t0.y, t0.z), v3(t1.x, t1.y, t1.z),
                                                   v3(t2.x, t2.y, t2.z), v3(t0.w, t1.w, t2.w), ct);
                    if (contact) {
                        ground = dot(ct.normal, v3(0.0f, 1.0f, 0.0f)) > p.ground_dot;
                        nonzero = ct.impulse.x != 0.0f || ct.impulse.y != 0.0f || ct.impulse.z != 0.0f;
                    }
                }
                const unsigned m_contact = __ballot_sync(FULL, contact);
                const unsigned m_push = __ballot_sync(FULL, nonzero);
                // commit everything up to and including the first pushing triangle; later lanes saw a stale pose
                const int last = m_push ? __ffs(m_push) - 1 : 31;
                const unsigned upto = last == 31 ? FULL : ((2u << last) - 1u);
                const unsigned m_commit = m_contact & upto;
                const unsigned m_ground = __ballot_sync(FULL, ground) & upto;
                if (m_commit) {
                    if (m_commit & ~m_push) pos = pos + v3(0.0f);             // `position += impulse` with a zero impulse (:242)
                    if (m_push) {
                        const V3 imp = v3(__shfl_sync(FULL, ct.impulse.x, last), __shfl_sync(FULL, ct.impulse.y, last), __shfl_sync(FULL, ct.impulse.z, last));
                        pos = pos + imp;
                    }
                    if (m_ground) {                                            // the LAST ground contact in order wins (:247-252)
                        const int src = 31 - __clz(m_ground);
                        gn = v3(__shfl_sync(FULL, ct.normal.x, src), __shfl_sync(FULL, ct.normal.y, src), __shfl_sync(FULL, ct.normal.z, src));
                        grounded = 1;
                    }
                    c_contact += __popc(m_commit);
                }
                const uint32_t consumed = m_push ? uint32_t(last) + 1u : (total - base < 32u ? total - base : 32u);
                c_tri += consumed;
                base += consumed;
                if (m_push) ps = make_pose(fr, pos);
            }
            __syncwarp(FULL);
            if (!overflow || onode >= p.n_nodes) break;
        }
        if (!found) break;                                        // `if (meshColIDs.empty()) break;`  :391-393
    }
    pos = pos + (float(stepsLeft) * vel) / fsub;                  //                                   :437

    // ---- phase 3 (physics_system.cpp:298-317) ----
    vel = prevVel;
    const float frictionRatio = fdiv(1.0f, 1.0f + (p.dt * p.friction));
    vel.x = vel.x * frictionRatio;
    vel.z = vel.z * frictionRatio;
    if (grounded) vel.y = gmax((0.0f * p.gravity) * p.dt, vel.y);
    else vel.y = vel.y + ((-1.0f * p.gravity) * p.dt);
    if (lane == 0) {
        tf->position[0] = pos.x; tf->position[1] = pos.y; tf->position[2] = pos.z;
        ph->velocity[0] = vel.x; ph->velocity[1] = vel.y; ph->velocity[2] = vel.z;
        ph->ground_normal[0] = gn.x; ph->ground_normal[1] = gn.y; ph->ground_normal[2] = gn.z;
        ph->grounded = grounded;
        if (p.counters) {
            atomicAdd(p.counters + 0, (unsigned long long)c_sub);
            atomicAdd(p.counters + 1, (unsigned long long)c_tri);
            atomicAdd(p.counters + 2, (unsigned long long)c_node);
            atomicAdd(p.counters + 3, (unsigned long long)c_contact);
            if (LITERAL) atomicAdd(p.counters + 4, (unsigned long long)c_cc);
        }
    }
}

// k_xc_refresh: keeps the per-slot frame cache of k_step_stream valid for the coming frame.  One thread per
// character: the region its frame can need (box of the current pose united with the pose after the pre-stepped
// velocity, +0.25 for push-outs); if the slot's stored region (built for this tree version) already contains it, done --
// resting or slow characters never walk the tree again; else one stackless walk with the region grown by xc_margin
// collects the overlapping leaves (emission order) and the slot is rewritten.  The list is a pure function of
// (tree, region): it does not matter which character occupied the slot before, and k_step_stream re-checks
// "query box inside region" at every substep, so a stale or too-small region can only cost time, never correctness.
__global__ void __launch_bounds__(128) k_xc_refresh(StepParams p, int reuse) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.n) return;
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
    vel.x = vel.x + ph->movement[0];
    vel.z = vel.z + ph->movement[2];
    V3 a0, b0, eA, eB;
    segment_ends(fr, pos, a0, b0);
    segment_ends(fr, pos + vel, eA, eB);
    Box R = box_union(capsule_box(a0, b0, cap.radius), capsule_box(eA, eB, cap.radius));
    R.lo = R.lo - v3(0.25f); R.hi = R.hi + v3(0.25f);
    const size_t s = size_t(p.xc_first) + i;
    if (reuse && p.xc_version[s] == p.tree_version) {
        if (p.xc_region[0 * size_t(p.xc_stride) + s] <= R.lo.x && p.xc_region[1 * size_t(p.xc_stride) + s] <= R.lo.y &&
            p.xc_region[2 * size_t(p.xc_stride) + s] <= R.lo.z && p.xc_region[3 * size_t(p.xc_stride) + s] >= R.hi.x &&
            p.xc_region[4 * size_t(p.xc_stride) + s] >= R.hi.y && p.xc_region[5 * size_t(p.xc_stride) + s] >= R.hi.z)
            return;
    }
    if (reuse) { R.lo = R.lo - v3(p.xc_margin); R.hi = R.hi + v3(p.xc_margin); }
    uint32_t node = 0, ncache = 0;
    while (node < p.n_nodes) {
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

__global__ void __launch_bounds__(32, 20) k_step_stream(StepParams p) {
    constexpr unsigned FULL = 0xffffffffu;
    __shared__ uint32_t s_leaf[LEAF_CAP][32];              // candidate leaves of the substep: node indices, emission order
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

    auto fetch_one = [&]() {
        while (f_cur == f_end) {
            if (f_leaf == nleaf) return;
            const uint32_t leaf = s_leaf[f_leaf][tid];
            ++f_leaf;
            f_cur = uint32_t(__float_as_int(__ldg(p.nodes + 2 * leaf).w));
            f_end = f_cur + uint32_t(__float_as_int(__ldg(p.nodes + 2 * leaf + 1).w));
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
                if (ov) { s_leaf[nleaf][tid] = node; ++nleaf; }
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
        if (m_load && (__popc(m_load) >= load_batch || (m_sub | m_stream) == 0u)) {
            unsigned base = 0;
            if (tid == 0) base = atomicAdd(p.work_counter, (unsigned)__popc(m_load));
            base = __shfl_sync(FULL, base, 0);
            bool loaded_now = false;
            if (state == ST_LOAD) {
                const uint32_t idx = base + __popc(m_load & lt_mask);
                if (idx >= p.n) {
                    state = ST_DONE;
                } else {
                    i = idx;
                    const prtc_transform * tf = p.transforms + i;
                    const prtc_char_physics * ph = p.physics + i;
                    uint32_t cid = ph->collider;
                    if (cid >= p.n_capsules) {                    // unknown capsule id: never dereferenced; reported, result undefined
                        if (p.counters) atomicAdd(p.counters + 6, 1ull);
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
                        for (uint32_t j = 0; j < ncache; ++j) {
                            const uint32_t leaf = s_cache[j][tid];
                            const float4 lo = __ldg(p.nodes + 2 * leaf);
                            const float4 hi = __ldg(p.nodes + 2 * leaf + 1);
                            ++c_node;
                            if (node_overlaps(lo, hi, qb)) {
                                if (nleaf == LEAF_CAP) { fits = false; break; }
                                s_leaf[nleaf][tid] = leaf;
                                ++nleaf;
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

// K1 for the dynamic pass: per own character, the phase-1 predicted box (physics_system.cpp:263-273: pose box united
// with the box after `velocity + gravity*dt`, displacement applied in the capsule's rotated frame), the stored fat box
// hysteresis of DynamicAABBTree::update/insertLeaf (aabb_tree.cpp:102-111,140-141) and the start-of-frame segment.
// Record = 4 float4: (A.xyz, radius) (B.xyz, 0) (stored.lo, 0) (stored.hi, 0) at global index dyn_first + i.
__global__ void k_dyn_prep(const prtc_transform * __restrict__ tf, const prtc_char_physics * __restrict__ ph,
                           const CapsuleDev * __restrict__ capsules, uint32_t n, uint32_t first, float4 * __restrict__ records,
                           float dt, float gravity, float margin, uint32_t init) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const CapsuleDev cap = capsules[ph[i].collider];
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
    } else {
        const float4 lo = rec[2], hi = rec[3];
        stored.lo = v3(lo.x, lo.y, lo.z); stored.hi = v3(hi.x, hi.y, hi.z);
    }
    if (!box_contains(stored, e)) { stored.lo = e.lo - margin; stored.hi = e.hi + margin; }
    rec[0] = make_float4(A.x, A.y, A.z, cap.radius);
    rec[1] = make_float4(B.x, B.y, B.z, 0.0f);
    rec[2] = make_float4(stored.lo.x, stored.lo.y, stored.lo.z, 0.0f);
    rec[3] = make_float4(stored.hi.x, stored.hi.y, stored.hi.z, 0.0f);
}

// counting-sort build of the uniform grid over ALL capsules' stored boxes (pass 0 counts, pass 1 fills)
__global__ void k_dyn_grid(const float4 * __restrict__ records, uint32_t n, float ox, float oz, float inv_cell, uint32_t gx, uint32_t gz,
                           uint32_t * __restrict__ cell_count, const uint32_t * __restrict__ cell_start, uint32_t * __restrict__ items, int fill) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const float4 lo = records[4 * size_t(j) + 2], hi = records[4 * size_t(j) + 3];
    const int x0 = dyn_cell(lo.x, ox, inv_cell, gx), x1 = dyn_cell(hi.x, ox, inv_cell, gx);
    const int z0 = dyn_cell(lo.z, oz, inv_cell, gz), z1 = dyn_cell(hi.z, oz, inv_cell, gz);
    // a NaN corner puts the box in every cell it could matter for: (0,0)..(x1,z1) is what dyn_cell yields
    for (int z = z0; z <= z1; ++z)
        for (int x = x0; x <= x1; ++x) {
            const uint32_t cell = uint32_t(z) * gx + uint32_t(x);
            const uint32_t slot = atomicAdd(cell_count + cell, 1u);
            if (fill) items[cell_start[cell] + slot] = j;
        }
}

// PhysicsSystem::raycast (physics_system.cpp:205-242): one thread per ray.  DynamicAABBTree::queryRaycast
// (aabb_tree.cpp:70-93) with AABB::intersectRay (aabb.cpp:13-48, Ericson's slab test), then
// intersectLineSegmentTriangle (physics_util.cpp:181-216) on every triangle of every mesh leaf hit; the result is
// the minimum t, so the visiting order does not matter.  Capsule leaves of a unified tree are skipped (:215).
__device__ __forceinline__ bool ray_box(float4 lo, float4 hi, V3 o, V3 d, float maxDistance) {
    float tmin = 0.0f;
    float tmax = maxDistance;
    const float lov[3] = {lo.x, lo.y, lo.z}, hiv[3] = {hi.x, hi.y, hi.z};
    const float ov[3] = {o.x, o.y, o.z}, dv[3] = {d.x, d.y, d.z};
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        if (double(fabsf(dv[i])) < 0.00001) {                     // `std::abs(direction[i]) < 0.00001`: a double compare
            if (ov[i] < lov[i] || ov[i] > hiv[i]) return false;
        } else {
            const float ood = fdiv(1.0f, dv[i]);
            float t1 = (lov[i] - ov[i]) * ood;
            float t2 = (hiv[i] - ov[i]) * ood;
            if (t1 > t2) { const float tmp = t1; t1 = t2; t2 = tmp; }
            if (t1 > tmin) tmin = t1;
            if (t2 < tmax) tmax = t2;
            if (tmin > tmax) return false;
        }
    }
    return true;
}

__device__ __forceinline__ bool segment_triangle(V3 origin, V3 end, V3 a, V3 b, V3 c, float & t) {
    const V3 ab = b - a;
    const V3 ac = c - a;
    const V3 qp = origin - end;
    const V3 n = cross(ab, ac);
    const float d = dot(qp, n);
    if (d <= 0.0f) return false;
    const V3 ap = origin - a;
    t = dot(ap, n);
    if (t < 0.0f) return false;
    if (t > d) return false;
    const V3 e = cross(qp, ap);
    const float v = dot(ac, e);
    if (v < 0.0f || v > d) return false;
    const float w = -dot(ab, e);
    if (w < 0.0f || v + w > d) return false;
    const float ood = fdiv(1.0f, d);
    t = t * ood;
    return true;
}

__global__ void k_raycast(const float4 * __restrict__ nodes, uint32_t n_nodes, const float4 * __restrict__ tris, uint32_t n_rays,
                          const float * __restrict__ origins, const float * __restrict__ dirs, const float * __restrict__ maxd,
                          float * __restrict__ hit_xyz, unsigned char * __restrict__ hit_mask) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rays) return;
    const V3 o = v3(origins[3 * r], origins[3 * r + 1], origins[3 * r + 2]);
    const V3 d = v3(dirs[3 * r], dirs[3 * r + 1], dirs[3 * r + 2]);
    const float md = maxd[r];
    const V3 end = o + d * md;
    float best = 3.402823466e+38f;                                // std::numeric_limits<float>::max()
    bool any = false;
    uint32_t node = 0;
    while (node < n_nodes) {
        const float4 lo = nodes[2 * node];
        const float4 hi = nodes[2 * node + 1];
        const bool ov = ray_box(lo, hi, o, d, md);
        const int link = __float_as_int(lo.w);
        const int cnt = __float_as_int(hi.w);
        if (cnt < 0) {
            node = ov ? node + 1 : uint32_t(link);
        } else {
            if (ov && !(cnt & PRT_CAPSULE_LEAF)) {
                for (uint32_t t = uint32_t(link); t < uint32_t(link) + uint32_t(cnt); ++t) {
                    const float4 t0 = tris[3 * size_t(t)], t1 = tris[3 * size_t(t) + 1], t2 = tris[3 * size_t(t) + 2];
                    float tt;
                    if (segment_triangle(o, end, v3(t0.x, t0.y, t0.z), v3(t1.x, t1.y, t1.z), v3(t2.x, t2.y, t2.z), tt)) {
                        any = true;
                        best = (tt < best) ? tt : best;           // std::min(intersectionTime, t)
                    }
                }
            }
            node = node + 1;
        }
    }
    hit_mask[r] = any ? 1 : 0;
    if (any) {
        const V3 h = o + (d * best) * md;                         // origin + direction * intersectionTime * maxDistance
        hit_xyz[3 * r] = h.x; hit_xyz[3 * r + 1] = h.y; hit_xyz[3 * r + 2] = h.z;
    }
}

} // namespace

void launch_raycast(const float4 * nodes, uint32_t n_nodes, const float4 * tris, uint32_t n_rays, const float * origins,
                    const float * dirs, const float * maxd, float * hit_xyz, unsigned char * hit_mask, cudaStream_t stream) {
    if (n_rays == 0) return;
    const uint32_t block = 64;
    k_raycast<<<(n_rays + block - 1) / block, block, 0, stream>>>(nodes, n_nodes, tris, n_rays, origins, dirs, maxd, hit_xyz, hit_mask);
}

int launch_step(StepParams const & p_in, bool traced, int mode, cudaStream_t stream) {
    if (p_in.n == 0) return 0;
    int launches = 1;
    StepParams p = p_in;
    const uint32_t block = STEP_BLOCK;
    // fewer characters than resident lanes: thin the warps out (1..32 characters per warp)
    const uint32_t resident_warps = p.persistent_blocks ? p.persistent_blocks : 2368u;
    uint32_t lpw = (p.n + resident_warps - 1) / resident_warps;
    lpw = lpw < 1u ? 1u : (lpw > 8u ? 32u : lpw);       // measured: thinning pays up to ~8 lanes/warp (C2 0.29 -> 0.11 ms), not at 22 (C3)
    const bool small = lpw < 32u;
    p.lanes_per_warp = lpw;
    if (small) p.work_counter = nullptr;                 // the persistent kernel is for batches that fill the machine
    const uint32_t warps = (p.n + lpw - 1) / lpw;
    const uint32_t grid = (warps * 32u + block - 1) / block;
    // untraced CANONICAL launches read per-slot leaf lists: bring them up to date first
    // (only where the extra launch pays: the persistent kernel and big JACOBI batches; thinned / chunked lock-step
    //  launches walk inside k_step -- measured: C2 0.113 vs 0.148 ms, chunked e2e 3.21 vs 3.32 ms)
    const bool slots = !traced && mode != MODE_LITERAL && p.xc_version != nullptr && !(p.tune & 1u) &&
                       ((mode == MODE_STATIC && p.work_counter != nullptr) || (mode == MODE_JACOBI && !small));
    if (!slots) p.xc_version = nullptr;
    if (slots) { k_xc_refresh<<<(p.n + 127) / 128, 128, 0, stream>>>(p, p.xc_reuse ? 1 : 0); ++launches; }
    const bool warp_per_character = !traced && lpw == 1u && mode != MODE_JACOBI && !(p.tune & 2u);
    if (warp_per_character) {
        if (mode == MODE_LITERAL) k_step_warp<MODE_LITERAL><<<p.n, 32, 0, stream>>>(p);
        else k_step_warp<MODE_STATIC><<<p.n, 32, 0, stream>>>(p);
    } else if (mode == MODE_LITERAL) {
        if (traced) k_step<true, MODE_LITERAL><<<grid, block, 0, stream>>>(p);
        else k_step<false, MODE_LITERAL><<<grid, block, 0, stream>>>(p);
    } else if (mode == MODE_JACOBI) {
        if (traced) k_step<true, MODE_JACOBI><<<grid, block, 0, stream>>>(p);
        else k_step<false, MODE_JACOBI><<<grid, block, 0, stream>>>(p);
    } else {
        if (traced) k_step<true, MODE_STATIC><<<grid, block, 0, stream>>>(p);
        else if (p.work_counter) {
            // persistent: one warp per block, as many blocks as fit (20 per SM), work handed out by a counter
            cudaMemsetAsync(p.work_counter, 0, sizeof(unsigned), stream);
            const uint32_t want = (p.n + 31) / 32;
            const uint32_t resident = uint32_t(p.persistent_blocks);
            k_step_stream<<<want < resident ? want : resident, 32, 0, stream>>>(p);
        }
        else k_step<false, MODE_STATIC><<<grid, block, 0, stream>>>(p);
    }
    return launches;
}

void launch_dyn_prep(const prtc_transform * tf, const prtc_char_physics * ph, const CapsuleDev * capsules, uint32_t n, uint32_t first,
                     float4 * records, float dt, float gravity, float margin, bool init, cudaStream_t stream) {
    if (n == 0) return;
    k_dyn_prep<<<(n + 127) / 128, 128, 0, stream>>>(tf, ph, capsules, n, first, records, dt, gravity, margin, init ? 1u : 0u);
}

void launch_dyn_grid(const float4 * records, uint32_t n, float ox, float oz, float inv_cell, uint32_t gx, uint32_t gz,
                     uint32_t * cell_count, const uint32_t * cell_start, uint32_t * items, bool fill, cudaStream_t stream) {
    if (n == 0) return;
    k_dyn_grid<<<(n + 127) / 128, 128, 0, stream>>>(records, n, ox, oz, inv_cell, gx, gz, cell_count, cell_start, items, fill ? 1 : 0);
}

void launch_transform_refit(const float * raw_xyz, float * world_xyz, const MeshDev * meshes, uint32_t first_mesh,
                            uint32_t n_meshes, const ModelXform * xforms, float * mesh_aabbs, cudaStream_t stream) {
    if (n_meshes == 0) return;
    const uint32_t block = 128;
    k_transform_refit<<<(n_meshes + block - 1) / block, block, 0, stream>>>(raw_xyz, world_xyz, meshes, first_mesh,
                                                                           n_meshes, xforms, mesh_aabbs);
}

void launch_tri_setup(const float * world_xyz, const uint32_t * tri_src_vertex, uint32_t n_tris, float4 * tris,
                      cudaStream_t stream) {
    if (n_tris == 0) return;
    const uint32_t block = 256;
    k_tri_setup<<<(n_tris + block - 1) / block, block, 0, stream>>>(world_xyz, tri_src_vertex, n_tris, tris);
}

void launch_query_box(const float4 * nodes, uint32_t n_nodes, const uint32_t * node_leaf_id, const float * aabb6,
                      uint32_t * out_ids, uint32_t capacity, uint32_t * out_count, cudaStream_t stream) {
    k_query_box<<<1, 32, 0, stream>>>(nodes, n_nodes, node_leaf_id, aabb6, out_ids, capacity, out_count);
}

} // namespace prt
