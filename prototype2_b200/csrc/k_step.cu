// k_step.cu -- the lock-step step kernel: one character per lane, warp-uniform loops (all modes, traced or not)
// Compile with -fmad=false: results must be bit-identical to the CPU reference (see prt_math.cuh).
#include "step_common.cuh"

namespace prt {

namespace {

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
            if (p.host_flags) p.host_flags[1] = 1u;
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
            bool searching = alive;
            uint32_t t_caps = 0;
            long long last = -1;                                  // largest neighbour index already processed
            while (true) {
                uint32_t ncand = 0;
                bool more = false;
                if (searching) ncand = dyn_collect(dyn_grid_of(p), qb, self, last, &s_dyn[0][tid], more);
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
                if (!more) searching = false;                      // this lane is finished
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

} // namespace

void launch_k_step(StepParams const & p, bool traced, int mode, uint32_t grid, cudaStream_t stream) {
    const uint32_t block = STEP_BLOCK;
    if (mode == MODE_LITERAL) {
        if (traced) k_step<true, MODE_LITERAL><<<grid, block, 0, stream>>>(p);
        else k_step<false, MODE_LITERAL><<<grid, block, 0, stream>>>(p);
    } else if (mode == MODE_JACOBI) {
        if (traced) k_step<true, MODE_JACOBI><<<grid, block, 0, stream>>>(p);
        else k_step<false, MODE_JACOBI><<<grid, block, 0, stream>>>(p);
    } else {
        if (traced) k_step<true, MODE_STATIC><<<grid, block, 0, stream>>>(p);
        else k_step<false, MODE_STATIC><<<grid, block, 0, stream>>>(p);
    }
}

} // namespace prt
