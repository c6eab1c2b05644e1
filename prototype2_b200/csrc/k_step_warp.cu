// k_step_warp.cu -- one warp per character, for small batches
// Compile with -fmad=false: results must be bit-identical to the CPU reference (see prt_math.cuh).
#include "step_common.cuh"

namespace prt {

namespace {

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
    __shared__ uint32_t s_rmesh[WF_CAP], s_rcaps[WF_CAP];   // the frame's cached candidates (emission order)
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
        if (lane == 0 && p.host_flags) p.host_flags[1] = 1u;
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
    // Level-by-level frontier walk for box `b`: the overlapping mesh leaves end up in s_mesh, the capsule leaves
    // (LITERAL) in s_caps, both sorted by node index = emission order.  false: a list or the frontier outgrew WF_CAP.
    uint32_t c_node_lane = 0;                                      // node tests of this lane (summed at the end)
    auto frontier_walk = [&](Box const & b, uint32_t & nmesh, uint32_t & ncaps) -> bool {
        nmesh = 0; ncaps = 0;
        const uint32_t tests_before = c_node_lane;                 // an overflowing walk is redone by ordered passes, which count themselves
        uint32_t cur = 0, nfront = p.n_nodes ? 1u : 0u;
        if (lane == 0) s_front[0][0] = 0u;
        __syncwarp(FULL);
        while (nfront != 0) {
            uint32_t nnext = 0;
            for (uint32_t base = 0; base < nfront; base += 32) {
                const uint32_t k = base + lane;
                bool internal = false, leaf_mesh = false, leaf_caps = false;
                uint32_t idx = 0, second = 0;
                if (k < nfront) {
                    idx = s_front[cur][k];
                    const float4 lo = __ldg(p.nodes + 2 * idx);
                    const float4 hi = __ldg(p.nodes + 2 * idx + 1);
                    ++c_node_lane;
                    if (node_overlaps(lo, hi, b)) {
                        const int cnt = __float_as_int(hi.w);
                        if (cnt < 0) { internal = true; second = uint32_t(-cnt); }
                        else if (!(cnt & PRT_CAPSULE_LEAF)) leaf_mesh = true;
                        else if (LITERAL && uint32_t(__float_as_int(lo.w)) != cid) leaf_caps = true;
                    }
                }
                const unsigned m_int = __ballot_sync(FULL, internal);
                const unsigned m_mesh = __ballot_sync(FULL, leaf_mesh);
                const unsigned m_caps = __ballot_sync(FULL, leaf_caps);
                if (nnext + 2 * __popc(m_int) > WF_CAP || nmesh + __popc(m_mesh) > WF_CAP || ncaps + __popc(m_caps) > WF_CAP) {
                    c_node_lane = tests_before;
                    return false;
                }
                if (internal) {
                    const uint32_t o = nnext + 2 * __popc(m_int & lt);
                    s_front[cur ^ 1][o] = idx + 1;
                    s_front[cur ^ 1][o + 1] = second;
                }
                if (leaf_mesh) s_found[nmesh + __popc(m_mesh & lt)] = idx;
                if (leaf_caps) s_caps[ncaps + __popc(m_caps & lt)] = idx;
                nnext += 2 * __popc(m_int);
                nmesh += __popc(m_mesh);
                ncaps += __popc(m_caps);
            }
            cur ^= 1;
            nfront = nnext;
            __syncwarp(FULL);
        }
        // sort by node index (rank by counting; the lists are short)
        for (uint32_t j = lane; j < nmesh; j += 32) {
            const uint32_t e = s_found[j];
            uint32_t r = 0;
            for (uint32_t t = 0; t < nmesh; ++t) r += s_found[t] < e;
            s_mesh[r] = e;
        }
        if (LITERAL) {
            uint32_t mine[WF_CAP / 32], rk[WF_CAP / 32];
#pragma unroll
            for (int q4 = 0; q4 < WF_CAP / 32; ++q4) {
                const uint32_t j = lane + 32 * q4;
                mine[q4] = j < ncaps ? s_caps[j] : 0xFFFFFFFFu;
                rk[q4] = 0;
                for (uint32_t t = 0; t < ncaps; ++t) rk[q4] += s_caps[t] < mine[q4];
            }
            __syncwarp(FULL);
#pragma unroll
            for (int q4 = 0; q4 < WF_CAP / 32; ++q4) if (lane + 32 * q4 < ncaps) s_caps[rk[q4]] = mine[q4];
        }
        __syncwarp(FULL);
        return true;
    };

    // Frame cache (PRTC_OPT_FRAME_CACHE): ONE frontier walk with a box that covers the whole frame's motion (pose now,
    // pose after the pre-stepped velocity, 0.25 of slack for push-outs); a substep whose query box lies inside that
    // region only filters the cached leaves -- one memory round trip instead of one per tree level.  The cached lists
    // are supersets in emission order, the filter is the exact box test, so the candidates are the walk's.
    Box region; region.lo = v3(0.0f); region.hi = v3(0.0f);
    bool region_ok = false;
    uint32_t nr_mesh = 0, nr_caps = 0;
    if (!(p.tune & 1u) && p.substeps > 1) {
        V3 eA, eB;
        segment_ends(fr, pos + vel, eA, eB);
        region = box_union(capsule_box(prevA, prevB, radius), capsule_box(eA, eB, radius));
        region.lo = region.lo - v3(0.25f); region.hi = region.hi + v3(0.25f);
        region_ok = frontier_walk(region, nr_mesh, nr_caps);
        if (region_ok) {
            for (uint32_t j = lane; j < nr_mesh; j += 32) s_rmesh[j] = s_mesh[j];
            for (uint32_t j = lane; j < nr_caps; j += 32) s_rcaps[j] = s_caps[j];
            __syncwarp(FULL);
        }
    }

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

        // ---- broadphase: the frame's cached candidates filtered by this substep's box, or a frontier walk ----
        uint32_t nmesh = 0, ncaps = 0;
        bool overflow = false;
        if (region_ok && box_contains(region, qb)) {
            // both cached lists are in emission order and ballot compaction keeps it
            for (uint32_t base = 0; base < nr_mesh; base += 32) {
                const uint32_t k = base + lane;
                bool ov = false;
                uint32_t idx = 0;
                if (k < nr_mesh) {
                    idx = s_rmesh[k];
                    ov = node_overlaps(__ldg(p.nodes + 2 * idx), __ldg(p.nodes + 2 * idx + 1), qb);
                    ++c_node_lane;
                }
                const unsigned m = __ballot_sync(FULL, ov);
                if (ov) s_mesh[nmesh + __popc(m & lt)] = idx;
                nmesh += __popc(m);
            }
            if (LITERAL) {
                for (uint32_t base = 0; base < nr_caps; base += 32) {
                    const uint32_t k = base + lane;
                    bool ov = false;
                    uint32_t idx = 0;
                    if (k < nr_caps) {
                        idx = s_rcaps[k];
                        ov = node_overlaps(__ldg(p.nodes + 2 * idx), __ldg(p.nodes + 2 * idx + 1), qb);
                        ++c_node_lane;
                    }
                    const unsigned m = __ballot_sync(FULL, ov);
                    if (ov) s_caps[ncaps + __popc(m & lt)] = idx;
                    ncaps += __popc(m);
                }
            }
            __syncwarp(FULL);
        } else {
            overflow = !frontier_walk(qb, nmesh, ncaps);
        }
        // ordered passes (overflow only): the stackless right-first walk, executed uniformly by the whole warp
        uint32_t onode = p.n_nodes, cnode = p.n_nodes;            // resume positions of the mesh / capsule passes
        if (overflow) { onode = 0; cnode = 0; nmesh = 0; ncaps = 0; }

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
    c_node += __reduce_add_sync(FULL, c_node_lane);
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
        if (LITERAL && p.host_flags && c_cc) p.host_flags[0] = 1u;
    }
    if (p.done_flag) {                                             // completion doorbell (StepParams::done_flag): the records are host-mapped
        __threadfence_system();
        __syncwarp(FULL);
        if (lane == 0 && atomicAdd(p.done_count, 1u) + 1u == gridDim.x) {
            *p.done_count = 0u;
            __threadfence_system();
            *reinterpret_cast<volatile uint32_t *>(p.done_flag) = p.done_value;
        }
    }
}

} // namespace

void launch_k_step_warp(StepParams const & p, int mode, cudaStream_t stream) {
    if (mode == MODE_LITERAL) k_step_warp<MODE_LITERAL><<<p.n, 32, 0, stream>>>(p);
    else k_step_warp<MODE_STATIC><<<p.n, 32, 0, stream>>>(p);
}

} // namespace prt
