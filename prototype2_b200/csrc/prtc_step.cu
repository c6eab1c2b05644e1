// prtc_step.cu -- host side of the C-ABI declared in include/prtc.h, part 2: the per-frame entry points
// (prtc_step_characters*, the LITERAL fixpoint, the staged and the streamed host-buffer steps, the dynamic pass plumbing,
// prtc_raycast, prtc_sweep_ellipsoids) and the trace read-back.
#include "prtc_internal.h"

using namespace prt;
using namespace prt::host;

namespace {

// Completion doorbell of the staged (mapped pinned memory) engine-sized paths: the kernel's last warp writes a sequence number
// to a word of the staging block, the host spins on it -- a few microseconds after the kernel ends instead of the wake-up
// of cudaStreamSynchronize.  Every few thousand spins the stream is asked whether it finished (or died) without ringing.
uint32_t arm_bell(prtc_ctx * c, StepParams & p, uint32_t * word) {
    if (!c->doorbell) return 0u;
    *word = 0u;
    uint32_t v = ++c->resident_frame;
    if (v == 0u) v = ++c->resident_frame;
    p.done_flag = word; p.done_count = c->d_work.ptr + 6; p.done_value = v;
    return v;
}
int wait_bell(prtc_ctx * c, const uint32_t * word, uint32_t value) {
    if (!value) { PRTC_CUDA(cudaStreamSynchronize(c->stream)); return PRTC_OK; }
    const volatile uint32_t * flag = word;
    uint32_t spins = 0;
    while (*flag != value) {
        if ((++spins & 0xFFFu) == 0u) {
            const cudaError_t q = cudaStreamQuery(c->stream);
            if (q == cudaSuccess) break;
            if (q != cudaErrorNotReady) return fail(PRTC_ERR_CUDA, std::string("stream failed: ") + cudaGetErrorString(q));
        }
    }
    return PRTC_OK;
}

// copies the device-side trace slots of one traced launch into the caller's prtc_trace
int gather_trace(prtc_ctx * c, uint32_t n, uint32_t cap_cand, uint32_t cap_hits, prtc_trace * tr) {
    cudaStream_t s = c->stream;
    const uint32_t nsub = c->cfg.substeps;
    const size_t slots = size_t(n) * nsub;
    std::vector<TraceQuery> tq(slots);
    std::vector<uint32_t> tmesh(slots * cap_cand), tcaps(slots * cap_cand);
    std::vector<TraceHit> thits(slots * cap_hits);
    PRTC_CUDA(cudaMemcpyAsync(tq.data(), c->d_tq.ptr, slots * sizeof(TraceQuery), cudaMemcpyDeviceToHost, s));
    PRTC_CUDA(cudaMemcpyAsync(tmesh.data(), c->d_tmesh.ptr, tmesh.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    PRTC_CUDA(cudaMemcpyAsync(tcaps.data(), c->d_tcaps.ptr, tcaps.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    PRTC_CUDA(cudaMemcpyAsync(thits.data(), c->d_thits.ptr, thits.size() * sizeof(TraceHit), cudaMemcpyDeviceToHost, s));
    PRTC_CUDA(cudaStreamSynchronize(s));
    uint32_t nq = 0, nm = 0, nc = 0, nh = 0;
    bool overflow = false;
    if (tr->q_mesh_off && tr->cap_queries) tr->q_mesh_off[0] = 0;
    if (tr->q_caps_off && tr->cap_queries) tr->q_caps_off[0] = 0;
    for (uint32_t i = 0; i < n; ++i) {
        for (uint32_t sub = 0; sub < nsub; ++sub) {
            size_t slot = size_t(i) * nsub + sub;
            TraceQuery const & q = tq[slot];
            if (!q.executed) continue;
            if (q.n_mesh > cap_cand || q.n_caps > cap_cand || q.n_hits > cap_hits) overflow = true;
            if (nq >= tr->cap_queries) { overflow = true; continue; }
            tr->q_char[nq] = i;
            tr->q_sub[nq] = sub;
            for (int k = 0; k < 6; ++k) tr->q_aabb[6 * size_t(nq) + k] = q.aabb[k];
            tr->q_tri_tests[nq] = q.tri_tests;
            for (uint32_t k = 0; k < std::min(q.n_mesh, cap_cand); ++k) {
                if (nm < tr->cap_mesh_ids) tr->mesh_ids[nm++] = tmesh[slot * cap_cand + k]; else overflow = true;
            }
            for (uint32_t k = 0; k < std::min(q.n_caps, cap_cand); ++k) {
                if (nc < tr->cap_caps_ids) tr->caps_ids[nc++] = tcaps[slot * cap_cand + k]; else overflow = true;
            }
            for (uint32_t k = 0; k < std::min(q.n_hits, cap_hits); ++k) {
                if (nh >= tr->cap_hits) { overflow = true; break; }
                TraceHit const & h = thits[slot * cap_hits + k];
                tr->h_query[nh] = nq;
                tr->h_poly[nh] = h.poly;
                for (int d = 0; d < 3; ++d) {
                    tr->h_impulse[3 * size_t(nh) + d] = h.impulse[d];
                    tr->h_normal[3 * size_t(nh) + d] = h.normal[d];
                    tr->h_pos[3 * size_t(nh) + d] = h.pos[d];
                }
                ++nh;
            }
            ++nq;
            tr->q_mesh_off[nq] = nm;
            tr->q_caps_off[nq] = nc;
        }
    }
    tr->n_queries = nq; tr->n_mesh_ids = nm; tr->n_caps_ids = nc; tr->n_hits = nh;
    if (overflow) return fail(PRTC_ERR_CAPACITY, "prtc_step_characters_traced: trace capacity exceeded (state is valid, trace is truncated)");
    return PRTC_OK;
}

constexpr uint32_t TRACE_CAP_CAND = 64, TRACE_CAP_HITS = 96;

int prepare_trace(prtc_ctx * c, uint32_t n, StepParams & p) {
    { int rc = upload_leaf_ids(c); if (rc) return rc; }
    p.node_leaf_id = c->d_node_leaf_id.ptr;
    const size_t slots = size_t(n) * c->cfg.substeps;
    PRTC_CUDA(c->d_tq.reserve(slots));
    PRTC_CUDA(c->d_tmesh.reserve(slots * TRACE_CAP_CAND));
    PRTC_CUDA(c->d_tcaps.reserve(slots * TRACE_CAP_CAND));
    PRTC_CUDA(c->d_thits.reserve(slots * TRACE_CAP_HITS));
    PRTC_CUDA(cudaMemsetAsync(c->d_tq.ptr, 0, slots * sizeof(TraceQuery), c->stream));
    p.trace.queries = c->d_tq.ptr; p.trace.mesh_ids = c->d_tmesh.ptr; p.trace.caps_ids = c->d_tcaps.ptr; p.trace.hits = c->d_thits.ptr;
    p.trace.cap_cand = TRACE_CAP_CAND; p.trace.cap_hits = TRACE_CAP_HITS;
    return PRTC_OK;
}

} // namespace

namespace prt {
namespace host {

int dyn_reserve(prtc_ctx * c, uint32_t n) {
    if (!c->dyn_configured) { c->dyn_first = 0; c->dyn_global = n; }
    if (uint64_t(c->dyn_first) + n > c->dyn_global) return fail(PRTC_ERR_INVALID, "prtc_dyn_prepare: slice exceeds the configured global count");
    // (two buffers when the records are exchanged across processes, see prtc_ctx::DynIpc)
    const size_t want = size_t(c->dyn_global) * 4 * (c->dyn_ipc.world ? 2 : 1);
    if (c->d_dyn_records.cap < want) {
        if (c->dyn_initialized) return fail(PRTC_ERR_INVALID, "prtc_dyn_prepare: the global character count may not grow after the first frame");
        PRTC_CUDA(c->d_dyn_records.reserve(want));
        PRTC_CUDA(cudaMemsetAsync(c->d_dyn_records.ptr, 0, want * sizeof(float4), c->stream));
    }
    if (!c->d_dyn_reach.ptr) {
        PRTC_CUDA(c->d_dyn_reach.reserve(2));
        PRTC_CUDA(cudaMemsetAsync(c->d_dyn_reach.ptr, 0, c->d_dyn_reach.cap * sizeof(uint32_t), c->stream));
    }
    return PRTC_OK;
}

// JACOBI dynamic pass: records of this context's characters (phase-1 box, stored-box hysteresis, start-of-frame segment)
int dyn_prepare(prtc_ctx * c, float dt, uint32_t n, const prtc_transform * d_tf, const prtc_char_physics * d_ph) {
    int rc = dyn_reserve(c, n);
    if (rc) return rc;
    if (c->dyn_ipc.on) {
        // this frame's buffer: own and peers'; the stored boxes of the last frame are in the other one; the kernel's last block
        // writes the frame number into this rank's word of everybody's arrival table
        prtc_ctx::DynIpc & x = c->dyn_ipc;
        const size_t half = size_t(c->dyn_global) * 4;
        const uint32_t b = x.frame & 1u;
        DynPeers peers = {};
        for (uint32_t k = 0; k < x.n_peers; ++k) { peers.records[peers.n++] = x.peer_records[k] + b * half; peers.flags[peers.n_flags++] = x.peer_flags[k] + x.rank; }
        peers.flags[peers.n_flags++] = x.flags.ptr + x.rank;
        peers.flag_value = x.frame + 1u;
        peers.block_count = x.flags.ptr + 31;
        peers.push_separately = 1u;
        launch_dyn_prep(d_tf, d_ph, c->d_capsules.ptr, uint32_t(c->capsules.size()), n, c->dyn_first, c->d_dyn_records.ptr + (b ^ 1u) * half,
                        c->d_dyn_records.ptr + b * half, peers, dt, c->cfg.gravity, c->cfg.leaf_margin, !c->dyn_initialized, c->d_dyn_reach.ptr, c->stream);
        // the peer stores leave on a stream of their own, so the slot refresh that follows on the context's stream runs beside
        // them; the context's stream meets them again in k_dyn_wait (this rank's own arrival word is written by the same kernel)
        PRTC_CUDA(cudaEventRecord(x.prep_done, c->stream));
        PRTC_CUDA(cudaStreamWaitEvent(x.push_stream, x.prep_done, 0));
        launch_dyn_push(c->d_dyn_records.ptr + b * half, c->dyn_first, n, peers, x.push_stream);
    } else {
        launch_dyn_prep(d_tf, d_ph, c->d_capsules.ptr, uint32_t(c->capsules.size()), n, c->dyn_first, c->d_dyn_records.ptr, c->d_dyn_records.ptr,
                        c->dyn_peers, dt, c->cfg.gravity, c->cfg.leaf_margin, !c->dyn_initialized, c->d_dyn_reach.ptr, c->stream);
    }
    PRTC_CUDA(cudaGetLastError());
    c->dyn_initialized = true;
    c->dyn_prepared = true;
    return PRTC_OK;
}

} // namespace host
} // namespace prt

namespace {

// uniform xz grid over the stored boxes of ALL characters (counting sort), wired into the launch parameters
int dyn_build_grid(prtc_ctx * c, StepParams & p) {
    cudaStream_t s = c->stream;
    float lox = c->world_lo[0], loz = c->world_lo[2], hix = c->world_hi[0], hiz = c->world_hi[2];
    if (!(hix > lox) || !(hiz > loz)) { lox = loz = 0.0f; hix = hiz = 1.0f; }          // no static geometry: one cell
    lox -= 2.0f; loz -= 2.0f; hix += 2.0f; hiz += 2.0f;
    const float ext = std::max(hix - lox, hiz - loz);
    const float cell = std::max(2.5f, ext / 1024.0f);
    const uint32_t gx = std::max(1u, uint32_t(std::ceil((hix - lox) / cell))), gz = std::max(1u, uint32_t(std::ceil((hiz - loz) / cell)));
    const size_t cells = size_t(gx) * gz;
    const uint32_t ng = c->dyn_global;
    // every stored box is binned in exactly one cell, so the item array has ng entries: nothing here depends on a device count
    // (cell_count is all zero between frames: zeroed when allocated, k_dyn_scan zeroes what a frame counted)
    if (c->d_dyn_count.cap < cells + 8) {
        PRTC_CUDA(c->d_dyn_count.reserve(cells + 8));
        PRTC_CUDA(cudaMemsetAsync(c->d_dyn_count.ptr, 0, c->d_dyn_count.cap * sizeof(uint32_t), s));
    }
    if (!c->d_dyn_scan.ptr) {
        PRTC_CUDA(c->d_dyn_scan.reserve(DYN_SCAN_WORDS + 2));
        PRTC_CUDA(cudaMemsetAsync(c->d_dyn_scan.ptr, 0, c->d_dyn_scan.cap * sizeof(unsigned long long), s));
        c->dyn_scan_tickets = 0; c->dyn_scan_launches = 0;
    }
    PRTC_CUDA(c->d_dyn_start.reserve(cells + 8));
    PRTC_CUDA(c->d_dyn_cursor.reserve(cells + 8));
    PRTC_CUDA(c->d_dyn_items.reserve(size_t(ng) + 1));
    const float4 * records = c->d_dyn_records.ptr;
    if (c->dyn_ipc.on) {
        // every rank's records of this frame must have landed in this device's buffer: wait for the arrival words, on the device
        prtc_ctx::DynIpc & x = c->dyn_ipc;
        launch_dyn_wait(x.flags.ptr, x.world, x.frame + 1u, c->d_counters.ptr, s);
        records = c->d_dyn_records.ptr + (x.frame & 1u) * size_t(c->dyn_global) * 4;
        ++x.frame;
        c->launches += 1;
    }
    c->launches += uint64_t(launch_dyn_grid_build(records, ng, lox, loz, 1.0f / cell, gx, gz, c->d_dyn_count.ptr, c->d_dyn_start.ptr,
                                                  c->d_dyn_cursor.ptr, c->d_dyn_items.ptr, c->d_dyn_reach.ptr, c->d_dyn_scan.ptr,
                                                  c->dyn_scan_tickets, c->dyn_scan_launches, s));
    PRTC_CUDA(cudaGetLastError());
    p.dyn_records = records;
    p.dyn_cell_start = c->d_dyn_start.ptr;
    p.dyn_items = c->d_dyn_items.ptr;
    p.dyn_reach = c->d_dyn_reach.ptr;
    p.dyn_first = c->dyn_first;
    p.dyn_gx = gx; p.dyn_gz = gz;
    p.dyn_origin_x = lox; p.dyn_origin_z = loz; p.dyn_inv_cell = 1.0f / cell;
    return PRTC_OK;
}

// One frame in LITERAL order (host buffers).  Everything the reference does to its unified tree per frame is
// replayed on the host with the same algorithm; the device runs the collide phase; the reference's sequential
// visibility between characters (T18) is reached as the fixpoint of repeated whole-batch launches.
// literal_step for an engine-sized batch (no trace): the records live in mapped pinned memory for the duration of the
// call, the warp-per-character kernel reads and writes them there, and "did any capsule-capsule test run" comes back as
// a flag in the same memory -- one launch (+ the counter fold) and one synchronisation per fixpoint round instead of
// six copies and three synchronisations.  Device counters of a round go to a scratch block so that only the last round
// is added to the context's counters.
int literal_step_staged(prtc_ctx * c, float dt, uint32_t n, prtc_transform * tf, prtc_char_physics * ph,
                        std::vector<float4> const & seg_start, std::vector<uint32_t> const & cap2char) {
    cudaStream_t s = c->stream;
    const size_t bytes = 64 + 64 * 8 + size_t(n) * (sizeof(prtc_transform) + sizeof(prtc_char_physics) + 4 * sizeof(float4) + 6 * sizeof(float)) +
                         cap2char.size() * sizeof(uint32_t);
    PRTC_CUDA(c->stage.reserve(bytes));
    StageCursor cur{c->stage.base};
    uint32_t * flags = cur.take<uint32_t>(16);
    prtc_transform * s_tf = cur.take<prtc_transform>(n);
    prtc_char_physics * s_ph = cur.take<prtc_char_physics>(n);
    float4 * s_seg_start = cur.take<float4>(2 * size_t(n));
    float4 * s_seg_prev = cur.take<float4>(2 * size_t(n));
    float * s_last_box = cur.take<float>(6 * size_t(n));
    uint32_t * s_cap2char = cur.take<uint32_t>(cap2char.size());
    std::memcpy(s_seg_start, seg_start.data(), seg_start.size() * sizeof(float4));
    std::memcpy(s_seg_prev, seg_start.data(), seg_start.size() * sizeof(float4));
    std::memcpy(s_cap2char, cap2char.data(), cap2char.size() * sizeof(uint32_t));
    unsigned long long * scratch = c->d_counters.ptr + 8;

    std::vector<prtc_transform> prev_tf;
    const uint32_t max_rounds = n + 2;
    for (uint32_t round = 0; round < max_rounds; ++round) {
        std::memcpy(s_tf, tf, size_t(n) * sizeof(prtc_transform));
        std::memcpy(s_ph, ph, size_t(n) * sizeof(prtc_char_physics));
        flags[0] = 0u; flags[1] = 0u;
        if (round > 0) PRTC_CUDA(cudaMemsetAsync(scratch, 0, 6 * sizeof(unsigned long long), s));       // only the last round counts
        StepParams p;
        { const int frc = fill_params(c, dt, n, s_tf, s_ph, p); if (frc) return frc; }
        p.cap2char = s_cap2char; p.seg_start = s_seg_start; p.seg_prev = s_seg_prev; p.last_box = s_last_box;
        p.counters = scratch;
        p.host_flags = flags;
        const uint32_t bell = arm_bell(c, p, flags + 2);
        c->launches += uint64_t(launch_step(p, false, STEP_LITERAL, s));
        PRTC_CUDA(cudaGetLastError());
        { const int wrc = wait_bell(c, flags + 2, bell); if (wrc) return wrc; }
        const bool any_capsule_test = flags[0] != 0u;
        const bool settled = !prev_tf.empty() && std::memcmp(prev_tf.data(), s_tf, size_t(n) * sizeof(prtc_transform)) == 0;
        if (!any_capsule_test || settled) break;
        // next round: characters j < i are seen where this round left them
        for (uint32_t i = 0; i < n; ++i) {
            CapsuleDev const & cap = c->capsules[ph[i].collider];
            Quat q; q.x = tf[i].rotation[0]; q.y = tf[i].rotation[1]; q.z = tf[i].rotation[2]; q.w = tf[i].rotation[3];
            V3 A, B;
            segment_ends(capsule_frame(cap, rotation_of(q)), v3(s_tf[i].position[0], s_tf[i].position[1], s_tf[i].position[2]), A, B);
            s_seg_prev[2 * size_t(i)] = make_float4(A.x, A.y, A.z, cap.radius);
            s_seg_prev[2 * size_t(i) + 1] = make_float4(B.x, B.y, B.z, 0.0f);
        }
        prev_tf.assign(s_tf, s_tf + n);
    }
    launch_fold_counters(c->d_counters.ptr, scratch, s);           // asynchronous: nobody waits for it here
    c->launches += 1;
    PRTC_CUDA(cudaGetLastError());
    // m_aabbData.capsuleAABBs[tag.index] is left at the last executed substep's query box (physics_system.cpp:355-356)
    for (uint32_t i = 0; i < n; ++i) {
        Box b;
        b.lo = v3(s_last_box[6 * size_t(i)], s_last_box[6 * size_t(i) + 1], s_last_box[6 * size_t(i) + 2]);
        b.hi = v3(s_last_box[6 * size_t(i) + 3], s_last_box[6 * size_t(i) + 4], s_last_box[6 * size_t(i) + 5]);
        c->capsule_aabbs[ph[i].collider] = b;
    }
    std::memcpy(tf, s_tf, size_t(n) * sizeof(prtc_transform));
    std::memcpy(ph, s_ph, size_t(n) * sizeof(prtc_char_physics));
    return PRTC_OK;
}

int literal_step(prtc_ctx * c, float dt, uint32_t n, prtc_transform * tf, prtc_char_physics * ph, prtc_trace * tr) {
    cudaStream_t s = c->stream;
    for (uint32_t i = 0; i < n; ++i)
        if (ph[i].collider >= c->capsules.size()) return fail(PRTC_ERR_INVALID, "prtc_step_characters: collider id out of range");

    // phase 1 AABBs (physics_system.cpp:263-273): pose box united with the box after `velocity + gravity*dt`,
    // the displacement being applied in the capsule's rotated frame (glm::translate(tform, v), trap T3)
    std::vector<float4> seg_start(2 * size_t(n));
    std::vector<uint32_t> cap2char(c->capsules.size(), 0u);
    for (uint32_t i = 0; i < n; ++i) {
        CapsuleDev const & cap = c->capsules[ph[i].collider];
        Quat q; q.x = tf[i].rotation[0]; q.y = tf[i].rotation[1]; q.z = tf[i].rotation[2]; q.w = tf[i].rotation[3];
        const Rot3 R = rotation_of(q);
        const CapsuleFrame fr = capsule_frame(cap, R);
        const V3 pos = v3(tf[i].position[0], tf[i].position[1], tf[i].position[2]);
        V3 A, B;
        segment_ends(fr, pos, A, B);
        Box e = capsule_box(A, B, cap.radius);
        const V3 vel = v3(ph[i].velocity[0], ph[i].velocity[1], ph[i].velocity[2]);
        const V3 d = vel + (v3(0.0f, -1.0f, 0.0f) * c->cfg.gravity) * dt;
        const V3 pos2 = ((R.c0 * d.x + R.c1 * d.y) + R.c2 * d.z) + pos;       // translate(): m[0]*v.x + m[1]*v.y + m[2]*v.z + m[3]
        V3 A2, B2;
        segment_ends(fr, pos, pos2, A2, B2);
        e = box_union(e, capsule_box(A2, B2, cap.radius));
        c->capsule_aabbs[ph[i].collider] = e;
        cap2char[ph[i].collider] = i;                                          // tagToCharacter, :275
        seg_start[2 * size_t(i)] = make_float4(A.x, A.y, A.z, cap.radius);
        seg_start[2 * size_t(i) + 1] = make_float4(B.x, B.y, B.z, 0.0f);
    }
    // m_aabbData.tree.update(all capsule leaves) (physics_system.cpp:289)
    if (!c->capsule_tree_index.empty() &&
        c->tree.update(c->capsule_tree_index.data(), c->capsule_aabbs.data(), c->capsule_tree_index.size()) != 0)
        c->tree_dirty = true;                                                   // some leaf left its fat box: the tree changed shape
    int rc = commit_impl(c);                                                    // re-flatten + upload (+ triangles in the new leaf order) if needed
    if (rc) return rc;

    if (!tr && n > 0 && n <= kSmallBatch && cap2char.size() <= 16 * kSmallBatch && c->warp_kernel && c->staged)
        return literal_step_staged(c, dt, n, tf, ph, seg_start, cap2char);

    PRTC_CUDA(c->d_tf.reserve(n));
    PRTC_CUDA(c->d_ph.reserve(n));
    PRTC_CUDA(c->d_cap2char.reserve(cap2char.size()));
    PRTC_CUDA(c->d_seg_start.reserve(seg_start.size()));
    PRTC_CUDA(c->d_seg_prev.reserve(seg_start.size()));
    PRTC_CUDA(c->d_last_box.reserve(6 * size_t(n)));
    if (n == 0) { if (tr) { tr->n_queries = tr->n_mesh_ids = tr->n_caps_ids = tr->n_hits = 0; if (tr->q_mesh_off) tr->q_mesh_off[0] = 0; if (tr->q_caps_off) tr->q_caps_off[0] = 0; } return PRTC_OK; }
    PRTC_CUDA(cudaMemcpyAsync(c->d_cap2char.ptr, cap2char.data(), cap2char.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
    PRTC_CUDA(cudaMemcpyAsync(c->d_seg_start.ptr, seg_start.data(), seg_start.size() * sizeof(float4), cudaMemcpyHostToDevice, s));

    unsigned long long base[8];
    PRTC_CUDA(cudaMemcpyAsync(base, c->d_counters.ptr, sizeof(base), cudaMemcpyDeviceToHost, s));
    PRTC_CUDA(cudaStreamSynchronize(s));

    std::vector<prtc_transform> out_tf(n), prev_tf;
    std::vector<prtc_char_physics> out_ph(n);
    std::vector<float4> seg_prev(seg_start);
    const uint32_t max_rounds = n + 2;
    for (uint32_t round = 0; round < max_rounds; ++round) {
        PRTC_CUDA(cudaMemcpyAsync(c->d_counters.ptr, base, sizeof(base), cudaMemcpyHostToDevice, s));   // only the last round counts
        PRTC_CUDA(cudaMemcpyAsync(c->d_tf.ptr, tf, size_t(n) * sizeof(prtc_transform), cudaMemcpyHostToDevice, s));
        PRTC_CUDA(cudaMemcpyAsync(c->d_ph.ptr, ph, size_t(n) * sizeof(prtc_char_physics), cudaMemcpyHostToDevice, s));
        PRTC_CUDA(cudaMemcpyAsync(c->d_seg_prev.ptr, seg_prev.data(), seg_prev.size() * sizeof(float4), cudaMemcpyHostToDevice, s));
        StepParams p;
        rc = fill_params(c, dt, n, c->d_tf.ptr, c->d_ph.ptr, p);
        if (rc) return rc;
        p.cap2char = c->d_cap2char.ptr; p.seg_start = c->d_seg_start.ptr; p.seg_prev = c->d_seg_prev.ptr; p.last_box = c->d_last_box.ptr;
        if (tr) { rc = prepare_trace(c, n, p); if (rc) return rc; }
        c->launches += uint64_t(launch_step(p, tr != nullptr, STEP_LITERAL, s));
        PRTC_CUDA(cudaGetLastError());
        PRTC_CUDA(cudaMemcpyAsync(out_tf.data(), c->d_tf.ptr, size_t(n) * sizeof(prtc_transform), cudaMemcpyDeviceToHost, s));
        PRTC_CUDA(cudaMemcpyAsync(out_ph.data(), c->d_ph.ptr, size_t(n) * sizeof(prtc_char_physics), cudaMemcpyDeviceToHost, s));
        unsigned long long now[8];
        PRTC_CUDA(cudaMemcpyAsync(now, c->d_counters.ptr, sizeof(now), cudaMemcpyDeviceToHost, s));
        PRTC_CUDA(cudaStreamSynchronize(s));
        const bool any_capsule_test = now[4] != base[4];
        const bool settled = !prev_tf.empty() && std::memcmp(prev_tf.data(), out_tf.data(), size_t(n) * sizeof(prtc_transform)) == 0;
        if (!any_capsule_test || settled) break;
        // next round: characters j < i are seen where this round left them
        for (uint32_t i = 0; i < n; ++i) {
            CapsuleDev const & cap = c->capsules[ph[i].collider];
            Quat q; q.x = tf[i].rotation[0]; q.y = tf[i].rotation[1]; q.z = tf[i].rotation[2]; q.w = tf[i].rotation[3];
            V3 A, B;
            segment_ends(capsule_frame(cap, rotation_of(q)), v3(out_tf[i].position[0], out_tf[i].position[1], out_tf[i].position[2]), A, B);
            seg_prev[2 * size_t(i)] = make_float4(A.x, A.y, A.z, cap.radius);
            seg_prev[2 * size_t(i) + 1] = make_float4(B.x, B.y, B.z, 0.0f);
        }
        prev_tf = out_tf;
    }
    // m_aabbData.capsuleAABBs[tag.index] is left at the last executed substep's query box (physics_system.cpp:355-356)
    std::vector<float> last_box(6 * size_t(n));
    PRTC_CUDA(cudaMemcpyAsync(last_box.data(), c->d_last_box.ptr, last_box.size() * sizeof(float), cudaMemcpyDeviceToHost, s));
    PRTC_CUDA(cudaStreamSynchronize(s));
    for (uint32_t i = 0; i < n; ++i) {
        Box b;
        b.lo = v3(last_box[6 * size_t(i)], last_box[6 * size_t(i) + 1], last_box[6 * size_t(i) + 2]);
        b.hi = v3(last_box[6 * size_t(i) + 3], last_box[6 * size_t(i) + 4], last_box[6 * size_t(i) + 5]);
        c->capsule_aabbs[ph[i].collider] = b;
    }
    std::memcpy(tf, out_tf.data(), size_t(n) * sizeof(prtc_transform));
    std::memcpy(ph, out_ph.data(), size_t(n) * sizeof(prtc_char_physics));
    if (tr) return gather_trace(c, n, TRACE_CAP_CAND, TRACE_CAP_HITS, tr);
    return PRTC_OK;
}

} // namespace

extern "C" {

int prtc_synchronize(prtc_ctx * c) {
    if (!c) return fail(PRTC_ERR_INVALID, "prtc_synchronize: null context");
    return check_bad_colliders(c);          // synchronises; reports characters skipped by device-pointer steps
}

int prtc_step_characters_device(prtc_ctx * c, float dt, uint32_t n, prtc_transform * d_tf, prtc_char_physics * d_ph) {
    if (!c || (n && (!d_tf || !d_ph))) return fail(PRTC_ERR_INVALID, "prtc_step_characters_device: null argument");
    int rc = check_mode(c, true);
    if (rc) return rc;
    PRTC_CUDA(cudaSetDevice(c->cfg.device));
    if (c->geometry_dirty || c->tree_dirty || c->capsules_dirty) { rc = commit_impl(c); if (rc) return rc; }
    if (c->capsules.empty() && n) return fail(PRTC_ERR_INVALID, "prtc_step_characters: no capsule collider registered");
    StepParams p;
    const bool chain_was_armed = c->chain.armed;
    rc = fill_params(c, dt, n, d_tf, d_ph, p);                   // (disarms the chain: every other kind of step breaks it)
    if (rc) return rc;
    int mode = STEP_STATIC;
    if (c->chain_frames && !c->jacobi() && n) {
        // PRTC_OPT_CHAIN_FRAMES: this frame may start on the device while the previous one (same batch) is draining
        if (c->d_chain_done.cap < size_t(n) + 1) {
            PRTC_CUDA(c->d_chain_done.reserve(size_t(n) + 1));
            PRTC_CUDA(cudaMemsetAsync(c->d_chain_done.ptr, 0, c->d_chain_done.cap * sizeof(uint32_t), c->stream));
            c->chain.armed = false;
        } else c->chain.armed = chain_was_armed;
        p.chain = &c->chain;
        p.chain_done = c->d_chain_done.ptr;
        if (!c->d_chain_ring.ptr) { PRTC_CUDA(c->d_chain_ring.reserve(CHAIN_RING)); c->chain.armed = false; }
        p.chain_ring = c->d_chain_ring.ptr;
    }
    if (c->jacobi()) {
        if (!c->dyn_prepared) { rc = dyn_prepare(c, dt, n, d_tf, d_ph); if (rc) return rc; }
        // (the records are on their way to the peers: bring the per-slot leaf lists up to date meanwhile, they do not depend on them)
        if (c->dyn_ipc.on && launch_slot_refresh(p, STEP_JACOBI, c->stream)) { p.xc_fresh = 1u; c->launches += 1; }
        rc = dyn_build_grid(c, p);
        if (rc) return rc;
        c->dyn_prepared = false;
        mode = STEP_JACOBI;
    }
    c->launches += uint64_t(launch_step(p, false, mode, c->stream));
    if (mode == STEP_JACOBI) c->launches += 1;                // k_dyn_prep (the grid build counts its own)
    PRTC_CUDA(cudaGetLastError());
    return PRTC_OK;
}

namespace {

// cuStreamWriteValue32 / cuStreamWaitValue32 through the runtime's driver entry-point query: libprtc.so does not link
// libcuda, so it still loads (and fails loudly in prtc_create) on a machine without a driver.
struct StreamMemOps {
    typedef CUresult (*WriteFn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
    typedef CUresult (*WaitFn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
    WriteFn write = nullptr;
    WaitFn wait = nullptr;
    bool ok = false;
};
StreamMemOps const & stream_memops() {
    static const StreamMemOps ops = [] {
        StreamMemOps o;
        void * w = nullptr;
        void * q = nullptr;
        cudaDriverEntryPointQueryResult rw, rq;
        if (cudaGetDriverEntryPoint("cuStreamWriteValue32", &w, cudaEnableDefault, &rw) == cudaSuccess && rw == cudaDriverEntryPointSuccess &&
            cudaGetDriverEntryPoint("cuStreamWaitValue32", &q, cudaEnableDefault, &rq) == cudaSuccess && rq == cudaDriverEntryPointSuccess) {
            o.write = reinterpret_cast<StreamMemOps::WriteFn>(w);
            o.wait = reinterpret_cast<StreamMemOps::WaitFn>(q);
            o.ok = o.write && o.wait;
        }
        (void)cudaGetLastError();
        return o;
    }();
    return ops;
}

// Host-buffer step of a big static batch as ONE persistent launch fed while it runs:
//   copy-in stream   chunk k of both record arrays, then `gate = characters arrived` (stream-ordered 32-bit write)
//   compute stream   k_step_stream<GATED>: LOAD waits for its indices behind the gate, finished characters are counted
//   copy-out stream  waits (stream-ordered 32-bit wait) until chunk k's counter is full, then copies chunk k back
// so the first triangle test starts after 5.5 MB instead of 84 MB have crossed PCIe, and the last copy-out is one chunk
// long.  The pipes already wait for the context's stream (pipe_ready).
int ingest_step(prtc_ctx * c, float dt, uint32_t n, prtc_transform * tf, prtc_char_physics * ph) {
    StreamMemOps const & ops = stream_memops();
    cudaStream_t compute = c->pipe[0], in = c->pipe[1], out = c->pipe[2];
    const uint32_t chunk = 1u << (INGEST_SUB_SHIFT + INGEST_CHUNK_SHIFT);
    const uint32_t n_sub = (n + (1u << INGEST_SUB_SHIFT) - 1) >> INGEST_SUB_SHIFT;
    const uint32_t n_chunks = (n + chunk - 1) / chunk;
    PRTC_CUDA(c->d_ingest.reserve(size_t(1) + n_chunks + n_sub));
    StepParams p;
    int rc = fill_params(c, dt, n, c->d_tf.ptr, c->d_ph.ptr, p);
    if (rc) return rc;
    p.gate = c->d_ingest.ptr;
    p.done_chunk = c->d_ingest.ptr + 1;
    p.done_sub = c->d_ingest.ptr + 1 + n_chunks;
    // compute stream first: it also carries the slot refresh of the previous call, which reads the record arrays
    PRTC_CUDA(cudaMemsetAsync(c->d_ingest.ptr, 0, (size_t(1) + n_chunks + n_sub) * sizeof(uint32_t), compute));
    PRTC_CUDA(cudaEventRecord(c->ingest_reset, compute));
    PRTC_CUDA(cudaStreamWaitEvent(in, c->ingest_reset, 0));
    PRTC_CUDA(cudaStreamWaitEvent(out, c->ingest_reset, 0));
    // SUBMISSION ORDER MATTERS: everything the kernel waits for (the copies in and their gate writes) is submitted
    // before the kernel, everything that waits for the kernel (the copies out) after it.  Hardware queues are FIFO, so
    // even if the driver folds these streams onto one queue the worst case is "no overlap", never a kernel spinning on
    // a write that is queued behind it.
    // copy-in pieces: 32 Ki characters first so the kernel starts early, 64 Ki after that.  Measured at C4 (1 M characters):
    // bigger pieces finish the copy sooner (1.7 vs 2.1 ms) but open the gate in coarser steps and the kernel ends later
    // (e2e 3.08 ms at 256 Ki, 3.33 at 1 Mi vs 2.96); smaller ones pay per-operation gaps (3.27 ms at 32 Ki throughout)
    // (experiment knobs, clamped: a piece is at least 32 characters -- whole lines of both record arrays -- and the pieces never shrink)
    static const uint32_t in_first = [] { const char * e = std::getenv("PRTC_INGEST_IN_FIRST"); const long v = e ? std::atol(e) : 32768; return uint32_t(std::min(std::max(v, 32L), 1L << 24)); }();
    static const uint32_t in_max = [] { const char * e = std::getenv("PRTC_INGEST_IN_MAX"); const long v = e ? std::atol(e) : 65536; return uint32_t(std::min(std::max(v, 32L), 1L << 24)); }();
    for (uint32_t first = 0, piece = in_first & ~31u; first < n; first += piece, piece = std::max(piece, std::min(2 * piece, in_max & ~31u))) {
        const uint32_t cnt = std::min(piece, n - first);
        PRTC_CUDA(cudaMemcpyAsync(c->d_tf.ptr + first, tf + first, size_t(cnt) * sizeof(prtc_transform), cudaMemcpyHostToDevice, in));
        PRTC_CUDA(cudaMemcpyAsync(c->d_ph.ptr + first, ph + first, size_t(cnt) * sizeof(prtc_char_physics), cudaMemcpyHostToDevice, in));
        if (ops.write(in, CUdeviceptr(reinterpret_cast<uintptr_t>(p.gate)), first + cnt, CU_STREAM_WRITE_VALUE_DEFAULT) != CUDA_SUCCESS)
            return fail(PRTC_ERR_CUDA, "cuStreamWriteValue32 failed");
    }
    c->launches += uint64_t(launch_step_ingest(p, compute));
    PRTC_CUDA(cudaGetLastError());
    for (uint32_t k = 0, first = 0; first < n; first += chunk, ++k) {
        const uint32_t cnt = std::min(chunk, n - first);
        const uint32_t blocks = (cnt + (1u << INGEST_SUB_SHIFT) - 1) >> INGEST_SUB_SHIFT;
        if (ops.wait(out, CUdeviceptr(reinterpret_cast<uintptr_t>(p.done_chunk + k)), blocks, CU_STREAM_WAIT_VALUE_GEQ) != CUDA_SUCCESS)
            return fail(PRTC_ERR_CUDA, "cuStreamWaitValue32 failed");
        PRTC_CUDA(cudaMemcpyAsync(tf + first, c->d_tf.ptr + first, size_t(cnt) * sizeof(prtc_transform), cudaMemcpyDeviceToHost, out));
        PRTC_CUDA(cudaMemcpyAsync(ph + first, c->d_ph.ptr + first, size_t(cnt) * sizeof(prtc_char_physics), cudaMemcpyDeviceToHost, out));
    }
    // the compute stream may still be refreshing slots for the next frame: later work on the context's stream follows
    // it, this call only waits for the last copy-out (every character was counted after its bad-collider report)
    PRTC_CUDA(cudaEventRecord(c->ingest_reset, compute));
    PRTC_CUDA(cudaStreamWaitEvent(c->stream, c->ingest_reset, 0));
    return check_bad_colliders(c, out);
}

} // namespace

int prtc_step_characters(prtc_ctx * c, float dt, uint32_t n, prtc_transform * tf, prtc_char_physics * ph) {
    if (!c || (n && (!tf || !ph))) return fail(PRTC_ERR_INVALID, "prtc_step_characters: null argument");
    int rc = check_mode(c, false);
    if (rc) return rc;
    PRTC_CUDA(cudaSetDevice(c->cfg.device));
    if (c->capsules.empty() && n) return fail(PRTC_ERR_INVALID, "prtc_step_characters: no capsule collider registered");
    if (c->literal()) return literal_step(c, dt, n, tf, ph, nullptr);
    PRTC_CUDA(c->d_tf.reserve(n));
    PRTC_CUDA(c->d_ph.reserve(n));
    // Static pass, big batch: characters are independent, so the batch is cut into chunks and the chunks' host->device
    // copies, steps and device->host copies overlap on a few streams (PCIe is full duplex).  Results are identical
    // to one launch -- every character's frame is computed by one thread either way.
    static const uint32_t kChunk = [] { const char * e = std::getenv("PRTC_CHUNK"); const long v = e ? std::atol(e) : 65536; return uint32_t(std::min(std::max(v, 1024L), 1L << 24)); }();
    static const bool kChunkPersistent = [] { const char * e = std::getenv("PRTC_CHUNK_PERSISTENT"); return e ? std::atoi(e) != 0 : false; }();   // chunks overlap: the lock-step kernel has no per-launch drain tail (3.13 vs 3.70 ms)
    if (!c->jacobi() && n >= 4 * kChunk) {
        if (c->geometry_dirty || c->tree_dirty || c->capsules_dirty) { rc = commit_impl(c); if (rc) return rc; }
        if (!c->pipe_ready) {
            PRTC_CUDA(cudaEventCreateWithFlags(&c->pipe_ready, cudaEventDisableTiming));
            PRTC_CUDA(cudaEventCreateWithFlags(&c->ingest_reset, cudaEventDisableTiming));
            for (int k = 0; k < prtc_ctx::PIPE_STREAMS; ++k) PRTC_CUDA(cudaStreamCreateWithFlags(&c->pipe[k], cudaStreamNonBlocking));
        }
        { StepParams whole; rc = fill_params(c, dt, n, c->d_tf.ptr, c->d_ph.ptr, whole); if (rc) return rc; }   // sizes the per-slot caches for n
        PRTC_CUDA(cudaEventRecord(c->pipe_ready, c->stream));          // whatever the context's stream still has queued
        for (int k = 0; k < prtc_ctx::PIPE_STREAMS; ++k) PRTC_CUDA(cudaStreamWaitEvent(c->pipe[k], c->pipe_ready, 0));
        if (c->persistent && c->streamed_copy && stream_memops().ok) return ingest_step(c, dt, n, tf, ph);
        uint32_t k = 0;
        for (uint32_t first = 0; first < n; first += kChunk, ++k) {
            const uint32_t cnt = std::min(kChunk, n - first);
            cudaStream_t s = c->pipe[k % prtc_ctx::PIPE_STREAMS];
            PRTC_CUDA(cudaMemcpyAsync(c->d_tf.ptr + first, tf + first, size_t(cnt) * sizeof(prtc_transform), cudaMemcpyHostToDevice, s));
            PRTC_CUDA(cudaMemcpyAsync(c->d_ph.ptr + first, ph + first, size_t(cnt) * sizeof(prtc_char_physics), cudaMemcpyHostToDevice, s));
            StepParams p;
            rc = fill_params(c, dt, cnt, c->d_tf.ptr + first, c->d_ph.ptr + first, p);
            if (rc) return rc;
            p.work_flip = nullptr;                             // concurrent launches on several streams: a counter each, zeroed by a memset
            if (!kChunkPersistent) p.work_counter = nullptr;
            if (p.work_counter) p.work_counter += 1 + k % prtc_ctx::PIPE_STREAMS;     // launches on one stream are ordered
            p.xc_first = first;                               // cache slots are indexed by the character's position in the batch
            c->launches += uint64_t(launch_step(p, false, STEP_STATIC, s));
            PRTC_CUDA(cudaGetLastError());
            PRTC_CUDA(cudaMemcpyAsync(tf + first, c->d_tf.ptr + first, size_t(cnt) * sizeof(prtc_transform), cudaMemcpyDeviceToHost, s));
            PRTC_CUDA(cudaMemcpyAsync(ph + first, c->d_ph.ptr + first, size_t(cnt) * sizeof(prtc_char_physics), cudaMemcpyDeviceToHost, s));
        }
        for (int q = 0; q < prtc_ctx::PIPE_STREAMS; ++q) PRTC_CUDA(cudaStreamSynchronize(c->pipe[q]));
        return check_bad_colliders(c);
    }
    if (n > 0 && n <= kSmallBatch && !c->jacobi() && c->warp_kernel && c->staged) {
        // engine-sized static batch: stage the records in mapped pinned memory, one launch, one synchronisation
        if (c->geometry_dirty || c->tree_dirty || c->capsules_dirty) { rc = commit_impl(c); if (rc) return rc; }
        PRTC_CUDA(c->stage.reserve(128 + size_t(n) * (sizeof(prtc_transform) + sizeof(prtc_char_physics)) + 128));
        StageCursor cur{c->stage.base};
        uint32_t * flags = cur.take<uint32_t>(16);
        prtc_transform * s_tf = cur.take<prtc_transform>(n);
        prtc_char_physics * s_ph = cur.take<prtc_char_physics>(n);
        std::memcpy(s_tf, tf, size_t(n) * sizeof(prtc_transform));
        std::memcpy(s_ph, ph, size_t(n) * sizeof(prtc_char_physics));
        flags[0] = 0u; flags[1] = 0u;
        StepParams p;
        rc = fill_params(c, dt, n, s_tf, s_ph, p);
        if (rc) return rc;
        p.host_flags = flags;
        const uint32_t bell = arm_bell(c, p, flags + 2);
        c->launches += uint64_t(launch_step(p, false, STEP_STATIC, c->stream));
        PRTC_CUDA(cudaGetLastError());
        rc = wait_bell(c, flags + 2, bell);
        if (rc) return rc;
        std::memcpy(tf, s_tf, size_t(n) * sizeof(prtc_transform));
        std::memcpy(ph, s_ph, size_t(n) * sizeof(prtc_char_physics));
        return flags[1] ? check_bad_colliders(c) : PRTC_OK;
    }
    if (n) {
        PRTC_CUDA(cudaMemcpyAsync(c->d_tf.ptr, tf, size_t(n) * sizeof(prtc_transform), cudaMemcpyHostToDevice, c->stream));
        PRTC_CUDA(cudaMemcpyAsync(c->d_ph.ptr, ph, size_t(n) * sizeof(prtc_char_physics), cudaMemcpyHostToDevice, c->stream));
    }
    rc = prtc_step_characters_device(c, dt, n, c->d_tf.ptr, c->d_ph.ptr);
    if (rc) return rc;
    if (n) {
        PRTC_CUDA(cudaMemcpyAsync(tf, c->d_tf.ptr, size_t(n) * sizeof(prtc_transform), cudaMemcpyDeviceToHost, c->stream));
        PRTC_CUDA(cudaMemcpyAsync(ph, c->d_ph.ptr, size_t(n) * sizeof(prtc_char_physics), cudaMemcpyDeviceToHost, c->stream));
    }
    return check_bad_colliders(c);
}

int prtc_step_characters_traced(prtc_ctx * c, float dt, uint32_t n, prtc_transform * tf, prtc_char_physics * ph, prtc_trace * tr) {
    if (!c || !tr || (n && (!tf || !ph))) return fail(PRTC_ERR_INVALID, "prtc_step_characters_traced: null argument");
    int rc = check_mode(c, false);
    if (rc) return rc;
    PRTC_CUDA(cudaSetDevice(c->cfg.device));
    if (c->capsules.empty() && n) return fail(PRTC_ERR_INVALID, "prtc_step_characters: no capsule collider registered");
    if (c->literal()) return literal_step(c, dt, n, tf, ph, tr);
    if (c->geometry_dirty || c->tree_dirty || c->capsules_dirty) { rc = commit_impl(c); if (rc) return rc; }
    cudaStream_t s = c->stream;
    PRTC_CUDA(c->d_tf.reserve(n));
    PRTC_CUDA(c->d_ph.reserve(n));
    tr->n_queries = tr->n_mesh_ids = tr->n_caps_ids = tr->n_hits = 0;
    if (n == 0) { if (tr->q_mesh_off) tr->q_mesh_off[0] = 0; if (tr->q_caps_off) tr->q_caps_off[0] = 0; return PRTC_OK; }
    PRTC_CUDA(cudaMemcpyAsync(c->d_tf.ptr, tf, size_t(n) * sizeof(prtc_transform), cudaMemcpyHostToDevice, s));
    PRTC_CUDA(cudaMemcpyAsync(c->d_ph.ptr, ph, size_t(n) * sizeof(prtc_char_physics), cudaMemcpyHostToDevice, s));
    StepParams p;
    rc = fill_params(c, dt, n, c->d_tf.ptr, c->d_ph.ptr, p);
    if (rc) return rc;
    rc = prepare_trace(c, n, p);
    if (rc) return rc;
    int mode = STEP_STATIC;
    if (c->jacobi()) {
        if (!c->dyn_prepared) { rc = dyn_prepare(c, dt, n, c->d_tf.ptr, c->d_ph.ptr); if (rc) return rc; }
        rc = dyn_build_grid(c, p);
        if (rc) return rc;
        c->dyn_prepared = false;
        mode = STEP_JACOBI;
    }
    c->launches += uint64_t(launch_step(p, true, mode, s));
    PRTC_CUDA(cudaGetLastError());
    PRTC_CUDA(cudaMemcpyAsync(tf, c->d_tf.ptr, size_t(n) * sizeof(prtc_transform), cudaMemcpyDeviceToHost, s));
    PRTC_CUDA(cudaMemcpyAsync(ph, c->d_ph.ptr, size_t(n) * sizeof(prtc_char_physics), cudaMemcpyDeviceToHost, s));
    return gather_trace(c, n, TRACE_CAP_CAND, TRACE_CAP_HITS, tr);
}

int prtc_raycast(prtc_ctx * c, uint32_t n, const float * origins, const float * directions, const float * max_distances,
                 float * hit_xyz, uint8_t * hit_mask) {
    if (!c || (n && (!origins || !directions || !max_distances || !hit_xyz || !hit_mask)))
        return fail(PRTC_ERR_INVALID, "prtc_raycast: null argument");
    PRTC_CUDA(cudaSetDevice(c->cfg.device));
    if (c->geometry_dirty || c->tree_dirty || c->capsules_dirty) { int rc = commit_impl(c); if (rc) return rc; }
    if (n == 0) return PRTC_OK;
    cudaStream_t s = c->stream;
    if (n <= kSmallBatch && c->staged) {
        // the engine casts four rays per frame, one call each (scene.cpp:248-255): rays and answers live in mapped pinned
        // memory, the call is one launch and one synchronisation instead of four uploads, a launch and two downloads
        PRTC_CUDA(c->stage.reserve(size_t(n) * 11 * sizeof(float) + 8 * 64));
        StageCursor cur{c->stage.base};
        float * s_o = cur.take<float>(3 * size_t(n)), * s_d = cur.take<float>(3 * size_t(n)), * s_m = cur.take<float>(n);
        float * s_h = cur.take<float>(3 * size_t(n));
        unsigned char * s_mask = cur.take<unsigned char>(n);
        std::memcpy(s_o, origins, size_t(n) * 3 * sizeof(float));
        std::memcpy(s_d, directions, size_t(n) * 3 * sizeof(float));
        std::memcpy(s_m, max_distances, size_t(n) * sizeof(float));
        std::memset(s_h, 0, size_t(n) * 3 * sizeof(float));
        uint32_t * bell_word = cur.take<uint32_t>(16);
        StepParams bp;
        std::memset(&bp, 0, sizeof(bp));
        const uint32_t bell = raycast_rings_doorbell(uint32_t(c->flat.nodes.size()), n) ? arm_bell(c, bp, bell_word) : 0u;
        launch_raycast(c->d_nodes.ptr, uint32_t(c->flat.nodes.size()), c->d_tris.ptr, n, s_o, s_d, s_m, s_h, s_mask, s, bp.done_flag, bp.done_count, bp.done_value);
        PRTC_CUDA(cudaGetLastError());
        { const int wrc = wait_bell(c, bell_word, bell); if (wrc) return wrc; }
        std::memcpy(hit_mask, s_mask, n);
        for (uint32_t i = 0; i < n; ++i)
            if (hit_mask[i]) { hit_xyz[3 * i] = s_h[3 * size_t(i)]; hit_xyz[3 * i + 1] = s_h[3 * size_t(i) + 1]; hit_xyz[3 * i + 2] = s_h[3 * size_t(i) + 2]; }
        return PRTC_OK;
    }
    PRTC_CUDA(c->d_scratch_f.reserve(size_t(n) * 10));
    PRTC_CUDA(c->d_scratch_u.reserve((size_t(n) + 3) / 4 + 1));
    float * d_o = c->d_scratch_f.ptr, * d_d = d_o + 3 * size_t(n), * d_m = d_d + 3 * size_t(n), * d_h = d_m + n;
    unsigned char * d_mask = reinterpret_cast<unsigned char *>(c->d_scratch_u.ptr);
    PRTC_CUDA(cudaMemcpyAsync(d_o, origins, size_t(n) * 3 * sizeof(float), cudaMemcpyHostToDevice, s));
    PRTC_CUDA(cudaMemcpyAsync(d_d, directions, size_t(n) * 3 * sizeof(float), cudaMemcpyHostToDevice, s));
    PRTC_CUDA(cudaMemcpyAsync(d_m, max_distances, size_t(n) * sizeof(float), cudaMemcpyHostToDevice, s));
    PRTC_CUDA(cudaMemsetAsync(d_h, 0, size_t(n) * 3 * sizeof(float), s));
    launch_raycast(c->d_nodes.ptr, uint32_t(c->flat.nodes.size()), c->d_tris.ptr, n, d_o, d_d, d_m, d_h, d_mask, s);
    PRTC_CUDA(cudaGetLastError());
    std::vector<float> h(size_t(n) * 3);
    PRTC_CUDA(cudaMemcpyAsync(h.data(), d_h, h.size() * sizeof(float), cudaMemcpyDeviceToHost, s));
    PRTC_CUDA(cudaMemcpyAsync(hit_mask, d_mask, n, cudaMemcpyDeviceToHost, s));
    PRTC_CUDA(cudaStreamSynchronize(s));
    for (uint32_t i = 0; i < n; ++i)
        if (hit_mask[i]) { hit_xyz[3 * i] = h[3 * size_t(i)]; hit_xyz[3 * i + 1] = h[3 * size_t(i) + 1]; hit_xyz[3 * i + 2] = h[3 * size_t(i) + 2]; }
    return PRTC_OK;
}

int prtc_sweep_ellipsoids(prtc_ctx * c, uint32_t n, const float * centers, const float * radii, const float * velocities,
                          uint32_t iterations, float very_close, float * out_pos, float * out_vel, int32_t * hit_tri,
                          float * hit_toi, float * hit_normal) {
    if (!c || (n && (!centers || !radii || !velocities || !out_pos || !out_vel || !hit_tri || !hit_toi || !hit_normal)) || iterations == 0)
        return fail(PRTC_ERR_INVALID, "prtc_sweep_ellipsoids: null argument or zero iterations");
    if (c->literal()) return fail(PRTC_ERR_INVALID, "prtc_sweep_ellipsoids needs a PRTC_ORDER_CANONICAL context (mesh colliders only in the tree)");
    for (uint32_t i = 0; i < 3 * n; ++i)
        if (!(radii[i] > 0.0f)) return fail(PRTC_ERR_INVALID, "prtc_sweep_ellipsoids: radii must be positive");
    PRTC_CUDA(cudaSetDevice(c->cfg.device));
    if (c->geometry_dirty || c->tree_dirty || c->capsules_dirty) { int rc = commit_impl(c); if (rc) return rc; }
    if (n == 0) return PRTC_OK;
    cudaStream_t s = c->stream;
    const size_t hits = size_t(n) * iterations;
    PRTC_CUDA(c->d_scratch_f.reserve(size_t(n) * 15 + hits * 4));
    PRTC_CUDA(c->d_scratch_u.reserve(hits));
    float * d_c = c->d_scratch_f.ptr, * d_r = d_c + 3 * size_t(n), * d_v = d_r + 3 * size_t(n), * d_p = d_v + 3 * size_t(n), * d_ov = d_p + 3 * size_t(n);
    float * d_toi = d_ov + 3 * size_t(n), * d_nrm = d_toi + hits;
    PRTC_CUDA(cudaMemcpyAsync(d_c, centers, size_t(n) * 3 * sizeof(float), cudaMemcpyHostToDevice, s));
    PRTC_CUDA(cudaMemcpyAsync(d_r, radii, size_t(n) * 3 * sizeof(float), cudaMemcpyHostToDevice, s));
    PRTC_CUDA(cudaMemcpyAsync(d_v, velocities, size_t(n) * 3 * sizeof(float), cudaMemcpyHostToDevice, s));
    SweepParams p;
    p.nodes = c->d_nodes.ptr; p.n_nodes = uint32_t(c->flat.nodes.size());
    p.tris = c->d_tris.ptr; p.tri_src = c->d_tri_src.ptr;
    p.n = n; p.iterations = iterations; p.very_close = very_close;
    p.centers = d_c; p.radii = d_r; p.velocities = d_v; p.out_pos = d_p; p.out_vel = d_ov;
    p.hit_tri = reinterpret_cast<int32_t *>(c->d_scratch_u.ptr); p.hit_toi = d_toi; p.hit_normal = d_nrm;
    p.counters = c->d_counters.ptr;
    launch_sweep(p, s);
    c->launches += 1;
    PRTC_CUDA(cudaGetLastError());
    PRTC_CUDA(cudaMemcpyAsync(out_pos, d_p, size_t(n) * 3 * sizeof(float), cudaMemcpyDeviceToHost, s));
    PRTC_CUDA(cudaMemcpyAsync(out_vel, d_ov, size_t(n) * 3 * sizeof(float), cudaMemcpyDeviceToHost, s));
    PRTC_CUDA(cudaMemcpyAsync(hit_tri, p.hit_tri, hits * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    PRTC_CUDA(cudaMemcpyAsync(hit_toi, d_toi, hits * sizeof(float), cudaMemcpyDeviceToHost, s));
    PRTC_CUDA(cudaMemcpyAsync(hit_normal, d_nrm, hits * 3 * sizeof(float), cudaMemcpyDeviceToHost, s));
    PRTC_CUDA(cudaStreamSynchronize(s));
    return PRTC_OK;
}

} // extern "C"

namespace prt {
namespace host {

int dyn_build_grid_for(prtc_ctx * c, StepParams & p) { return dyn_build_grid(c, p); }

int resident_begin_impl(prtc_ctx * c, uint32_t n, const prtc_transform * tf, const prtc_char_physics * ph, uint32_t * io_flags, float * io_move,
                        float * io_result) {
    if (!c || (n && (!tf || !ph))) return fail(PRTC_ERR_INVALID, "prtc_resident_begin: null argument");
    int rc = check_mode(c, true);                                  // LITERAL order maintains the tree on the host from host records
    if (rc) return rc;
    PRTC_CUDA(cudaSetDevice(c->cfg.device));
    if (c->capsules.empty() && n) return fail(PRTC_ERR_INVALID, "prtc_resident_begin: no capsule collider registered");
    if (c->geometry_dirty || c->tree_dirty || c->capsules_dirty) { rc = commit_impl(c); if (rc) return rc; }
    PRTC_CUDA(c->d_tf.reserve(n));
    PRTC_CUDA(c->d_ph.reserve(n));
    if (io_flags) {
        c->resident_flags = io_flags; c->resident_move = io_move; c->resident_result = io_result;
    } else {
        PRTC_CUDA(c->resident_io.reserve(64 + size_t(n) * 7 * sizeof(float) + 64));
        c->resident_flags = reinterpret_cast<uint32_t *>(c->resident_io.base);
        c->resident_move = reinterpret_cast<float *>(c->resident_io.base + 64);
        c->resident_result = c->resident_move + 3 * size_t(n) + ((4 - (3 * size_t(n)) % 4) % 4);       // float4-aligned
    }
    c->resident_flags[0] = 0u; c->resident_flags[1] = 0u; c->resident_flags[2] = 0u;
    c->resident_frame = 0;
    c->resident_prefresh = 0u;
    if (n) {
        PRTC_CUDA(cudaMemcpyAsync(c->d_tf.ptr, tf, size_t(n) * sizeof(prtc_transform), cudaMemcpyHostToDevice, c->stream));
        PRTC_CUDA(cudaMemcpyAsync(c->d_ph.ptr, ph, size_t(n) * sizeof(prtc_char_physics), cudaMemcpyHostToDevice, c->stream));
        PRTC_CUDA(cudaStreamSynchronize(c->stream));
    }
    for (uint32_t i = 0; i < n; ++i) {
        for (int k = 0; k < 3; ++k) c->resident_move[3 * size_t(i) + k] = ph[i].movement[k];
        c->resident_result[4 * size_t(i)] = tf[i].position[0]; c->resident_result[4 * size_t(i) + 1] = tf[i].position[1];
        c->resident_result[4 * size_t(i) + 2] = tf[i].position[2]; c->resident_result[4 * size_t(i) + 3] = ph[i].grounded ? 1.0f : 0.0f;
    }
    c->resident = true;
    c->resident_n = n;
    return PRTC_OK;
}

int resident_launch(prtc_ctx * c, float dt) {
    if (!c || !c->resident) return fail(PRTC_ERR_INVALID, "prtc_resident_step: call prtc_resident_begin first");
    int rc = check_mode(c, true);
    if (rc) return rc;
    PRTC_CUDA(cudaSetDevice(c->cfg.device));
    if (c->geometry_dirty || c->tree_dirty || c->capsules_dirty) { rc = commit_impl(c); if (rc) return rc; }
    const uint32_t n = c->resident_n;
    c->resident_pending = 0u;
    if (n == 0) return PRTC_OK;
    StepParams p;
    rc = fill_params(c, dt, n, c->d_tf.ptr, c->d_ph.ptr, p);
    if (rc) return rc;
    p.io_move = c->resident_move;
    p.io_result = reinterpret_cast<float4 *>(c->resident_result);
    p.host_flags = c->resident_flags;
    int mode = STEP_STATIC;
    if (c->jacobi()) {
        // the records of the dynamic pass are built from the state BEFORE this frame's movement input is applied: movement
        // only enters the step (physics_system.cpp:281-282), not the predicted box (:263-273)
        if (!c->dyn_prepared) { rc = dyn_prepare(c, dt, n, c->d_tf.ptr, c->d_ph.ptr); if (rc) return rc; }
        rc = dyn_build_grid_for(c, p);
        if (rc) return rc;
        c->dyn_prepared = false;
        mode = STEP_JACOBI;
    }
    // completion doorbell: the last warp of the step's last kernel writes the frame number to a host-mapped word; spinning on
    // it returns a few microseconds after the kernel ends, where cudaStreamSynchronize's wake-up costs 10-20
    if (c->doorbell) {
        p.done_flag = c->resident_flags + 2;
        p.done_count = c->d_work.ptr + 6;
        p.done_value = ++c->resident_frame;
        if (p.done_value == 0u) p.done_value = ++c->resident_frame;
        c->resident_pending = p.done_value;
    }
    // the per-slot leaf lists for THIS frame were refreshed behind the previous one (below) unless something intervened
    if (mode == STEP_STATIC && c->resident_prefresh == c->tree_version && c->resident_prefresh != 0u) p.xc_fresh = 1u;
    c->launches += uint64_t(launch_step(p, false, mode, c->stream));
    PRTC_CUDA(cudaGetLastError());
    // ... and the lists for the NEXT frame are refreshed right behind this one, from the state it leaves (with this frame's
    // movement input as the guess for the next): the doorbell has rung by then, the host is already back with the caller.
    // A wrong guess costs a tree walk inside the step, never a result (the step checks "query box inside region" per substep).
    c->resident_prefresh = 0u;
    if (mode == STEP_STATIC) {
        StepParams r = p;
        r.done_flag = nullptr; r.io_move = nullptr; r.io_result = nullptr;
        if (launch_slot_refresh(r, STEP_STATIC, c->stream)) { c->resident_prefresh = c->tree_version; c->launches += 1; }
        PRTC_CUDA(cudaGetLastError());
    }
    return PRTC_OK;
}

int resident_wait(prtc_ctx * c) {
    if (!c || !c->resident) return fail(PRTC_ERR_INVALID, "prtc_resident_step: call prtc_resident_begin first");
    if (c->resident_n == 0) return PRTC_OK;
    PRTC_CUDA(cudaSetDevice(c->cfg.device));
    if (c->resident_pending) {
        volatile uint32_t * flag = c->resident_flags + 2;
        uint32_t spins = 0;
        while (*flag != c->resident_pending) {
            if ((++spins & 0xFFFu) == 0u) {                        // now and then: did the stream die, or finish without ringing?
                const cudaError_t q = cudaStreamQuery(c->stream);
                if (q == cudaSuccess) break;
                if (q != cudaErrorNotReady) return fail(PRTC_ERR_CUDA, std::string("prtc_resident_step: ") + cudaGetErrorString(q));
            }
        }
        c->resident_pending = 0u;
    } else {
        PRTC_CUDA(cudaStreamSynchronize(c->stream));
    }
    if (c->resident_flags[1]) {
        c->resident_flags[1] = 0u;
        return check_bad_colliders(c);
    }
    return PRTC_OK;
}

} // namespace host
} // namespace prt

extern "C" {

// ---- resident state ------------------------------------------------------------------------------------------------
// What the engine's caller does around updateCharacters is gather -> call -> scatter of positions (character_system.cpp:
// 55-78); the records themselves need not cross the bus every frame.  With resident state they stay on the device, the
// caller writes this frame's movementVector into a mapped pinned array the kernel reads in place, and position + isGrounded
// come back through a second mapped array the kernel writes in place: 12 bytes in and 16 bytes out per character instead
// of 84 + 84, no separate copies, one launch and one synchronisation per frame.
int prtc_resident_begin(prtc_ctx * c, uint32_t n, const prtc_transform * tf, const prtc_char_physics * ph) {
    return resident_begin_impl(c, n, tf, ph, nullptr, nullptr, nullptr);
}

int prtc_resident_io(prtc_ctx * c, float ** movement, float ** result) {
    if (!c || !c->resident) return fail(PRTC_ERR_INVALID, "prtc_resident_io: call prtc_resident_begin first");
    if (movement) *movement = c->resident_move;
    if (result) *result = c->resident_result;
    return PRTC_OK;
}

int prtc_resident_step(prtc_ctx * c, float dt) {
    const int rc = resident_launch(c, dt);
    return rc ? rc : resident_wait(c);
}

int prtc_resident_edit(prtc_ctx * c, uint32_t k, const uint32_t * indices, const prtc_transform * tf, const prtc_char_physics * ph) {
    if (!c || !c->resident || (k && !indices) || (k && !tf && !ph)) return fail(PRTC_ERR_INVALID, "prtc_resident_edit: null argument or no resident state");
    PRTC_CUDA(cudaSetDevice(c->cfg.device));
    for (uint32_t j = 0; j < k; ++j) {
        const uint32_t i = indices[j];
        if (i >= c->resident_n) return fail(PRTC_ERR_INVALID, "prtc_resident_edit: index out of range");
        if (tf) { PRTC_CUDA(cudaMemcpyAsync(c->d_tf.ptr + i, tf + j, sizeof(prtc_transform), cudaMemcpyHostToDevice, c->stream)); c->resident_prefresh = 0u; }   // (moved: refresh the leaf lists in front of the next frame)
        if (ph) {
            PRTC_CUDA(cudaMemcpyAsync(c->d_ph.ptr + i, ph + j, sizeof(prtc_char_physics), cudaMemcpyHostToDevice, c->stream));
            for (int d = 0; d < 3; ++d) c->resident_move[3 * size_t(i) + d] = ph[j].movement[d];
        }
    }
    PRTC_CUDA(cudaStreamSynchronize(c->stream));
    return PRTC_OK;
}

int prtc_resident_fetch(prtc_ctx * c, prtc_transform * tf, prtc_char_physics * ph) {
    if (!c || !c->resident) return fail(PRTC_ERR_INVALID, "prtc_resident_fetch: no resident state");
    PRTC_CUDA(cudaSetDevice(c->cfg.device));
    const uint32_t n = c->resident_n;
    if (n && tf) PRTC_CUDA(cudaMemcpyAsync(tf, c->d_tf.ptr, size_t(n) * sizeof(prtc_transform), cudaMemcpyDeviceToHost, c->stream));
    if (n && ph) PRTC_CUDA(cudaMemcpyAsync(ph, c->d_ph.ptr, size_t(n) * sizeof(prtc_char_physics), cudaMemcpyDeviceToHost, c->stream));
    PRTC_CUDA(cudaStreamSynchronize(c->stream));
    return PRTC_OK;
}

int prtc_resident_end(prtc_ctx * c) {
    if (!c) return fail(PRTC_ERR_INVALID, "prtc_resident_end: null context");
    c->resident = false;
    c->resident_n = 0;
    return PRTC_OK;
}

int prtc_dyn_configure(prtc_ctx * c, uint32_t first_global, uint32_t n_global) {
    if (!c) return fail(PRTC_ERR_INVALID, "prtc_dyn_configure: null context");
    if (c->dyn_initialized && (n_global != c->dyn_global)) return fail(PRTC_ERR_INVALID, "prtc_dyn_configure: the global count is fixed after the first frame");
    c->dyn_first = first_global;
    c->dyn_global = n_global;
    c->dyn_configured = true;
    return PRTC_OK;
}

int prtc_dyn_prepare(prtc_ctx * c, float dt, uint32_t n, const prtc_transform * d_tf, const prtc_char_physics * d_ph) {
    if (c && c->resident && !d_tf && !d_ph) { d_tf = c->d_tf.ptr; d_ph = c->d_ph.ptr; n = c->resident_n; }   // resident state: its own arrays
    if (!c || (n && (!d_tf || !d_ph))) return fail(PRTC_ERR_INVALID, "prtc_dyn_prepare: null argument");
    if (!c->jacobi()) return fail(PRTC_ERR_INVALID, "prtc_dyn_prepare: context is not in PRTC_ORDER_CANONICAL + PRTC_DYN_JACOBI");
    PRTC_CUDA(cudaSetDevice(c->cfg.device));
    if (c->geometry_dirty || c->tree_dirty || c->capsules_dirty) { int rc = commit_impl(c); if (rc) return rc; }
    return dyn_prepare(c, dt, n, d_tf, d_ph);
}

int prtc_dyn_ipc_export(prtc_ctx * c, uint32_t rank, uint32_t world, void * handle) {
    if (!c || !handle || world < 2 || world > 16 || rank >= world) return fail(PRTC_ERR_INVALID, "prtc_dyn_ipc_export: bad argument (2..16 ranks)");
    if (!c->jacobi() || !c->dyn_configured) return fail(PRTC_ERR_INVALID, "prtc_dyn_ipc_export: needs PRTC_DYN_JACOBI and prtc_dyn_configure first");
    if (c->dyn_initialized) return fail(PRTC_ERR_INVALID, "prtc_dyn_ipc_export: call it before the first frame");
    PRTC_CUDA(cudaSetDevice(c->cfg.device));
    prtc_ctx::DynIpc & x = c->dyn_ipc;
    x.rank = rank; x.world = world; x.frame = 0;
    int rc = dyn_reserve(c, 0);
    if (rc) return rc;
    PRTC_CUDA(x.flags.reserve(32));
    PRTC_CUDA(cudaMemsetAsync(x.flags.ptr, 0, 32 * sizeof(uint32_t), c->stream));
    PRTC_CUDA(cudaStreamSynchronize(c->stream));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle layout");
    cudaIpcMemHandle_t h[2];
    PRTC_CUDA(cudaIpcGetMemHandle(&h[0], c->d_dyn_records.ptr));
    PRTC_CUDA(cudaIpcGetMemHandle(&h[1], x.flags.ptr));
    std::memcpy(handle, h, sizeof(h));
    return PRTC_OK;
}

int prtc_dyn_ipc_import(prtc_ctx * c, const void * handles) {
    if (!c) return fail(PRTC_ERR_INVALID, "prtc_dyn_ipc_import: null context");
    prtc_ctx::DynIpc & x = c->dyn_ipc;
    if (!handles) {                                                // null: give the exchange up again (some rank could not map its peers)
        for (void * q : x.opened) cudaIpcCloseMemHandle(q);
        x.opened.clear();
        x.n_peers = 0;
        x.on = false;
        (void)cudaGetLastError();
        return PRTC_OK;
    }
    if (!x.world || !x.flags.ptr) return fail(PRTC_ERR_INVALID, "prtc_dyn_ipc_import: call prtc_dyn_ipc_export first");
    PRTC_CUDA(cudaSetDevice(c->cfg.device));
    const unsigned char * hb = static_cast<const unsigned char *>(handles);
    x.n_peers = 0;
    for (uint32_t r = 0; r < x.world; ++r) {
        if (r == x.rank) continue;
        cudaIpcMemHandle_t h[2];
        std::memcpy(h, hb + size_t(r) * sizeof(h), sizeof(h));
        void * rec = nullptr, * flg = nullptr;
        PRTC_CUDA(cudaIpcOpenMemHandle(&rec, h[0], cudaIpcMemLazyEnablePeerAccess));
        x.opened.push_back(rec);
        PRTC_CUDA(cudaIpcOpenMemHandle(&flg, h[1], cudaIpcMemLazyEnablePeerAccess));
        x.opened.push_back(flg);
        x.peer_records[x.n_peers] = static_cast<float4 *>(rec);
        x.peer_flags[x.n_peers] = static_cast<uint32_t *>(flg);
        ++x.n_peers;
    }
    if (!x.push_stream) PRTC_CUDA(cudaStreamCreateWithFlags(&x.push_stream, cudaStreamNonBlocking));
    if (!x.prep_done) PRTC_CUDA(cudaEventCreateWithFlags(&x.prep_done, cudaEventDisableTiming));
    x.on = true;
    return PRTC_OK;
}

int prtc_dyn_buffer(prtc_ctx * c, void ** d_records, size_t * record_bytes, uint32_t * first_global, uint32_t * n_global) {
    if (!c || !d_records) return fail(PRTC_ERR_INVALID, "prtc_dyn_buffer: null argument");
    if (!c->d_dyn_records.ptr) return fail(PRTC_ERR_INVALID, "prtc_dyn_buffer: call prtc_dyn_prepare first");
    *d_records = c->d_dyn_records.ptr + (c->dyn_ipc.on ? (c->dyn_ipc.frame & 1u) * size_t(c->dyn_global) * 4 : 0);
    if (record_bytes) *record_bytes = 4 * sizeof(float4);
    if (first_global) *first_global = c->dyn_first;
    if (n_global) *n_global = c->dyn_global;
    return PRTC_OK;
}


} // extern "C"
