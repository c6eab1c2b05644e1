// prtc.cu -- host side of the C-ABI declared in include/prtc.h: collider registry, commit (vertex transform on
// the device, reference-algorithm tree build on the host, flatten, upload) and the per-frame step launches.
// There is no CPU implementation of the compute path in this library: without a compute-capability-10 device
// prtc_create fails and nothing else can be called.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <cstdio>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/prtc.h"
#include "bvh_builder.h"
#include "kernels.cuh"

using namespace prt;

namespace {

thread_local std::string g_last_error;

int fail(int code, const std::string & msg) {
    g_last_error = msg;
    return code;
}

#define PRTC_CUDA(expr)                                                                                   \
    do {                                                                                                  \
        cudaError_t _e = (expr);                                                                          \
        if (_e != cudaSuccess)                                                                            \
            return fail(PRTC_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));               \
    } while (0)

template <class T>
struct DevBuf {
    T * ptr = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (ptr) cudaFree(ptr);
        ptr = nullptr; cap = 0;
        size_t want = std::max<size_t>(n, 16);
        cudaError_t e = cudaMalloc(reinterpret_cast<void **>(&ptr), want * sizeof(T));
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (ptr) cudaFree(ptr); ptr = nullptr; cap = 0; }
};

// Pinned host memory the device reads and writes in place (unified addressing: one pointer for both sides).  Engine-sized
// host-buffer steps stage their records here instead of paying two uploads, two downloads and a counter read-back.
struct HostStage {
    unsigned char * base = nullptr;
    size_t bytes = 0;
    cudaError_t reserve(size_t need) {
        if (need <= bytes) return cudaSuccess;
        if (base) cudaFreeHost(base);
        base = nullptr; bytes = 0;
        const size_t want = std::max<size_t>(need, size_t(64) << 10);
        cudaError_t e = cudaHostAlloc(reinterpret_cast<void **>(&base), want, cudaHostAllocMapped);
        if (e == cudaSuccess) bytes = want;
        return e;
    }
    void release() { if (base) cudaFreeHost(base); base = nullptr; bytes = 0; }
};
// carves 64-byte aligned arrays out of a HostStage
struct StageCursor {
    unsigned char * p;
    template <typename T> T * take(size_t n) { T * r = reinterpret_cast<T *>(p); p += (n * sizeof(T) + 63) & ~size_t(63); return r; }
};
constexpr uint32_t kSmallBatch = 256;       // the engine's entity table holds 100 (entity.h:15)

struct MeshHost { uint32_t first_vertex, n_vertices, model; int32_t tree_index; };
struct ModelHost { uint32_t first_mesh, n_meshes; prtc_transform transform; };

Affine affine_from(prtc_transform const & t) {
    Quat q; q.x = t.rotation[0]; q.y = t.rotation[1]; q.z = t.rotation[2]; q.w = t.rotation[3];
    return affine_of(v3(t.position[0], t.position[1], t.position[2]), q, v3(t.scale[0], t.scale[1], t.scale[2]));
}

} // namespace

struct prtc_ctx {
    prtc_config cfg;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    // host-buffer steps of large batches are cut into chunks pipelined over these streams (copy-in, step, copy-out)
    static constexpr int PIPE_STREAMS = 3;
    cudaStream_t pipe[PIPE_STREAMS] = {nullptr, nullptr, nullptr};
    cudaEvent_t pipe_ready = nullptr;

    // host registry
    std::vector<float> raw;                 // model-space xyz of every registered model, back to back
    std::vector<MeshHost> meshes;
    std::vector<ModelHost> models;
    std::vector<CapsuleDev> capsules;
    RefTree tree;
    FlatTree flat;
    std::vector<uint32_t> tri_src;          // per leaf-ordered triangle: index of its first vertex in `raw`/`world`
    uint32_t meshes_in_tree = 0;            // meshes that have a leaf in `tree`
    uint32_t meshes_on_device = 0;          // meshes whose world cache / AABB the device has computed
    // LITERAL order: the reference's unified tree also holds one leaf per capsule (physics_system.cpp:36-39)
    std::vector<int32_t> capsule_tree_index;
    std::vector<Box> capsule_aabbs;         // m_aabbData.capsuleAABBs
    bool geometry_dirty = false;            // raw / meshes / xforms changed since the last upload
    bool tree_dirty = false;                // tree changed since the last flatten
    bool leaf_ids_stale = true;             // d_node_leaf_id lags behind flat.node_leaf_id (upload_leaf_ids)
    bool tri_dirty = true;                  // triangle records (layout or world positions) must be rebuilt
    std::vector<uint32_t> mesh_tri_first;   // LITERAL: first triangle record of every mesh (registration order)
    bool capsules_dirty = false;
    float world_max_abs = 0.0f;             // largest |coordinate| of any mesh AABB (margin of the pre-cull)
    int frame_cache = 1;                    // PRTC_OPT_FRAME_CACHE
    int persistent = 1;                     // PRTC_OPT_PERSISTENT
    uint32_t tune_hi = 0;                   // PRTC_TUNE: batch thresholds of k_step_stream (bits 8-15 load, 16-23 substep), experiments
    int sm_count = 148;
    int quick_reject = 1;                   // PRTC_OPT_QUICK_REJECT
    std::vector<uint32_t> moved_models;     // models whose transform changed (prtc_update_models)

    // device
    DevBuf<float> d_raw, d_world, d_aabbs;
    DevBuf<MeshDev> d_meshes;
    DevBuf<ModelXform> d_xforms;
    DevBuf<float4> d_nodes, d_tris;
    DevBuf<uint32_t> d_node_leaf_id, d_tri_src;
    DevBuf<CapsuleDev> d_capsules;
    DevBuf<unsigned long long> d_counters;
    // cross-frame leaf cache of the persistent kernel (PRTC_OPT_CROSS_FRAME_CACHE)
    DevBuf<uint32_t> d_xc_version, d_xc_count, d_xc_leaf;
    DevBuf<float> d_xc_region;
    uint32_t xc_capacity = 0;
    uint32_t tree_version = 1;
    uint64_t launches = 0;                  // kernels launched by step calls since the last counter reset
    int cross_frame = 1;
    int streamed_copy = 1;                  // PRTC_OPT_STREAMED_COPY: big host-buffer batches feed one persistent launch while it runs
    int warp_kernel = 1;                    // small batches: one warp per character (k_step_warp); 0 = thinned lock-step kernel
    float xc_margin = 0.5f;
    DevBuf<unsigned> d_work;                // work counters of the persistent kernel, one per concurrently running launch
    HostStage stage;                        // small host-buffer batches: records staged in mapped pinned memory
    int staged = 1;                         // PRTC_STAGED=0: small batches take the copy path too (A/B)
    DevBuf<uint32_t> d_ingest;              // streamed host-buffer step: [0] arrival gate, then per-chunk and per-block completion counters
    cudaEvent_t ingest_reset = nullptr;
    DevBuf<prtc_transform> d_tf;
    DevBuf<prtc_char_physics> d_ph;
    // trace scratch
    DevBuf<TraceQuery> d_tq;
    DevBuf<uint32_t> d_tmesh, d_tcaps;
    DevBuf<TraceHit> d_thits;
    DevBuf<float> d_scratch_f;
    DevBuf<uint32_t> d_scratch_u;
    // LITERAL order
    DevBuf<uint32_t> d_cap2char;
    DevBuf<float4> d_seg_start, d_seg_prev;
    DevBuf<float> d_last_box;
    // CANONICAL order, JACOBI dynamic pass
    DevBuf<float4> d_dyn_records;
    DevBuf<uint32_t> d_dyn_count, d_dyn_start, d_dyn_items;
    DevBuf<unsigned char> d_cub_temp;
    uint32_t dyn_first = 0, dyn_global = 0;  // this context's slice [dyn_first, dyn_first + n) of dyn_global characters
    bool dyn_configured = false;             // prtc_dyn_configure called (multi-rank); otherwise the slice is the whole batch
    bool dyn_initialized = false;            // stored boxes exist (first prepare creates them at the identity pose)
    bool dyn_prepared = false;               // prtc_dyn_prepare ran for the coming step (records may have been exchanged since)
    float world_lo[3] = {0, 0, 0}, world_hi[3] = {0, 0, 0};
    bool literal() const { return cfg.order_mode == PRTC_ORDER_LITERAL; }
    bool jacobi() const { return cfg.order_mode == PRTC_ORDER_CANONICAL && cfg.dyn_mode == PRTC_DYN_JACOBI; }
};

namespace {

int commit_impl(prtc_ctx * c) {
    cudaStream_t s = c->stream;
    const uint32_t n_meshes = uint32_t(c->meshes.size());
    const uint32_t n_vertices = uint32_t(c->raw.size() / 3);

    if (c->capsules_dirty) {
        PRTC_CUDA(c->d_capsules.reserve(c->capsules.size()));
        if (!c->capsules.empty())
            PRTC_CUDA(cudaMemcpyAsync(c->d_capsules.ptr, c->capsules.data(), c->capsules.size() * sizeof(CapsuleDev),
                                      cudaMemcpyHostToDevice, s));
        c->capsules_dirty = false;
    }

    if (c->geometry_dirty) {
        // (re)upload the registry.  The world cache and AABBs of meshes that are already in the tree stay valid
        // unless a device buffer has to grow, in which case everything is re-derived (simple and rare).
        PRTC_CUDA(c->d_raw.reserve(c->raw.size()));
        bool world_realloc = c->d_world.cap < c->raw.size() || c->d_aabbs.cap < size_t(n_meshes) * 6;
        PRTC_CUDA(c->d_world.reserve(c->raw.size()));
        PRTC_CUDA(c->d_aabbs.reserve(size_t(n_meshes) * 6));
        PRTC_CUDA(c->d_meshes.reserve(n_meshes));
        PRTC_CUDA(c->d_xforms.reserve(c->models.size()));
        if (n_vertices) PRTC_CUDA(cudaMemcpyAsync(c->d_raw.ptr, c->raw.data(), c->raw.size() * sizeof(float), cudaMemcpyHostToDevice, s));
        std::vector<MeshDev> md(n_meshes);
        for (uint32_t m = 0; m < n_meshes; ++m) md[m] = MeshDev{c->meshes[m].first_vertex, c->meshes[m].n_vertices, c->meshes[m].model, 0u};
        std::vector<ModelXform> xf(c->models.size());
        for (size_t m = 0; m < c->models.size(); ++m) xf[m].m = affine_from(c->models[m].transform);
        if (n_meshes) PRTC_CUDA(cudaMemcpyAsync(c->d_meshes.ptr, md.data(), md.size() * sizeof(MeshDev), cudaMemcpyHostToDevice, s));
        if (!xf.empty()) PRTC_CUDA(cudaMemcpyAsync(c->d_xforms.ptr, xf.data(), xf.size() * sizeof(ModelXform), cudaMemcpyHostToDevice, s));
        PRTC_CUDA(cudaStreamSynchronize(s));     // md / xf are stack-owned

        // K8 over every mesh that is new, moved, or whose cache was lost to a reallocation
        std::vector<float> aabbs(size_t(n_meshes) * 6);
        uint32_t first_new = world_realloc ? 0u : c->meshes_on_device;
        if (n_meshes > first_new) {
            launch_transform_refit(c->d_raw.ptr, c->d_world.ptr, c->d_meshes.ptr, first_new, n_meshes - first_new,
                                   c->d_xforms.ptr, c->d_aabbs.ptr, s);
        }
        for (uint32_t model : c->moved_models) {
            ModelHost const & mh = c->models[model];
            if (mh.first_mesh >= first_new) continue;     // already covered
            launch_transform_refit(c->d_raw.ptr, c->d_world.ptr, c->d_meshes.ptr, mh.first_mesh,
                                   std::min(mh.n_meshes, first_new - mh.first_mesh), c->d_xforms.ptr, c->d_aabbs.ptr, s);
        }
        PRTC_CUDA(cudaGetLastError());
        if (n_meshes) PRTC_CUDA(cudaMemcpyAsync(aabbs.data(), c->d_aabbs.ptr, aabbs.size() * sizeof(float), cudaMemcpyDeviceToHost, s));
        PRTC_CUDA(cudaStreamSynchronize(s));

        auto box_of = [&](uint32_t m) {
            Box b;
            b.lo = v3(aabbs[6 * size_t(m)], aabbs[6 * size_t(m) + 1], aabbs[6 * size_t(m) + 2]);
            b.hi = v3(aabbs[6 * size_t(m) + 3], aabbs[6 * size_t(m) + 4], aabbs[6 * size_t(m) + 5]);
            return b;
        };
        // moved models first (they were registered earlier): DynamicAABBTree::update per model, physics_system.cpp:199-200
        // (LITERAL order has done this in prtc_update_models already, where the reference does it)
        for (uint32_t model : c->moved_models) {
            if (c->literal()) break;
            ModelHost const & mh = c->models[model];
            for (uint32_t m = mh.first_mesh; m < mh.first_mesh + mh.n_meshes && m < c->meshes_in_tree; ++m) {
                Box b = box_of(m);
                c->tree.update(&c->meshes[m].tree_index, &b, 1);
            }
        }
        c->moved_models.clear();
        c->world_max_abs = 0.0f;
        for (float v : aabbs) { float a = std::fabs(v); if (a > c->world_max_abs) c->world_max_abs = a; }   // NaN never wins
        for (int k = 0; k < 3; ++k) { c->world_lo[k] = 3.0e38f; c->world_hi[k] = -3.0e38f; }
        for (uint32_t m = 0; m < n_meshes; ++m)
            for (int k = 0; k < 3; ++k) {
                if (aabbs[6 * size_t(m) + k] < c->world_lo[k]) c->world_lo[k] = aabbs[6 * size_t(m) + k];
                if (aabbs[6 * size_t(m) + 3 + k] > c->world_hi[k]) c->world_hi[k] = aabbs[6 * size_t(m) + 3 + k];
            }
        // new meshes: DynamicAABBTree::insert in registration order, physics_system.cpp:102
        for (uint32_t m = c->meshes_in_tree; m < n_meshes; ++m)
            if (c->models[c->meshes[m].model].n_meshes != 0)      // not removed before its first commit
                c->meshes[m].tree_index = c->tree.insert_leaf(m, LEAF_MESH, box_of(m));
        c->meshes_in_tree = n_meshes;
        c->meshes_on_device = n_meshes;
        c->geometry_dirty = false;
        c->tree_dirty = true;
        c->tri_dirty = true;                      // new / moved / removed geometry: the triangle records are stale
    }

    if (c->tree_dirty) {
        // Triangle records are laid out in LEAF order in CANONICAL mode (the tree is built once; neighbours in the
        // traversal are neighbours in memory) and in REGISTRATION order in LITERAL mode, where the tree changes shape
        // whenever a capsule leaves its fat box: there a re-flatten only uploads the nodes, the triangles stay put.
        const bool by_leaf_order = !c->literal();
        const bool rebuild_tris = by_leaf_order || c->tri_dirty;
        if (rebuild_tris) {
            c->tri_src.clear();
            c->tri_src.reserve(n_vertices / 3);
            if (!by_leaf_order) {
                c->mesh_tri_first.resize(n_meshes);
                for (uint32_t m = 0; m < n_meshes; ++m) {
                    MeshHost const & mh = c->meshes[m];
                    c->mesh_tri_first[m] = uint32_t(c->tri_src.size());
                    for (uint32_t k = 0; k < mh.n_vertices / 3; ++k) c->tri_src.push_back(mh.first_vertex + 3 * k);
                }
            }
        }
        ++c->tree_version;                       // every cached leaf list refers to node indices of the old layout
        c->tree.flatten(c->flat, [&](uint32_t mesh_id, uint32_t & first, uint32_t & count) {
            MeshHost const & mh = c->meshes[mesh_id];
            count = mh.n_vertices / 3;
            if (by_leaf_order) {
                first = uint32_t(c->tri_src.size());
                for (uint32_t k = 0; k < count; ++k) c->tri_src.push_back(mh.first_vertex + 3 * k);
            } else {
                first = c->mesh_tri_first[mesh_id];
            }
        });
        const size_t nn = c->flat.nodes.size();
        const size_t nt = c->tri_src.size();
        PRTC_CUDA(c->d_nodes.reserve(nn * 2));
        PRTC_CUDA(c->d_node_leaf_id.reserve(nn));
        static_assert(sizeof(FlatNode) == 32, "FlatNode must be two float4");
        if (nn) {
            // pageable sources: the runtime stages them before it returns, so the vectors may be rewritten afterwards
            PRTC_CUDA(cudaMemcpyAsync(c->d_nodes.ptr, c->flat.nodes.data(), nn * sizeof(FlatNode), cudaMemcpyHostToDevice, s));
            c->leaf_ids_stale = true;            // only traces and prtc_query_box read them: uploaded on demand
        }
        if (rebuild_tris) {
            PRTC_CUDA(c->d_tri_src.reserve(nt));
            PRTC_CUDA(c->d_tris.reserve(nt * 3));
            if (nt) {
                PRTC_CUDA(cudaMemcpyAsync(c->d_tri_src.ptr, c->tri_src.data(), nt * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
                launch_tri_setup(c->d_world.ptr, c->d_tri_src.ptr, uint32_t(nt), c->d_tris.ptr, s);
                PRTC_CUDA(cudaGetLastError());
            }
            PRTC_CUDA(cudaStreamSynchronize(s));
        }
        c->tree_dirty = false;
        c->tri_dirty = false;
    }
    return PRTC_OK;
}

int fill_params(prtc_ctx * c, float dt, uint32_t n, prtc_transform * d_tf, prtc_char_physics * d_ph, StepParams & p) {
    std::memset(&p, 0, sizeof(p));
    p.nodes = c->d_nodes.ptr;
    p.n_nodes = uint32_t(c->flat.nodes.size());
    p.tris = c->d_tris.ptr;
    p.node_leaf_id = c->d_node_leaf_id.ptr;
    p.capsules = c->d_capsules.ptr;
    p.n_capsules = uint32_t(c->capsules.size());
    p.transforms = d_tf;
    p.physics = d_ph;
    p.n = n;
    p.dt = dt;
    p.gravity = c->cfg.gravity;
    p.ground_dot = c->cfg.ground_dot;
    p.friction = c->cfg.friction;
    p.stick = c->cfg.stick;
    p.substeps = c->cfg.substeps;
    p.abs_mode = c->cfg.abs_mode;
    p.counters = c->d_counters.ptr;
    // pre-cull margin: 64 ulp of the largest world coordinate, at least 1e-3; switched off for absurd extents
    {
        float m = c->world_max_abs;
        float ulp = std::nextafter(m, INFINITY) - m;
        p.qr_eps = std::max(1e-3f, 64.0f * ulp);
        p.qr_enable = (c->quick_reject && m < 1.0e6f) ? 1u : 0u;
        p.tune = (c->frame_cache ? 0u : 1u) | (c->warp_kernel ? 0u : 2u) | c->tune_hi;
        p.work_counter = c->persistent ? c->d_work.ptr : nullptr;
        p.persistent_blocks = uint32_t(c->sm_count) * 20u;
        p.tree_version = c->tree_version;
        p.xc_margin = c->xc_margin;
        if (c->persistent && n > c->xc_capacity) {
            const uint32_t cap = n;
            PRTC_CUDA(c->d_xc_version.reserve(cap)); PRTC_CUDA(c->d_xc_count.reserve(cap));
            PRTC_CUDA(c->d_xc_region.reserve(size_t(cap) * 6)); PRTC_CUDA(c->d_xc_leaf.reserve(size_t(cap) * 12));
            PRTC_CUDA(cudaMemsetAsync(c->d_xc_version.ptr, 0, size_t(cap) * sizeof(uint32_t), c->stream));
            c->xc_capacity = cap;
        }
        if (c->persistent) {
            p.xc_version = c->d_xc_version.ptr; p.xc_region = c->d_xc_region.ptr; p.xc_count = c->d_xc_count.ptr; p.xc_leaf = c->d_xc_leaf.ptr;
            p.xc_stride = c->xc_capacity; p.xc_first = 0;
            p.xc_reuse = c->cross_frame ? 1u : 0u;
        }
    }
    return PRTC_OK;
}

// characters whose collider id names no registered capsule are skipped by the kernels and counted in slot 6
int check_bad_colliders(prtc_ctx * c, cudaStream_t on = nullptr) {
    cudaStream_t s = on ? on : c->stream;
    unsigned long long bad = 0;
    PRTC_CUDA(cudaMemcpyAsync(&bad, c->d_counters.ptr + 6, sizeof(bad), cudaMemcpyDeviceToHost, s));
    PRTC_CUDA(cudaStreamSynchronize(s));
    if (bad) {
        PRTC_CUDA(cudaMemsetAsync(c->d_counters.ptr + 6, 0, sizeof(bad), s));
        return fail(PRTC_ERR_INVALID, "prtc_step_characters: " + std::to_string(bad) + " character(s) name a capsule collider that was never added; their results are undefined");
    }
    return PRTC_OK;
}

int check_mode(prtc_ctx * c, bool device_pointers) {
    if (c->cfg.order_mode != PRTC_ORDER_CANONICAL && c->cfg.order_mode != PRTC_ORDER_LITERAL)
        return fail(PRTC_ERR_INVALID, "unknown order_mode");
    if (c->literal() && device_pointers)
        return fail(PRTC_ERR_INVALID, "PRTC_ORDER_LITERAL maintains the unified tree on the host every frame and needs the "
                                      "host-buffer entry points (prtc_step_characters / _traced)");
    if (c->cfg.dyn_mode != PRTC_DYN_OFF && c->cfg.dyn_mode != PRTC_DYN_JACOBI)
        return fail(PRTC_ERR_INVALID, "unknown dyn_mode");
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

// collider id per node of the flattened tree: what traced kernels and prtc_query_box report
int upload_leaf_ids(prtc_ctx * c) {
    if (!c->leaf_ids_stale) return PRTC_OK;
    const size_t nn = c->flat.node_leaf_id.size();
    PRTC_CUDA(c->d_node_leaf_id.reserve(nn));
    if (nn) PRTC_CUDA(cudaMemcpyAsync(c->d_node_leaf_id.ptr, c->flat.node_leaf_id.data(), nn * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
    c->leaf_ids_stale = false;
    return PRTC_OK;
}

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

// JACOBI dynamic pass: records of this context's characters (phase-1 box, stored-box hysteresis, start-of-frame segment)
int dyn_prepare(prtc_ctx * c, float dt, uint32_t n, const prtc_transform * d_tf, const prtc_char_physics * d_ph) {
    if (!c->dyn_configured) { c->dyn_first = 0; c->dyn_global = n; }
    if (uint64_t(c->dyn_first) + n > c->dyn_global) return fail(PRTC_ERR_INVALID, "prtc_dyn_prepare: slice exceeds the configured global count");
    if (c->d_dyn_records.cap < size_t(c->dyn_global) * 4) {
        if (c->dyn_initialized) return fail(PRTC_ERR_INVALID, "prtc_dyn_prepare: the global character count may not grow after the first frame");
        PRTC_CUDA(c->d_dyn_records.reserve(size_t(c->dyn_global) * 4));
        PRTC_CUDA(cudaMemsetAsync(c->d_dyn_records.ptr, 0, size_t(c->dyn_global) * 4 * sizeof(float4), c->stream));
    }
    launch_dyn_prep(d_tf, d_ph, c->d_capsules.ptr, n, c->dyn_first, c->d_dyn_records.ptr, dt, c->cfg.gravity, c->cfg.leaf_margin,
                    !c->dyn_initialized, c->stream);
    PRTC_CUDA(cudaGetLastError());
    c->dyn_initialized = true;
    c->dyn_prepared = true;
    return PRTC_OK;
}

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
    PRTC_CUDA(c->d_dyn_count.reserve(cells + 1));
    PRTC_CUDA(c->d_dyn_start.reserve(cells + 1));
    PRTC_CUDA(cudaMemsetAsync(c->d_dyn_count.ptr, 0, (cells + 1) * sizeof(uint32_t), s));
    launch_dyn_grid(c->d_dyn_records.ptr, ng, lox, loz, 1.0f / cell, gx, gz, c->d_dyn_count.ptr, nullptr, nullptr, false, s);
    size_t temp_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, temp_bytes, c->d_dyn_count.ptr, c->d_dyn_start.ptr, int(cells + 1), s);
    PRTC_CUDA(c->d_cub_temp.reserve(temp_bytes));
    PRTC_CUDA(cub::DeviceScan::ExclusiveSum(c->d_cub_temp.ptr, temp_bytes, c->d_dyn_count.ptr, c->d_dyn_start.ptr, int(cells + 1), s));
    uint32_t total = 0;
    PRTC_CUDA(cudaMemcpyAsync(&total, c->d_dyn_start.ptr + cells, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    PRTC_CUDA(cudaStreamSynchronize(s));
    PRTC_CUDA(c->d_dyn_items.reserve(size_t(total) + 1));
    PRTC_CUDA(cudaMemsetAsync(c->d_dyn_count.ptr, 0, (cells + 1) * sizeof(uint32_t), s));
    launch_dyn_grid(c->d_dyn_records.ptr, ng, lox, loz, 1.0f / cell, gx, gz, c->d_dyn_count.ptr, c->d_dyn_start.ptr, c->d_dyn_items.ptr, true, s);
    PRTC_CUDA(cudaGetLastError());
    p.dyn_records = c->d_dyn_records.ptr;
    p.dyn_cell_start = c->d_dyn_start.ptr;
    p.dyn_items = c->d_dyn_items.ptr;
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
        fill_params(c, dt, n, s_tf, s_ph, p);
        p.cap2char = s_cap2char; p.seg_start = s_seg_start; p.seg_prev = s_seg_prev; p.last_box = s_last_box;
        p.counters = scratch;
        p.host_flags = flags;
        c->launches += uint64_t(launch_step(p, false, STEP_LITERAL, s));
        PRTC_CUDA(cudaGetLastError());
        PRTC_CUDA(cudaStreamSynchronize(s));
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
        fill_params(c, dt, n, c->d_tf.ptr, c->d_ph.ptr, p);
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

void prtc_default_config(prtc_config * cfg) {
    if (!cfg) return;
    cfg->gravity = 1.0f;
    cfg->leaf_margin = 0.05f;
    cfg->ground_dot = 0.3f;
    cfg->friction = 10.0f;
    cfg->stick = 0.05f;
    cfg->substeps = 4;
    cfg->abs_mode = PRTC_ABS_INT;
    cfg->order_mode = PRTC_ORDER_CANONICAL;
    cfg->dyn_mode = PRTC_DYN_OFF;
    cfg->device = 0;
}

const char * prtc_last_error(void) { return g_last_error.c_str(); }

int prtc_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    int ok = 0;
    for (int d = 0; d < n; ++d) {
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, d) == cudaSuccess && prop.major == 10) ++ok;
    }
    return ok;
}

int prtc_create(const prtc_config * cfg, prtc_ctx ** out) {
    if (!out) return fail(PRTC_ERR_INVALID, "prtc_create: out is null");
    *out = nullptr;
    prtc_config c;
    if (cfg) c = *cfg; else prtc_default_config(&c);
    if (c.substeps == 0) return fail(PRTC_ERR_INVALID, "prtc_create: substeps must be > 0");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(PRTC_ERR_CUDA, std::string("prtc_create: no CUDA device (") + (e != cudaSuccess ? cudaGetErrorString(e) : "count 0") +
                                   "); this library has no CPU path");
    }
    if (c.device < 0 || c.device >= n) return fail(PRTC_ERR_INVALID, "prtc_create: device ordinal out of range");
    cudaDeviceProp prop;
    PRTC_CUDA(cudaGetDeviceProperties(&prop, c.device));
    if (prop.major != 10)
        return fail(PRTC_ERR_CUDA, std::string("prtc_create: device '") + prop.name + "' is not compute capability 10.x; kernels are built for sm_100a only");
    PRTC_CUDA(cudaSetDevice(c.device));
    prtc_ctx * ctx = new (std::nothrow) prtc_ctx();
    if (!ctx) return fail(PRTC_ERR_ALLOC, "prtc_create: out of host memory");
    ctx->cfg = c;
    // environment overrides of the option defaults, for A/B measurements without touching the caller
    if (const char * e = std::getenv("PRTC_QUICK_REJECT")) ctx->quick_reject = std::atoi(e);
    if (const char * e = std::getenv("PRTC_FRAME_CACHE")) ctx->frame_cache = std::atoi(e);
    if (const char * e = std::getenv("PRTC_PERSISTENT")) ctx->persistent = std::atoi(e);
    if (const char * e = std::getenv("PRTC_CROSS_FRAME")) ctx->cross_frame = std::atoi(e);
    if (const char * e = std::getenv("PRTC_WARP_KERNEL")) ctx->warp_kernel = std::atoi(e);
    if (const char * e = std::getenv("PRTC_STREAMED_COPY")) ctx->streamed_copy = std::atoi(e);
    if (const char * e = std::getenv("PRTC_STAGED")) ctx->staged = std::atoi(e);
    if (const char * e = std::getenv("PRTC_XC_MARGIN")) ctx->xc_margin = float(std::atof(e));
    if (const char * e = std::getenv("PRTC_TUNE")) ctx->tune_hi = uint32_t(std::strtoul(e, nullptr, 0)) & 0xFFFFFF00u;
    ctx->sm_count = prop.multiProcessorCount;
    ctx->tree.set_margin(c.leaf_margin);
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete ctx;
        return fail(PRTC_ERR_CUDA, "prtc_create: cudaStreamCreate failed");
    }
    ctx->own_stream = true;
    if (ctx->d_work.reserve(16) != cudaSuccess || ctx->d_counters.reserve(16) != cudaSuccess || cudaMemsetAsync(ctx->d_counters.ptr, 0, 16 * sizeof(unsigned long long), ctx->stream) != cudaSuccess) {
        prtc_destroy(ctx);
        return fail(PRTC_ERR_CUDA, "prtc_create: counter allocation failed");
    }
    *out = ctx;
    return PRTC_OK;
}

void prtc_destroy(prtc_ctx * c) {
    if (!c) return;
    cudaSetDevice(c->cfg.device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    c->d_raw.release(); c->d_world.release(); c->d_aabbs.release(); c->d_meshes.release(); c->d_xforms.release();
    c->d_nodes.release(); c->d_tris.release(); c->d_node_leaf_id.release(); c->d_tri_src.release();
    c->d_capsules.release(); c->d_counters.release(); c->d_work.release(); c->d_ingest.release(); c->stage.release();
    c->d_xc_version.release(); c->d_xc_count.release(); c->d_xc_leaf.release(); c->d_xc_region.release(); c->d_tf.release(); c->d_ph.release();
    c->d_tq.release(); c->d_tmesh.release(); c->d_tcaps.release(); c->d_thits.release();
    c->d_scratch_f.release(); c->d_scratch_u.release();
    c->d_cap2char.release(); c->d_seg_start.release(); c->d_seg_prev.release(); c->d_last_box.release();
    c->d_dyn_records.release(); c->d_dyn_count.release(); c->d_dyn_start.release(); c->d_dyn_items.release(); c->d_cub_temp.release();
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    for (int k = 0; k < prtc_ctx::PIPE_STREAMS; ++k) if (c->pipe[k]) cudaStreamDestroy(c->pipe[k]);
    if (c->pipe_ready) cudaEventDestroy(c->pipe_ready);
    if (c->ingest_reset) cudaEventDestroy(c->ingest_reset);
    delete c;
}

namespace {
// world-space AABB of one registered mesh under `xf`, with the arithmetic of k_transform_refit (physics_system.cpp:80-91)
Box host_mesh_box(prtc_ctx const * c, MeshHost const & me, Affine const & xf) {
    V3 mn = v3(3.402823466e+38f), mx = v3(-3.402823466e+38f);
    for (uint32_t k = 0; k < me.n_vertices; ++k) {
        const float * r = &c->raw[3 * size_t(me.first_vertex + k)];
        V3 w = affine_point(xf, v3(r[0], r[1], r[2]));
        mn = gmin(mn, w);
        mx = gmax(mx, w);
    }
    Box b; b.lo = mn; b.hi = mx;
    return b;
}
} // namespace

int prtc_add_model(prtc_ctx * c, const float * xyz, uint32_t n_vertices, const uint32_t * mesh_first,
                   const uint32_t * mesh_count, uint32_t n_meshes, const prtc_transform * transform,
                   uint32_t * out_model_id, uint32_t * out_first_mesh_id) {
    if (!c || (!xyz && n_vertices) || (n_meshes && (!mesh_first || !mesh_count)) || !transform)
        return fail(PRTC_ERR_INVALID, "prtc_add_model: null argument");
    for (uint32_t m = 0; m < n_meshes; ++m) {
        if (mesh_count[m] % 3 != 0 || uint64_t(mesh_first[m]) + mesh_count[m] > n_vertices)
            return fail(PRTC_ERR_INVALID, "prtc_add_model: mesh range must be a multiple of 3 vertices inside the model");
    }
    // physics_system.cpp:66-91: vertices are copied mesh by mesh in mesh order (geometry.raw[i], i running)
    uint32_t model_id = uint32_t(c->models.size());
    ModelHost mh;
    mh.first_mesh = uint32_t(c->meshes.size());
    mh.n_meshes = n_meshes;
    mh.transform = *transform;
    for (uint32_t m = 0; m < n_meshes; ++m) {
        MeshHost me;
        me.first_vertex = uint32_t(c->raw.size() / 3);
        me.n_vertices = mesh_count[m];
        me.model = model_id;
        me.tree_index = -1;
        c->raw.insert(c->raw.end(), xyz + 3 * size_t(mesh_first[m]), xyz + 3 * (size_t(mesh_first[m]) + mesh_count[m]));
        c->meshes.push_back(me);
    }
    c->models.push_back(mh);
    c->geometry_dirty = true;
    if (c->literal()) {
        // the reference inserts the mesh leaves right here (physics_system.cpp:102), interleaved with capsule leaves
        // in call order, and the tree's shape depends on that order: derive the AABBs on the host with the same
        // arithmetic as k_transform_refit instead of waiting for the commit
        const Affine xf = affine_from(mh.transform);
        for (uint32_t m = mh.first_mesh; m < mh.first_mesh + n_meshes; ++m) {
            MeshHost & me = c->meshes[m];
            me.tree_index = c->tree.insert_leaf(m, LEAF_MESH, host_mesh_box(c, me, xf));
        }
        c->meshes_in_tree = uint32_t(c->meshes.size());
        c->tree_dirty = true;
    }
    if (out_model_id) *out_model_id = model_id;
    if (out_first_mesh_id) *out_first_mesh_id = mh.first_mesh;
    return PRTC_OK;
}

int prtc_add_capsule(prtc_ctx * c, float height, float radius, const float offset[3], uint32_t * out_id) {
    if (!c || !offset) return fail(PRTC_ERR_INVALID, "prtc_add_capsule: null argument");
    CapsuleDev d;
    d.height = height; d.radius = radius; d.ox = offset[0]; d.oy = offset[1]; d.oz = offset[2];
    d.pad0 = d.pad1 = d.pad2 = 0.0f;
    if (out_id) *out_id = uint32_t(c->capsules.size());
    c->capsules.push_back(d);
    c->capsules_dirty = true;
    if (c->literal()) {
        // physics_system.cpp:36-39: the capsule's leaf is created at the IDENTITY pose
        Quat qi; qi.x = qi.y = qi.z = 0.0f; qi.w = 1.0f;
        V3 A, B;
        segment_ends(capsule_frame(d, rotation_of(qi)), v3(0.0f), A, B);
        Box b = capsule_box(A, B, d.radius);
        c->capsule_aabbs.push_back(b);
        c->capsule_tree_index.push_back(c->tree.insert_leaf(uint32_t(c->capsules.size() - 1), LEAF_CAPSULE, b));
        c->tree_dirty = true;
    }
    return PRTC_OK;
}

int prtc_update_capsule(prtc_ctx * c, uint32_t id, float height, float radius, const float offset[3]) {
    if (!c || !offset || id >= c->capsules.size()) return fail(PRTC_ERR_INVALID, "prtc_update_capsule: bad capsule id");
    CapsuleDev & d = c->capsules[id];
    d.height = height; d.radius = radius; d.ox = offset[0]; d.oy = offset[1]; d.oz = offset[2];
    c->capsules_dirty = true;
    return PRTC_OK;
}

int prtc_commit(prtc_ctx * c) {
    if (!c) return fail(PRTC_ERR_INVALID, "prtc_commit: null context");
    PRTC_CUDA(cudaSetDevice(c->cfg.device));
    return commit_impl(c);
}

int prtc_update_models(prtc_ctx * c, const uint32_t * model_ids, const prtc_transform * transforms, uint32_t n) {
    if (!c || (n && (!model_ids || !transforms))) return fail(PRTC_ERR_INVALID, "prtc_update_models: null argument");
    for (uint32_t i = 0; i < n; ++i) {
        if (model_ids[i] >= c->models.size()) return fail(PRTC_ERR_INVALID, "prtc_update_models: bad model id");
        c->models[model_ids[i]].transform = transforms[i];
        c->moved_models.push_back(model_ids[i]);
        if (c->literal()) {
            // the reference updates the model's leaves right here (physics_system.cpp:199-200), i.e. BEFORE the capsule
            // leaves of the next updateCharacters move -- and the tree's shape depends on that order
            ModelHost const & mh = c->models[model_ids[i]];
            const Affine xf = affine_from(mh.transform);
            for (uint32_t m = mh.first_mesh; m < mh.first_mesh + mh.n_meshes && m < c->meshes_in_tree; ++m) {
                Box b = host_mesh_box(c, c->meshes[m], xf);
                if (c->tree.update(&c->meshes[m].tree_index, &b, 1) != 0) c->tree_dirty = true;
            }
        }
    }
    if (n) c->geometry_dirty = true;
    return PRTC_OK;
}

int prtc_set_stream(prtc_ctx * c, void * cuda_stream) {
    if (!c) return fail(PRTC_ERR_INVALID, "prtc_set_stream: null context");
    if (c->own_stream && c->stream) { cudaStreamSynchronize(c->stream); cudaStreamDestroy(c->stream); }
    c->stream = static_cast<cudaStream_t>(cuda_stream);
    c->own_stream = false;
    return PRTC_OK;
}

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
    fill_params(c, dt, n, d_tf, d_ph, p);
    int mode = STEP_STATIC;
    if (c->jacobi()) {
        if (!c->dyn_prepared) { rc = dyn_prepare(c, dt, n, d_tf, d_ph); if (rc) return rc; }
        rc = dyn_build_grid(c, p);
        if (rc) return rc;
        c->dyn_prepared = false;
        mode = STEP_JACOBI;
    }
    c->launches += uint64_t(launch_step(p, false, mode, c->stream));
    if (mode == STEP_JACOBI) c->launches += 3;                // k_dyn_prep + two k_dyn_grid passes (+ cub scan)
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
    static const uint32_t in_first = [] { const char * e = std::getenv("PRTC_INGEST_IN_FIRST"); return e ? uint32_t(std::atoi(e)) : 32768u; }();
    static const uint32_t in_max = [] { const char * e = std::getenv("PRTC_INGEST_IN_MAX"); return e ? uint32_t(std::atoi(e)) : 65536u; }();
    for (uint32_t first = 0, piece = in_first & ~31u; first < n; first += piece, piece = std::min(2 * piece, in_max & ~31u)) {
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
    static const uint32_t kChunk = [] { const char * e = std::getenv("PRTC_CHUNK"); return e ? uint32_t(std::atoi(e)) : 65536u; }();
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
            fill_params(c, dt, cnt, c->d_tf.ptr + first, c->d_ph.ptr + first, p);
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
        c->launches += uint64_t(launch_step(p, false, STEP_STATIC, c->stream));
        PRTC_CUDA(cudaGetLastError());
        PRTC_CUDA(cudaStreamSynchronize(c->stream));
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
    fill_params(c, dt, n, c->d_tf.ptr, c->d_ph.ptr, p);
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
        PRTC_CUDA(c->stage.reserve(size_t(n) * 11 * sizeof(float) + 6 * 64));
        StageCursor cur{c->stage.base};
        float * s_o = cur.take<float>(3 * size_t(n)), * s_d = cur.take<float>(3 * size_t(n)), * s_m = cur.take<float>(n);
        float * s_h = cur.take<float>(3 * size_t(n));
        unsigned char * s_mask = cur.take<unsigned char>(n);
        std::memcpy(s_o, origins, size_t(n) * 3 * sizeof(float));
        std::memcpy(s_d, directions, size_t(n) * 3 * sizeof(float));
        std::memcpy(s_m, max_distances, size_t(n) * sizeof(float));
        std::memset(s_h, 0, size_t(n) * 3 * sizeof(float));
        launch_raycast(c->d_nodes.ptr, uint32_t(c->flat.nodes.size()), c->d_tris.ptr, n, s_o, s_d, s_m, s_h, s_mask, s);
        PRTC_CUDA(cudaGetLastError());
        PRTC_CUDA(cudaStreamSynchronize(s));
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

int prtc_remove_model(prtc_ctx * c, uint32_t model_id) {
    if (!c || model_id >= c->models.size()) return fail(PRTC_ERR_INVALID, "prtc_remove_model: bad model id");
    ModelHost & mh = c->models[model_id];
    // DynamicAABBTree::remove over the model's mesh leaves in order (physics_system.cpp:136)
    for (uint32_t m = mh.first_mesh; m < mh.first_mesh + mh.n_meshes; ++m) {
        MeshHost & me = c->meshes[m];
        if (m < c->meshes_in_tree && me.tree_index >= 0) c->tree.remove(me.tree_index);
        me.tree_index = -1;
        me.n_vertices = 0;          // the vertices stay in the registry but are never referenced again
    }
    mh.n_meshes = 0;
    c->geometry_dirty = true;
    c->tree_dirty = true;
    return PRTC_OK;
}

int prtc_get_counters(prtc_ctx * c, prtc_counters * out, int reset) {
    if (!c || !out) return fail(PRTC_ERR_INVALID, "prtc_get_counters: null argument");
    PRTC_CUDA(cudaSetDevice(c->cfg.device));
    unsigned long long h[8];
    PRTC_CUDA(cudaMemcpyAsync(h, c->d_counters.ptr, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    if (reset) PRTC_CUDA(cudaMemsetAsync(c->d_counters.ptr, 0, sizeof(h), c->stream));
    PRTC_CUDA(cudaStreamSynchronize(c->stream));
    out->substeps = h[0]; out->tri_tests = h[1]; out->node_tests = h[2]; out->contacts = h[3]; out->cc_tests = h[4]; out->exact_tests = h[5];
    out->launches = c->launches;
    if (reset) c->launches = 0;
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
    if (!c || (n && (!d_tf || !d_ph))) return fail(PRTC_ERR_INVALID, "prtc_dyn_prepare: null argument");
    if (!c->jacobi()) return fail(PRTC_ERR_INVALID, "prtc_dyn_prepare: context is not in PRTC_ORDER_CANONICAL + PRTC_DYN_JACOBI");
    PRTC_CUDA(cudaSetDevice(c->cfg.device));
    if (c->geometry_dirty || c->tree_dirty || c->capsules_dirty) { int rc = commit_impl(c); if (rc) return rc; }
    return dyn_prepare(c, dt, n, d_tf, d_ph);
}

int prtc_dyn_buffer(prtc_ctx * c, void ** d_records, size_t * record_bytes, uint32_t * first_global, uint32_t * n_global) {
    if (!c || !d_records) return fail(PRTC_ERR_INVALID, "prtc_dyn_buffer: null argument");
    if (!c->d_dyn_records.ptr) return fail(PRTC_ERR_INVALID, "prtc_dyn_buffer: call prtc_dyn_prepare first");
    *d_records = c->d_dyn_records.ptr;
    if (record_bytes) *record_bytes = 4 * sizeof(float4);
    if (first_global) *first_global = c->dyn_first;
    if (n_global) *n_global = c->dyn_global;
    return PRTC_OK;
}

int prtc_set_option(prtc_ctx * c, int option, int value) {
    if (!c) return fail(PRTC_ERR_INVALID, "prtc_set_option: null context");
    switch (option) {
        case PRTC_OPT_FRAME_CACHE: c->frame_cache = value != 0; return PRTC_OK;
        case PRTC_OPT_QUICK_REJECT: c->quick_reject = value != 0; return PRTC_OK;
        case PRTC_OPT_PERSISTENT: c->persistent = value != 0; return PRTC_OK;
        case PRTC_OPT_CROSS_FRAME_CACHE: c->cross_frame = value != 0; return PRTC_OK;
        case PRTC_OPT_WARP_KERNEL: c->warp_kernel = value != 0; return PRTC_OK;
        case PRTC_OPT_STREAMED_COPY: c->streamed_copy = value != 0; return PRTC_OK;
        default: return fail(PRTC_ERR_INVALID, "prtc_set_option: unknown option");
    }
}

int prtc_get_tree_info(prtc_ctx * c, uint32_t * n_nodes, uint32_t * n_leaves, uint32_t * n_triangles) {
    if (!c) return fail(PRTC_ERR_INVALID, "prtc_get_tree_info: null context");
    if (n_nodes) *n_nodes = uint32_t(c->flat.nodes.size());
    if (n_leaves) *n_leaves = uint32_t(c->flat.leaf_order.size());
    if (n_triangles) *n_triangles = uint32_t(c->tri_src.size());
    return PRTC_OK;
}

int prtc_host_build_tree(const float * aabbs6, uint32_t n, float margin, uint32_t * leaf_order, float * nodes8,
                         uint32_t nodes_capacity, uint32_t * n_nodes) {
    if ((n && !aabbs6) || !n_nodes) return fail(PRTC_ERR_INVALID, "prtc_host_build_tree: null argument");
    RefTree tree(margin);
    for (uint32_t i = 0; i < n; ++i) {
        Box b;
        b.lo = v3(aabbs6[6 * size_t(i)], aabbs6[6 * size_t(i) + 1], aabbs6[6 * size_t(i) + 2]);
        b.hi = v3(aabbs6[6 * size_t(i) + 3], aabbs6[6 * size_t(i) + 4], aabbs6[6 * size_t(i) + 5]);
        tree.insert_leaf(i, LEAF_MESH, b);
    }
    FlatTree flat;
    tree.flatten(flat, [&](uint32_t id, uint32_t & first, uint32_t & count) { first = id; count = 1; });
    *n_nodes = uint32_t(flat.nodes.size());
    if (leaf_order) std::copy(flat.leaf_order.begin(), flat.leaf_order.end(), leaf_order);
    if (nodes8) {
        if (nodes_capacity < flat.nodes.size()) return fail(PRTC_ERR_CAPACITY, "prtc_host_build_tree: node buffer too small");
        std::memcpy(nodes8, flat.nodes.data(), flat.nodes.size() * sizeof(FlatNode));
    }
    return PRTC_OK;
}

int prtc_get_leaf_order(prtc_ctx * c, uint32_t * leaf_mesh_ids, uint32_t capacity) {
    if (!c || !leaf_mesh_ids) return fail(PRTC_ERR_INVALID, "prtc_get_leaf_order: null argument");
    if (capacity < c->flat.leaf_order.size()) return fail(PRTC_ERR_CAPACITY, "prtc_get_leaf_order: buffer too small");
    std::copy(c->flat.leaf_order.begin(), c->flat.leaf_order.end(), leaf_mesh_ids);
    return PRTC_OK;
}

int prtc_get_world_vertices(prtc_ctx * c, float * xyz, uint32_t capacity_vertices, uint32_t * n_vertices) {
    if (!c) return fail(PRTC_ERR_INVALID, "prtc_get_world_vertices: null context");
    uint32_t nv = uint32_t(c->raw.size() / 3);
    if (n_vertices) *n_vertices = nv;
    if (!xyz) return PRTC_OK;
    if (capacity_vertices < nv) return fail(PRTC_ERR_CAPACITY, "prtc_get_world_vertices: buffer too small");
    PRTC_CUDA(cudaSetDevice(c->cfg.device));
    if (nv) PRTC_CUDA(cudaMemcpyAsync(xyz, c->d_world.ptr, size_t(nv) * 3 * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    PRTC_CUDA(cudaStreamSynchronize(c->stream));
    return PRTC_OK;
}

int prtc_query_box(prtc_ctx * c, const float aabb6[6], uint32_t * mesh_ids, uint32_t capacity, uint32_t * n_found) {
    if (!c || !aabb6 || !n_found || (capacity && !mesh_ids)) return fail(PRTC_ERR_INVALID, "prtc_query_box: null argument");
    PRTC_CUDA(cudaSetDevice(c->cfg.device));
    if (c->geometry_dirty || c->tree_dirty || c->capsules_dirty) { int rc = commit_impl(c); if (rc) return rc; }
    { int rc = upload_leaf_ids(c); if (rc) return rc; }
    PRTC_CUDA(c->d_scratch_f.reserve(6));
    PRTC_CUDA(c->d_scratch_u.reserve(size_t(capacity) + 1));
    PRTC_CUDA(cudaMemcpyAsync(c->d_scratch_f.ptr, aabb6, 6 * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    launch_query_box(c->d_nodes.ptr, uint32_t(c->flat.nodes.size()), c->d_node_leaf_id.ptr, c->d_scratch_f.ptr,
                     c->d_scratch_u.ptr + 1, capacity, c->d_scratch_u.ptr, c->stream);
    PRTC_CUDA(cudaGetLastError());
    uint32_t found = 0;
    PRTC_CUDA(cudaMemcpyAsync(&found, c->d_scratch_u.ptr, sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    PRTC_CUDA(cudaStreamSynchronize(c->stream));
    *n_found = found;
    uint32_t ncopy = std::min(found, capacity);
    if (ncopy) PRTC_CUDA(cudaMemcpy(mesh_ids, c->d_scratch_u.ptr + 1, ncopy * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    if (found > capacity) return fail(PRTC_ERR_CAPACITY, "prtc_query_box: more candidates than capacity");
    return PRTC_OK;
}

} // extern "C"
