// prtc_internal.h -- what the two host translation units of libprtc.so share: error reporting, device / staging buffers,
// the context behind prtc_ctx and the commit / parameter helpers.  Not installed; include/prtc.h is the interface.
#ifndef PRTC_INTERNAL_H
#define PRTC_INTERNAL_H

#include <cuda.h>
#include <cuda_runtime.h>

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

namespace prt {
namespace host {

int fail(int code, const std::string & msg);          // records the thread's last error text, returns `code`

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

inline Affine affine_from(prtc_transform const & t) {
    Quat q; q.x = t.rotation[0]; q.y = t.rotation[1]; q.z = t.rotation[2]; q.w = t.rotation[3];
    return affine_of(v3(t.position[0], t.position[1], t.position[2]), q, v3(t.scale[0], t.scale[1], t.scale[2]));
}

} // namespace host
} // namespace prt

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
    std::vector<prt::host::MeshHost> meshes;
    std::vector<prt::host::ModelHost> models;
    std::vector<prt::CapsuleDev> capsules;
    prt::RefTree tree;
    prt::FlatTree flat;
    void * flat_pinned = nullptr;           // flat.nodes' storage when it is page-locked (cudaHostRegister)
    size_t flat_pinned_bytes = 0;
    std::vector<uint32_t> tri_src;          // per leaf-ordered triangle: index of its first vertex in `raw`/`world`
    uint32_t meshes_in_tree = 0;            // meshes that have a leaf in `tree`
    uint32_t meshes_on_device = 0;          // meshes whose world cache / AABB the device has computed
    // LITERAL order: the reference's unified tree also holds one leaf per capsule (physics_system.cpp:36-39)
    std::vector<int32_t> capsule_tree_index;
    std::vector<prt::Box> capsule_aabbs;         // m_aabbData.capsuleAABBs
    bool geometry_dirty = false;            // raw / meshes / xforms changed since the last upload
    bool registry_dirty = false;            // a model was added or removed since the last commit (a moved model alone takes the incremental path)
    uint32_t incremental_commits = 0, full_commits = 0;     // statistics (prtc_get_commit_stats)
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
    int tight_reject = 1;                   // PRTC_OPT_TIGHT_REJECT (k_step_pool)
    float push_slack = 0.0025f;               // k_step_pool: push budget of a culled window (PRTC_PUSH_SLACK)
    int pool_kernel = 1;                    // PRTC_OPT_POOL_KERNEL: persistent launches run k_step_pool (0: k_step_stream)
    std::vector<uint32_t> moved_models;     // models whose transform changed (prtc_update_models)
    std::vector<float> h_aabbs;             // host copy of the mesh AABBs the device computed (6 floats per mesh)

    // device
    prt::host::DevBuf<float> d_raw, d_world, d_aabbs;
    prt::host::DevBuf<prt::MeshDev> d_meshes;
    prt::host::DevBuf<prt::ModelXform> d_xforms;
    prt::host::DevBuf<float4> d_nodes, d_tris, d_cull;
    prt::host::DevBuf<float4> d_leaves;     // BY_ID: leaf records handed to the device tree build
    std::vector<float> leaf_fat;            // BY_ID: stored fat box of every mesh leaf (6 floats), kept with the reference's update rule
    prt::host::DevBuf<float4> d_top;        // top-of-tree table (bvh_builder.h build_top_table), CANONICAL order only
    std::vector<prt::FlatNode> top;
    std::vector<uint32_t> top_src;          // node index behind every table entry
    prt::host::DevBuf<uint32_t> d_top_src;
    std::vector<uint32_t> leaf_slot;        // BY_ID: position of every mesh in the device's leaf-record array (-1: removed)
    prt::host::DevBuf<uint32_t> d_node_leaf_id, d_tri_src;
    prt::host::DevBuf<prt::CapsuleDev> d_capsules;
    prt::host::DevBuf<unsigned long long> d_counters;
    // cross-frame leaf cache of the persistent kernel (PRTC_OPT_CROSS_FRAME_CACHE)
    prt::host::DevBuf<uint32_t> d_xc_version, d_xc_count, d_xc_leaf;
    prt::host::DevBuf<float> d_xc_region;
    uint32_t xc_capacity = 0;
    uint32_t tree_version = 1;
    uint64_t launches = 0;                  // kernels launched by step calls since the last counter reset
    int cross_frame = 1;
    int top_staging = 1;                    // PRTC_OPT_TOP_STAGING: k_xc_refresh walks the top levels out of shared memory
    int streamed_copy = 1;                  // PRTC_OPT_STREAMED_COPY: big host-buffer batches feed one persistent launch while it runs
    int warp_kernel = 1;                    // small batches: one warp per character (k_step_warp); 0 = thinned lock-step kernel
    float xc_margin = 0.5f;
    prt::host::DevBuf<unsigned> d_work;                // work counters of the persistent kernel, one per concurrently running launch ([8], [9]: the pooled kernel's alternating pair)
    prt::ChainState chain;                             // chained frames (PRTC_OPT_CHAIN_FRAMES)
    int chain_frames = 0;
    prt::host::DevBuf<uint32_t> d_chain_done;
    prt::host::DevBuf<unsigned> d_chain_ring;          // CHAIN_RING work counters
    uint32_t work_flip = 0;                            // which of the pair the next pooled launch uses
    prt::host::HostStage stage;                        // small host-buffer batches: records staged in mapped pinned memory
    int staged = 1;                         // PRTC_STAGED=0: small batches take the copy path too (A/B)
    prt::host::DevBuf<uint32_t> d_ingest;              // streamed host-buffer step: [0] arrival gate, then per-chunk and per-block completion counters
    cudaEvent_t ingest_reset = nullptr;
    prt::host::DevBuf<prtc_transform> d_tf;
    prt::host::DevBuf<prtc_char_physics> d_ph;
    // trace scratch
    prt::host::DevBuf<prt::TraceQuery> d_tq;
    prt::host::DevBuf<uint32_t> d_tmesh, d_tcaps;
    prt::host::DevBuf<prt::TraceHit> d_thits;
    prt::host::DevBuf<float> d_scratch_f;
    prt::host::DevBuf<uint32_t> d_scratch_u;
    // LITERAL order
    prt::host::DevBuf<uint32_t> d_cap2char;
    prt::host::DevBuf<float4> d_seg_start, d_seg_prev;
    prt::host::DevBuf<float> d_last_box;
    // CANONICAL order, JACOBI dynamic pass
    prt::host::DevBuf<float4> d_dyn_records;
    prt::host::DevBuf<uint32_t> d_dyn_count, d_dyn_start, d_dyn_items, d_dyn_reach, d_dyn_cursor;
    prt::host::DevBuf<unsigned long long> d_dyn_scan;      // look-back words of k_dyn_scan + its ticket counter
    uint32_t dyn_scan_tickets = 0, dyn_scan_launches = 0;
    prt::host::DevBuf<unsigned long long> d_timeline;      // PRTC_TIMELINE=1 (diagnostic): claim timeline of the last pooled launch
    int timeline = 0;
    uint32_t dyn_first = 0, dyn_global = 0;  // this context's slice [dyn_first, dyn_first + n) of dyn_global characters
    bool dyn_configured = false;             // prtc_dyn_configure called (multi-rank); otherwise the slice is the whole batch
    bool dyn_initialized = false;            // stored boxes exist (first prepare creates them at the identity pose)
    bool dyn_prepared = false;               // prtc_dyn_prepare ran for the coming step (records may have been exchanged since)
    prt::DynPeers dyn_peers = {};            // prtc_multi: the other devices' record arrays, written by this device's k_dyn_prep
    // exchange of the records ACROSS PROCESSES (prtc_dyn_ipc_*): the record array is double-buffered (a fast rank writes
    // frame f+1's records into its peers while they still step frame f) and every rank's arrival is a word in device memory
    struct DynIpc {
        bool on = false;
        uint32_t rank = 0, world = 0, frame = 0;
        prt::host::DevBuf<uint32_t> flags;           // [32] arrival table (frame numbers) + block counter at [31]
        float4 * peer_records[15] = {};              // base of the peers' (2 x n_global records) arrays
        uint32_t * peer_flags[15] = {};
        uint32_t n_peers = 0;
        std::vector<void *> opened;
        cudaStream_t push_stream = nullptr;          // the peer stores run beside the slot refresh of the context's stream
        cudaEvent_t prep_done = nullptr;
    } dyn_ipc;
    float world_lo[3] = {0, 0, 0}, world_hi[3] = {0, 0, 0};
    // resident state (prtc_resident_*): the records stay in d_tf / d_ph between frames; per-frame I/O through mapped pinned arrays
    bool resident = false;
    uint32_t resident_n = 0;
    prt::host::HostStage resident_io;        // [flags 64 B][movement 3n floats][result 4n floats]
    float * resident_move = nullptr;
    float * resident_result = nullptr;
    uint32_t * resident_flags = nullptr;     // [0] unused, [1] unknown capsule id, [2] completion doorbell
    uint32_t resident_frame = 0;
    uint32_t resident_prefresh = 0;          // tree version the per-slot leaf lists were refreshed for BEHIND the last resident frame (0: not)
    uint32_t resident_pending = 0;           // doorbell value of the frame in flight (0: wait with cudaStreamSynchronize)
    int doorbell = 1;                        // PRTC_DOORBELL=0: resident steps wait in cudaStreamSynchronize
    bool literal() const { return cfg.order_mode == PRTC_ORDER_LITERAL; }
    bool by_id() const { return cfg.order_mode == PRTC_ORDER_BY_ID; }
    bool jacobi() const { return cfg.order_mode != PRTC_ORDER_LITERAL && cfg.dyn_mode == PRTC_DYN_JACOBI; }
};

namespace prt {
namespace host {

// ---- prtc.cu ---------------------------------------------------------------------------------------------------
int commit_impl(prtc_ctx * c);
int fill_params(prtc_ctx * c, float dt, uint32_t n, prtc_transform * d_tf, prtc_char_physics * d_ph, StepParams & p);
int check_bad_colliders(prtc_ctx * c, cudaStream_t on = nullptr);
int check_mode(prtc_ctx * c, bool device_pointers);
int upload_leaf_ids(prtc_ctx * c);               // collider id per node, for traced kernels and prtc_query_box
// ---- prtc_step.cu ----------------------------------------------------------------------------------------------
int dyn_reserve(prtc_ctx * c, uint32_t n);       // allocates the record array of the JACOBI pass (dyn_first / dyn_global set)
int dyn_prepare(prtc_ctx * c, float dt, uint32_t n, const prtc_transform * d_tf, const prtc_char_physics * d_ph);
// resident state, in pieces (prtc_multi drives several contexts through them): `io` = caller-provided mapped pinned memory
// ([flags 64 B] / movement / result) or null for the context's own
int resident_begin_impl(prtc_ctx * c, uint32_t n, const prtc_transform * tf, const prtc_char_physics * ph, uint32_t * io_flags, float * io_move, float * io_result);
int resident_launch(prtc_ctx * c, float dt);     // enqueues one frame
int resident_wait(prtc_ctx * c);                 // returns when it is complete
int dyn_build_grid_for(prtc_ctx * c, StepParams & p);   // grid over the (exchanged) records, wired into the launch parameters

} // namespace host
} // namespace prt

#endif
