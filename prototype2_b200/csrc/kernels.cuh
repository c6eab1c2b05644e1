// kernels.cuh -- launch parameter blocks shared by kernels.cu (device) and prtc.cu (host side of the C-ABI).
#ifndef PRT_KERNELS_CUH
#define PRT_KERNELS_CUH

#include <cuda_runtime.h>
#include <cstdint>

#include "../../include/prtc.h"
#include "collide.cuh"

namespace prt {

// per-(character, substep) trace slot written by the traced kernel variant
struct TraceQuery {
    float aabb[6];
    uint32_t executed;      // 1 when the substep ran
    uint32_t n_mesh;        // candidate mesh colliders found (may exceed the slot capacity)
    uint32_t n_caps;
    uint32_t n_hits;
    uint32_t tri_tests;
    uint32_t pad;
};
struct TraceHit {
    int32_t poly;
    float impulse[3];
    float normal[3];
    float pos[3];
};

struct TraceBuffers {
    TraceQuery * queries;   // [n * substeps]
    uint32_t * mesh_ids;    // [n * substeps * cap_cand]
    uint32_t * caps_ids;    // [n * substeps * cap_cand]
    TraceHit * hits;        // [n * substeps * cap_hits]
    uint32_t cap_cand, cap_hits;
};

struct StepParams {
    // static geometry (replicated on every GPU)
    const float4 * nodes;           // 2 float4 per node, right-first preorder (bvh_builder.h FlatNode)
    uint32_t n_nodes;
    const float4 * tris;            // 3 float4 per triangle, leaf order: (a, n.x) (b, n.y) (c, n.z)
    const float4 * cull;            // 4 float4 per triangle: plane and edge planes of the pooled kernel's pre-cull (collide.cuh)
    float push_slack;               // k_step_pool: how far pushes may move a capsule before its window of triangles is culled again
    const uint32_t * node_leaf_id;  // mesh-collider id per node (trace only)
    const CapsuleDev * capsules;
    uint32_t n_capsules;
    // characters (in/out)
    prtc_transform * transforms;
    prtc_char_physics * physics;
    uint32_t n;
    // LITERAL order: capsule-capsule pass against the other characters' capsule segments
    const uint32_t * cap2char;      // capsule id -> character index (tagToCharacter, physics_system.cpp:275)
    const float4 * seg_start;       // 2 float4 per character: (A.xyz, radius) (B.xyz, 0) at the start-of-frame pose
    const float4 * seg_prev;        // same at the end-of-frame pose of the previous fixpoint round
    float * last_box;               // out, 6 floats per character: the last executed substep's query box
    // CANONICAL order, JACOBI dynamic pass: records of ALL characters (4 float4 each) + uniform xz grid over their stored boxes
    const float4 * dyn_records;
    const uint32_t * dyn_cell_start;  // [gx*gz + 1]
    const uint32_t * dyn_items;       // [n_global] record indices, cell by cell (every record in the cell of its stored box's lower corner)
    const uint32_t * dyn_reach;       // [2] how many cells (x, z) a stored box reaches beyond its own, maximum over all records this frame
    uint32_t dyn_first;             // global index of this launch's character 0
    uint32_t dyn_gx, dyn_gz;
    float dyn_origin_x, dyn_origin_z, dyn_inv_cell;
    // constants (prtc_config)
    float dt, gravity, ground_dot, friction, stick;
    uint32_t substeps, abs_mode;
    // conservative pre-culls: margin and enables (bit 0: box cull quick_reject, step_common.cuh; bit 1: plane/edge cull
    // tight_reject, collide.cuh, k_step_pool only), derived from the scene extent at commit
    float qr_eps;
    uint32_t qr_enable;
    uint32_t tune;                  // bit 0: frame cache OFF (walk the tree every substep, reference-equivalent counters)
    unsigned long long * counters;  // prtc_counters layout
    uint32_t * host_flags;          // host-mapped words the kernels raise (small host-buffer batches, resident steps) --
                                    // [0] a capsule-capsule test ran (LITERAL), [1] a character named an unknown capsule
    unsigned * work_counter;        // k_step_stream / k_step_pool: next character to hand out (null => lock-step kernel)
    // k_step_pool: two work counters used alternately -- a launch hands characters out of one and zeroes the OTHER for the
    // next launch (stream order makes that safe), so no memset is queued in front of every frame
    unsigned * work_pair;           // device: the two counters
    // chained frames (k_step_pool<CHAINED>, PRTC_OPT_CHAIN_FRAMES): a word per claim; this launch writes `chain_serial` into the
    // word of every claim it finishes and starts a claim only when its word holds `chain_prev` (0: first of a chain, no wait)
    uint32_t * chain_done; uint32_t chain_serial, chain_prev;
    unsigned * chain_ring;          // CHAIN_RING work counters used in rotation by the frames of a chain (see CHAIN_RING)
    struct ChainState * chain;      // HOST: what the previous launch of this context was; null => this launch may not chain
    unsigned long long * timeline;  // diagnostic (PRTC_TIMELINE=1): per claim of k_step_pool {start ns, end ns, SM}; null otherwise
    uint32_t * work_flip;           // HOST: which one the next launch uses (toggled by the launcher); null => memset + work_counter
    unsigned * work_next;           // device: the counter this launch zeroes (set by the launcher)
    // k_step_stream: cross-frame leaf cache, one slot per character index (SoA, stride xc_stride); null => off
    uint32_t * xc_version;          // tree version the slot was built for (0 = empty)
    float * xc_region;              // [6][stride] region the leaf list is complete for
    uint32_t * xc_count;            // [stride]
    uint32_t * xc_leaf;             // [LEAF_CAP][stride] node indices, emission order
    uint32_t xc_stride, xc_first;   // slot of this launch's character 0
    uint32_t tree_version;          // bumped on every re-flatten (>= 1)
    float xc_margin;                // how far beyond the frame's needed region a fresh list reaches
    uint32_t xc_reuse;              // 1: keep slots across frames (PRTC_OPT_CROSS_FRAME_CACHE), 0: rebuild every frame
    const float4 * top;             // k_xc_refresh: top-of-tree table staged in shared memory (2 float4 per entry), null => plain walk
    uint32_t n_top;
    uint32_t persistent_blocks;     // k_step_stream: resident blocks to launch
    uint32_t xc_fresh;              // 1: the per-slot leaf lists were brought up to date by the caller (launch_step does not launch k_xc_refresh)
    uint32_t pool;                  // 1: the pooled kernel k_step_pool runs the persistent launches (PRTC_OPT_POOL_KERNEL)
    uint32_t pool_blocks;           // k_step_pool: resident blocks to launch
    uint32_t n_tris;                // triangle records behind `tris`
    // k_step_stream fed by a host->device copy that is still in flight (prtc_step_characters on big host batches):
    const uint32_t * gate;          // characters [0, *gate) have arrived; advanced by the copy-in stream, polled by LOAD
    uint32_t * done_sub;            // finished characters per block of 2^INGEST_SUB_SHIFT characters
    uint32_t * done_chunk;          // finished blocks per chunk of 2^INGEST_CHUNK_SHIFT blocks: what the copy-out stream waits on
    uint32_t lanes_per_warp;        // k_step: characters per warp (set by launch_step; < 32 for small batches)
    // resident state (prtc_resident_step): per-frame input / output arrays, host-mapped or device memory; null => the records only
    const float * io_move;          // 3 floats per character: this frame's movementVector (replaces and updates physics[i].movement)
    float4 * io_result;             // per character: (position.xyz, isGrounded as 0.0f / 1.0f) after the frame
    // completion doorbell of a resident step: the last warp / block to finish writes `done_value` to host-mapped `done_flag`
    // (after a system-wide fence), so the host can spin on a word instead of sleeping in cudaStreamSynchronize
    uint32_t * done_flag;
    unsigned * done_count;          // device counter of finished warps / blocks (reset by the one that rings)
    uint32_t done_value;
    TraceBuffers trace;
};

constexpr uint32_t TOP_CAP = 1024;             // entries of the top-of-tree table k_xc_refresh stages in shared memory (32 KB)
constexpr uint32_t INGEST_SUB_SHIFT = 10;      // 1024 characters per completion counter (keeps the atomics spread out)
constexpr uint32_t INGEST_CHUNK_SHIFT = 6;     // 64 of those = 65536 characters per copy chunk (5.5 MB in, 5.5 MB out)

enum StepMode { STEP_STATIC = 0, STEP_LITERAL = 1, STEP_JACOBI = 2 };
int launch_step(StepParams const & p, bool traced, int mode, cudaStream_t stream);   // returns the number of kernels launched
uint32_t pool_blocks_per_sm();                  // resident one-warp blocks per SM k_step_pool is built for
// static pass over a batch whose input is still being copied in (gate / done_* set): the persistent kernel alone, then the
// refresh of the per-slot leaf lists for the NEXT frame from the state it leaves behind; returns the number of launches
int launch_step_ingest(StepParams const & p, cudaStream_t stream);
// the refresh of the per-slot leaf lists on its own (callers that want it ahead of something else set xc_fresh afterwards)
bool launch_slot_refresh(StepParams const & p, int mode, cudaStream_t stream);
// dynamic pass, CANONICAL / JACOBI: per-character record (4 float4) and the uniform grid over all stored boxes (k_dyn.cu)
// Where k_dyn_prep stores a record besides the device's own array, and how it tells the others that it is done:
//   records[k]   the other devices' record arrays (peer-mapped: cudaDeviceEnablePeerAccess inside one process, CUDA IPC across
//                processes)
//   flags[k]     (IPC exchange only) one word per destination -- this rank's slot in that device's arrival table; the LAST
//                block of the kernel writes `flag_value` (the frame number) into all of them behind a system-wide fence, and
//                the consumer's k_dyn_wait spins on its own table until every rank's word has reached the frame number
// Work counters of chained frames: frame f claims from counter f % CHAIN_RING and zeroes counter (f + 1) % CHAIN_RING before it lets
// frame f + 1 in.  That counter was last used by frame f + 1 - CHAIN_RING, which must be gone: a frame with an unfinished claim
// keeps every later frame unfinished too (their claim of the same characters waits), every unfinished frame holds at least one
// resident block, and the machine has room for 148 x 16 = 2368 one-warp blocks -- so no more than 2368 consecutive frames can be
// unfinished at the same time, however long a single claim takes.
constexpr uint32_t CHAIN_RING = 4096;

// host-side memory of chained launches (one per context)
struct ChainState {
    bool armed = false;             // the last step launched on the context was a chained frame with the signature below
    uint32_t serial = 0, ring = 0;
    unsigned long long sig[6] = {0, 0, 0, 0, 0, 0};
};

struct DynPeers {
    float4 * records[15]; uint32_t n;
    uint32_t * flags[16]; uint32_t n_flags; uint32_t flag_value; unsigned * block_count;
    uint32_t push_separately;       // 1: the records go out in a kernel of their own (k_dyn_push) behind k_dyn_prep
};
void launch_dyn_prep(const prtc_transform * tf, const prtc_char_physics * ph, const CapsuleDev * capsules, uint32_t n_capsules, uint32_t n,
                     uint32_t first, const float4 * prev_records, float4 * records, DynPeers const & peers, float dt, float gravity, float margin,
                     bool init, uint32_t * reach, cudaStream_t stream);
// IPC exchange: this rank's slice into every peer's array, then the arrival words (k_dyn_push)
void launch_dyn_push(const float4 * records, uint32_t first, uint32_t n, DynPeers const & peers, cudaStream_t stream);
// IPC exchange: wait (on the device, stream-ordered) until every rank's arrival word has reached `value`; gives up after a
// few seconds and raises counters[7] instead of hanging the GPU
void launch_dyn_wait(const uint32_t * flags, uint32_t n_ranks, uint32_t value, unsigned long long * counters, cudaStream_t stream);
// counting sort of all n stored boxes into the uniform xz grid (one cell per box) + the per-frame reach; returns kernels launched
int launch_dyn_grid_build(const float4 * records, uint32_t n, float ox, float oz, float inv_cell, uint32_t gx, uint32_t gz,
                          uint32_t * cell_count, uint32_t * cell_start, uint32_t * cursor, uint32_t * items, uint32_t * reach,
                          unsigned long long * scan_state, uint32_t & scan_tickets, uint32_t & scan_launches, cudaStream_t stream);
constexpr uint32_t DYN_SCAN_WORDS = 1024;           // look-back words of k_dyn_scan (2^20 cells / 2048 per tile + 1, rounded up); the ticket follows

// main[k] += scratch[k], scratch[k] = 0 for the device-side work counters (per-frame scratch of the LITERAL fixpoint rounds)
void launch_fold_counters(unsigned long long * main_counters, unsigned long long * scratch, cudaStream_t stream);
// device self-test of the written-out IEEE division / square-root fast paths (16 operand sets per thread)
void launch_selftest_math(uint64_t seed, uint32_t n, int mode, unsigned long long * bad, cudaStream_t stream);

// K8: world-space vertex cache + per-mesh AABB (physics_system.cpp:68-93 / :161-202)
struct MeshDev { uint32_t first_vertex, n_vertices, model; uint32_t pad; };
struct ModelXform { Affine m; };
void launch_transform_refit(const float * raw_xyz, float * world_xyz, const MeshDev * meshes, uint32_t first_mesh,
                            uint32_t n_meshes, const ModelXform * xforms, float * mesh_aabbs /*6 per mesh*/,
                            cudaStream_t stream);
// leaf-ordered triangle records with precomputed unit normals (3 float4 each) and their cull records (4 float4 each)
void launch_tri_setup(const float * world_xyz, const uint32_t * tri_src_vertex, uint32_t n_tris, float4 * tris, float4 * cull,
                      cudaStream_t stream);
// PRTC_ORDER_BY_ID: left-complete tree over n_leaves leaf records (2 float4 each, ascending id) built into nodes (k_tree_build.cu)
void launch_tree_build(const float4 * leaves, uint32_t n_leaves, float4 * nodes, cudaStream_t stream);
// K7: batched PhysicsSystem::raycast
void launch_raycast(const float4 * nodes, uint32_t n_nodes, const float4 * tris, uint32_t n_rays, const float * origins,
                    const float * dirs, const float * maxd, float * hit_xyz, unsigned char * hit_mask, cudaStream_t stream,
                    uint32_t * done_flag = nullptr, unsigned * done_count = nullptr, uint32_t done_value = 0);
bool raycast_rings_doorbell(uint32_t n_nodes, uint32_t n_rays);      // the launch above will write done_flag when it finishes
// TIER B (parity unpinned, see k_sweep.cu): swept ellipsoids with collide-and-slide, one warp per ellipsoid
struct SweepParams {
    const float4 * nodes; uint32_t n_nodes;
    const float4 * tris;
    const uint32_t * tri_src;       // per triangle record: index of its first vertex in registration order
    uint32_t n, iterations;
    float very_close;
    const float * centers, * radii, * velocities;       // 3 floats per ellipsoid, world space
    float * out_pos, * out_vel;
    int32_t * hit_tri; float * hit_toi, * hit_normal;    // per (ellipsoid, iteration)
    unsigned long long * counters;
};
void launch_sweep(SweepParams const & p, cudaStream_t stream);
// boxes of the staged top-of-tree table re-read from the node array (topology unchanged)
void launch_top_refresh(const float4 * nodes, float4 * top, const uint32_t * src, uint32_t n_top, cudaStream_t stream);
// single box query (introspection / tests)
void launch_query_box(const float4 * nodes, uint32_t n_nodes, const uint32_t * node_leaf_id, const float * aabb6,
                      uint32_t * out_ids, uint32_t capacity, uint32_t * out_count, cudaStream_t stream);

} // namespace prt
#endif
