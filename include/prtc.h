/*
 * prtc.h -- C-ABI of the B200-native character-collision path ("prtc" = PRoTotype2 Collision).
 *
 * This is the drop-in boundary for the one hot path of ArneStenkrona/prototype2 that this repository
 * accelerates: PhysicsSystem::updateCharacters and what it calls (broadphase AABB-tree query,
 * capsule-vs-triangle narrowphase with push-out response, capsule-vs-capsule pass).  The reference has no
 * FFI layer today -- the C++ class PhysicsSystem *is* the seam (reference
 * src/game/system/physics/physics_system.h:24-77).  Every entry point below names the PhysicsSystem member
 * it stands behind; the C++ drop-in class that keeps the reference's signatures and forwards to these
 * functions lives in prototype2_b200/engine/physics_system.h, and INTEGRATION.md shows the binding.
 *
 * Conventions: extern "C", opaque context handle, int status (0 = PRTC_OK), no exceptions cross the
 * boundary, caller-owned buffers, plain pointers and sizes.  All calls on one context must come from one
 * host thread (the reference is single-threaded).  There is NO CPU fallback: every compute
 * entry point fails with PRTC_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef PRTC_H
#define PRTC_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PRTC_VERSION 1

enum prtc_status {
    PRTC_OK = 0,
    PRTC_ERR_INVALID = 1,   /* bad argument / call order */
    PRTC_ERR_CUDA = 2,      /* CUDA runtime error or no usable device; see prtc_last_error() */
    PRTC_ERR_CAPACITY = 3,  /* a caller-provided buffer was too small */
    PRTC_ERR_ALLOC = 4
};

/* How the unqualified `abs(dist)` at reference collision_system.cpp:72 is resolved (it decides which
 * triangles survive the range cull and therefore isGrounded):
 *   PRTC_ABS_INT    libstdc++ builds: only `int abs(int)` is visible -> dist is truncated to int first
 *                   (this is what the reference does when compiled on this image; default)
 *   PRTC_ABS_FLOAT  libc++ / MSVC builds: the float overload is chosen */
enum prtc_abs_mode { PRTC_ABS_INT = 0, PRTC_ABS_FLOAT = 1 };

/* Candidate order (SURVEY.md findings 3, H3, H4):
 *   PRTC_ORDER_CANONICAL  static-only tree built by the reference insertion algorithm, then frozen;
 *                         capsule neighbours in ascending index order.  Scales to millions of capsules.
 *   PRTC_ORDER_LITERAL    the reference's unified mesh+capsule tree is maintained on the host by the same
 *                         algorithm every frame and re-flattened, and characters keep the reference's
 *                         sequential visibility of one another; for engine-sized scenes.
 *   PRTC_ORDER_BY_ID      like CANONICAL, but the mesh candidates of a query -- the same SET, which depends only on
 *                         the leaves' stored fat boxes, not on the tree -- are taken in ascending collider id
 *                         (registration order) instead of the insertion-history order of the reference's tree.  No
 *                         tree is built on the host: the device puts a left-complete tree over the id sequence
 *                         together in closed form (commit of 1.25 M leaves in milliseconds instead of seconds; a
 *                         moved model costs one bottom-up pass).  For streamed / very large levels. */
enum prtc_order_mode { PRTC_ORDER_LITERAL = 0, PRTC_ORDER_CANONICAL = 1, PRTC_ORDER_BY_ID = 2 };

/* Dynamic-vs-dynamic (capsule-capsule) pass in CANONICAL order:
 *   PRTC_DYN_OFF     capsules do not see one another (static pass only)
 *   PRTC_DYN_JACOBI  every capsule collides against the start-of-frame pose of its neighbours (see prtc_dyn_*) */
enum prtc_dyn_mode { PRTC_DYN_OFF = 0, PRTC_DYN_JACOBI = 2 };

/* Compile-time constants of the reference turned into fields (defaults = reference values). */
typedef struct prtc_config {
    float gravity;         /* 1.0    physics_system.h:122 (m_gravity) */
    float leaf_margin;     /* 0.05   aabb_tree.h:90 (buffer) */
    float ground_dot;      /* 0.3    collision_system.cpp:245 */
    float friction;        /* 10.0   physics_system.cpp:307 */
    float stick;           /* 0.05   physics_system.cpp:280 */
    uint32_t substeps;     /* 4      physics_system.cpp:346 (nTimeSteps) */
    uint32_t abs_mode;     /* enum prtc_abs_mode */
    uint32_t order_mode;   /* enum prtc_order_mode */
    uint32_t dyn_mode;     /* enum prtc_dyn_mode */
    int32_t device;        /* CUDA device ordinal */
} prtc_config;

/* Bit-compatible with the reference's `Transform` (src/game/component/component.h:14-17):
 * glm::vec3 position; glm::quat rotation (glm 0.9.9 stores x,y,z,w); glm::vec3 scale.  40 bytes. */
typedef struct prtc_transform {
    float position[3];
    float rotation[4];     /* x, y, z, w */
    float scale[3];
} prtc_transform;

/* The fields of the reference's `CharacterPhysics` (src/game/system/character/character.h:19-27) that
 * updateCharacters reads or writes, with the collider index widened to 32 bits.  44 bytes. */
typedef struct prtc_char_physics {
    float velocity[3];         /* in/out */
    float movement[3];         /* in      movementVector */
    float ground_normal[3];    /* in/out  groundNormal */
    uint32_t collider;         /* in      colliderTag.index (capsule id from prtc_add_capsule); an id that was never
                                *         added is not dereferenced: the step returns PRTC_ERR_INVALID and that
                                *         character's result is undefined */
    uint32_t grounded;         /* in/out  isGrounded (0/1) */
} prtc_char_physics;

/* Work counters accumulated on the device since the last reset. */
typedef struct prtc_counters {
    uint64_t substeps;     /* executed query-substeps (one broadphase query each) */
    uint64_t tri_tests;    /* triangles entering the narrowphase loop (collision_system.cpp:29) */
    uint64_t node_tests;   /* BVH nodes tested by the GPU traversal */
    uint64_t contacts;     /* handleCollision calls (collision_system.cpp:236) */
    uint64_t cc_tests;     /* capsule-capsule narrowphase tests */
    uint64_t exact_tests;  /* triangles that survived the pre-cull and ran the exact capsule-triangle test */
    uint64_t launches;     /* kernels launched by the step entry points (host-side count) */
} prtc_counters;

/* Parity trace of one traced step (prtc_step_characters_traced); same record layout the CPU checkers under tests/ use.
 * Arrays are caller-allocated; capacities in, counts out.  Queries are sorted by (character, substep). */
typedef struct prtc_trace {
    uint32_t cap_queries, cap_mesh_ids, cap_caps_ids, cap_hits;     /* in  */
    uint32_t n_queries, n_mesh_ids, n_caps_ids, n_hits;              /* out */
    uint32_t * q_char;        /* [cap_queries]      character index */
    uint32_t * q_sub;         /* [cap_queries]      substep 0.. */
    float * q_aabb;           /* [cap_queries*6]    query box lo.xyz hi.xyz */
    uint32_t * q_mesh_off;    /* [cap_queries+1]    CSR offsets into mesh_ids */
    uint32_t * q_caps_off;    /* [cap_queries+1]    CSR offsets into caps_ids */
    uint32_t * mesh_ids;      /* [cap_mesh_ids]     candidate mesh colliders, in test order */
    uint32_t * caps_ids;      /* [cap_caps_ids]     candidate capsules, in test order */
    uint32_t * q_tri_tests;   /* [cap_queries] */
    uint32_t * h_query;       /* [cap_hits]         index into q_* */
    int32_t * h_poly;         /* [cap_hits]         triangle index within the query's candidate list, -1 = capsule */
    float * h_impulse;        /* [cap_hits*3] */
    float * h_normal;         /* [cap_hits*3] */
    float * h_pos;            /* [cap_hits*3]       position right after the push */
} prtc_trace;

typedef struct prtc_ctx prtc_ctx;

void prtc_default_config(prtc_config * cfg);
const char * prtc_last_error(void);
/* number of CUDA devices with compute capability 10.x; 0 when there is none (never an error) */
int prtc_device_count(void);

/* PhysicsSystem::PhysicsSystem()  physics_system.cpp:15-16 */
int prtc_create(const prtc_config * cfg, prtc_ctx ** out);
void prtc_destroy(prtc_ctx * ctx);

/* PhysicsSystem::addModelCollider(Model const&, Transform const&)  physics_system.cpp:44-106.
 * xyz: 3*n_vertices floats, index-expanded model-space positions (what the reference copies out of
 * Model::vertexBuffer via indexBuffer, :85); mesh_first/mesh_count: one entry per Model::Mesh, in vertices
 * (multiples of 3).  One mesh collider (= one BVH leaf) is created per mesh; ids are consecutive. */
int prtc_add_model(prtc_ctx * ctx, const float * xyz, uint32_t n_vertices, const uint32_t * mesh_first,
                   const uint32_t * mesh_count, uint32_t n_meshes, const prtc_transform * transform,
                   uint32_t * out_model_id, uint32_t * out_first_mesh_id);

/* PhysicsSystem::addCapsuleCollider(float h, float r, vec3 const& offset)  physics_system.cpp:22-42 */
int prtc_add_capsule(prtc_ctx * ctx, float height, float radius, const float offset[3], uint32_t * out_id);
/* PhysicsSystem::updateCapsuleCollider  physics_system.cpp:151-159 */
int prtc_update_capsule(prtc_ctx * ctx, uint32_t id, float height, float radius, const float offset[3]);

/* Makes registered colliders visible to the device: transforms vertices to world space, builds / extends
 * the AABB tree with the reference's insertion algorithm (aabb_tree.cpp:134-180,243-339), flattens it in
 * the reference's right-first DFS order and uploads.  Called implicitly by the step functions when needed. */
int prtc_commit(prtc_ctx * ctx);

/* PhysicsSystem::removeCollider(tag) for a MODEL tag -> removeModelCollider  physics_system.cpp:108-149: the model's
 * mesh leaves leave the tree.  (Ids of other colliders stay stable here; the reference renumbers later meshes.) */
int prtc_remove_model(prtc_ctx * ctx, uint32_t model_id);

/* PhysicsSystem::updateModelColliders(ColliderTag const*, Transform const*, size_t)  physics_system.cpp:161-202 */
int prtc_update_models(prtc_ctx * ctx, const uint32_t * model_ids, const prtc_transform * transforms, uint32_t n);

/* PhysicsSystem::updateCharacters(float dt, Scene&, Transform*)  physics_system.cpp:245-318.
 * HOST buffers, updated in place; copies to and from the device are part of the call. */
int prtc_step_characters(prtc_ctx * ctx, float dt, uint32_t n, prtc_transform * transforms,
                         prtc_char_physics * physics);
/* Same step on DEVICE-resident arrays (asynchronous on the context's stream). */
int prtc_step_characters_device(prtc_ctx * ctx, float dt, uint32_t n, prtc_transform * d_transforms,
                                prtc_char_physics * d_physics);
/* Host-buffer step that also records the parity trace (slower; for tests). */
int prtc_step_characters_traced(prtc_ctx * ctx, float dt, uint32_t n, prtc_transform * transforms,
                                prtc_char_physics * physics, prtc_trace * trace);

/* RESIDENT STATE.  The engine's caller gathers the characters' transforms, calls updateCharacters and scatters the positions
 * back (CharacterSystem::updatePhysics, character_system.cpp:55-78); between frames it only touches movementVector (input /
 * AI, character_system.cpp:20-35) and, now and then, a velocity (jump, :92) or a transform (teleport).  With resident state
 * the records stay on the device:
 *   prtc_resident_begin   uploads n records once (CANONICAL / BY_ID order; LITERAL keeps its host records)
 *   prtc_resident_io      two host arrays owned by the context (pinned, mapped into the device's address space):
 *                         movement[3n]  IN   this frame's movementVector per character -- the caller writes it before a step
 *                         result[4n]    OUT  position.xyz and isGrounded (0.0f / 1.0f) per character -- valid when a step returns
 *   prtc_resident_step    one frame (updateCharacters, physics_system.cpp:245-318): the kernels read movement[] and write
 *                         result[] in place over the bus while they run -- 12 + 16 bytes per character and frame instead of
 *                         84 + 84, no separate copies; returns after the frame is complete
 *   prtc_resident_edit    overwrites k records (transform and / or physics; either may be null) -- jumps, teleports, spawns
 *   prtc_resident_fetch   downloads all records (either pointer may be null)
 * Same results, bit for bit, as prtc_step_characters on the same records with physics[i].movement = movement[3i..]. */
int prtc_resident_begin(prtc_ctx * ctx, uint32_t n, const prtc_transform * transforms, const prtc_char_physics * physics);
int prtc_resident_io(prtc_ctx * ctx, float ** movement, float ** result);
int prtc_resident_step(prtc_ctx * ctx, float dt);
int prtc_resident_edit(prtc_ctx * ctx, uint32_t k, const uint32_t * indices, const prtc_transform * transforms,
                       const prtc_char_physics * physics);
int prtc_resident_fetch(prtc_ctx * ctx, prtc_transform * transforms, prtc_char_physics * physics);
int prtc_resident_end(prtc_ctx * ctx);

/* SEVERAL GPUs OF ONE NODE BEHIND ONE HANDLE.  One process, one context per device; static geometry and its
 * tree are replicated (every registration call is repeated on every device, the ids come out the same), the characters of a
 * call are partitioned into contiguous slices, device k steps slice k on its own stream.  The static pass needs no exchange.
 * The dynamic-vs-dynamic pass (PRTC_DYN_JACOBI) has one per frame: every device's record kernel stores its characters' records
 * straight into the other devices' record arrays (peer stores over NVLink; cudaMemcpyPeerAsync where there is no peer access).
 * Results are byte-identical to one context stepping the whole batch.  PRTC_ORDER_CANONICAL / PRTC_ORDER_BY_ID (LITERAL order
 * keeps the reference's sequential visibility between characters and does not shard).  `cfg->device` is ignored.
 * prtc_multi_context() exposes the per-device contexts (options, counters, introspection). */
typedef struct prtc_multi prtc_multi;
int prtc_multi_create(const prtc_config * cfg, const int32_t * devices, uint32_t n_devices, prtc_multi ** out);
void prtc_multi_destroy(prtc_multi * m);
uint32_t prtc_multi_device_count(const prtc_multi * m);
prtc_ctx * prtc_multi_context(prtc_multi * m, uint32_t k);
int prtc_multi_add_model(prtc_multi * m, const float * xyz, uint32_t n_vertices, const uint32_t * mesh_first, const uint32_t * mesh_count,
                         uint32_t n_meshes, const prtc_transform * transform, uint32_t * out_model_id, uint32_t * out_first_mesh_id);
int prtc_multi_add_capsule(prtc_multi * m, float height, float radius, const float offset[3], uint32_t * out_id);
int prtc_multi_update_capsule(prtc_multi * m, uint32_t id, float height, float radius, const float offset[3]);
int prtc_multi_remove_model(prtc_multi * m, uint32_t model_id);
int prtc_multi_update_models(prtc_multi * m, const uint32_t * model_ids, const prtc_transform * transforms, uint32_t n);
int prtc_multi_commit(prtc_multi * m);
int prtc_multi_set_option(prtc_multi * m, int option, int value);
int prtc_multi_raycast(prtc_multi * m, uint32_t n, const float * origins, const float * directions, const float * max_distances,
                       float * hit_xyz, uint8_t * hit_mask);
int prtc_multi_get_counters(prtc_multi * m, prtc_counters * out, int reset);          /* summed over the devices */
/* PhysicsSystem::updateCharacters, HOST buffers, over all devices */
int prtc_multi_step_characters(prtc_multi * m, float dt, uint32_t n, prtc_transform * transforms, prtc_char_physics * physics);
/* resident state over all devices: ONE movement[3n] and ONE result[4n] host array, slice k read / written by device k */
int prtc_multi_resident_begin(prtc_multi * m, uint32_t n, const prtc_transform * transforms, const prtc_char_physics * physics);
int prtc_multi_resident_io(prtc_multi * m, float ** movement, float ** result);
int prtc_multi_resident_step(prtc_multi * m, float dt);
int prtc_multi_resident_fetch(prtc_multi * m, prtc_transform * transforms, prtc_char_physics * physics);
int prtc_multi_resident_end(prtc_multi * m);

/* PhysicsSystem::raycast(origin, direction, maxDistance, hit&)  physics_system.cpp:205-242, batched:
 * n rays; hit_xyz[3*i..] is written and hit_mask[i]=1 where ray i hits a mesh collider. */
int prtc_raycast(prtc_ctx * ctx, uint32_t n, const float * origins, const float * directions,
                 const float * max_distances, float * hit_xyz, uint8_t * hit_mask);

/* TIER B -- PARITY UNPINNED.  The swept ellipsoid-vs-triangle collide-and-slide of the engine's older collision code:
 * n ellipsoids (centre, radius vector, displacement of this frame; world space) against the committed mesh colliders,
 * `iterations` rounds of { earliest time of impact over all candidate triangles in ellipsoid space; stop very_close
 * short of it; project the rest of the motion onto the slide plane }.  The snapshot of the reference has no caller
 * for this any more (PhysicsSystem::collisionResponse is declared, physics_system.h:131-136, never defined); its
 * helper functions survive in src/util/physics_util.cpp:51-178 and are what the kernel restates.  PRTC_ORDER_CANONICAL
 * contexts only.  Outputs: out_pos / out_vel (3 floats per ellipsoid: where it ends up, what is left of the motion),
 * and per (ellipsoid, iteration): hit_tri (triangle index in registration order, -1 = none), hit_toi (fraction of
 * that iteration's motion), hit_normal (slide-plane normal in ellipsoid space, 3 floats). */
int prtc_sweep_ellipsoids(prtc_ctx * ctx, uint32_t n, const float * centers, const float * radii, const float * velocities,
                          uint32_t iterations, float very_close, float * out_pos, float * out_vel, int32_t * hit_tri,
                          float * hit_toi, float * hit_normal);

/* Dynamic-vs-dynamic pass at scale (PRTC_ORDER_CANONICAL + PRTC_DYN_JACOBI): every character
 * collides with the start-of-frame pose of every other character whose STORED fat box (the hysteresis of
 * DynamicAABBTree::update / insertLeaf, aabb_tree.cpp:102-111,140-141) overlaps its query box, in ascending global
 * character index (collideCapsuleCapsule, collision_system.cpp:158-234; pass order physics_system.cpp:369-389).
 * On one GPU nothing extra has to be called.  With characters partitioned over ranks, per frame and rank:
 *   prtc_dyn_configure(ctx, first, n_global)             once: this rank owns global indices [first, first + n)
 *   prtc_dyn_prepare(ctx, dt, n, d_transforms, d_physics) writes this rank's records (64 bytes per character)
 *   prtc_dyn_buffer(...) -> the caller all-gathers the record array IN PLACE (e.g. ncclAllGather over NVLink)
 *   prtc_step_characters_device(...)                      grid over all records + the step                       */
int prtc_dyn_configure(prtc_ctx * ctx, uint32_t first_global, uint32_t n_global);
int prtc_dyn_prepare(prtc_ctx * ctx, float dt, uint32_t n, const prtc_transform * d_transforms, const prtc_char_physics * d_physics);
int prtc_dyn_buffer(prtc_ctx * ctx, void ** d_records, size_t * record_bytes, uint32_t * first_global, uint32_t * n_global);
/* The same exchange ACROSS PROCESSES without a collective: every rank exports CUDA IPC handles of its record array and of
 * its arrival table (PRTC_IPC_HANDLE_BYTES bytes), the caller distributes them (any transport), every rank imports all of
 * them in rank order.  From then on prtc_dyn_prepare stores this rank's records straight into every peer's array (peer
 * stores over NVLink from the kernel that computes them; the array is double-buffered) and its last block writes the frame
 * number into the peers' arrival tables; prtc_step_characters_device waits on the device until every rank has arrived
 * (a rank that does not arrive within a few seconds makes the step fail with PRTC_ERR_CUDA instead of hanging the GPU).
 * No all-gather, no host involvement per frame.  Call export after prtc_dyn_configure and before the first frame. */
#define PRTC_IPC_HANDLE_BYTES 128
int prtc_dyn_ipc_export(prtc_ctx * ctx, uint32_t rank, uint32_t world, void * handle /* PRTC_IPC_HANDLE_BYTES */);
int prtc_dyn_ipc_import(prtc_ctx * ctx, const void * handles /* world x PRTC_IPC_HANDLE_BYTES, rank order; NULL: give the exchange up again */);

/* stream control for callers that own device memory (bench / multi-GPU drivers) */
int prtc_set_stream(prtc_ctx * ctx, void * cuda_stream);
int prtc_synchronize(prtc_ctx * ctx);

int prtc_get_counters(prtc_ctx * ctx, prtc_counters * out, int reset);

/* Work-saving devices and kernel choices of the GPU path.  None changes any result bit; they change how much work is
 * done and how it is scheduled.
 *   PRTC_OPT_FRAME_CACHE   (default 1) one tree walk per character and frame with a box covering the whole
 *                          frame's motion; substeps re-test only the cached leaves.  With 0 the device walks the
 *                          tree once per substep exactly like DynamicAABBTree::query, and prtc_counters.node_tests
 *                          then equals the reference's count (the numerator of the roofline's algorithmic bytes).
 *   PRTC_OPT_QUICK_REJECT  (default 1) conservative per-triangle pre-cull ahead of the exact test.
 *   PRTC_OPT_PERSISTENT    (default 1) static pass runs as a persistent kernel whose lanes advance through their
 *                          characters independently (0: one character per lane in lock step).
 *   PRTC_OPT_CROSS_FRAME_CACHE (default 1) the persistent kernel keeps every character slot's leaf list (valid for
 *                          an enlarged region around it) in device memory between frames; a character whose
 *                          frame still fits inside that region skips the tree walk.  The list is a pure function
 *                          of (tree, region), so it does not matter which character occupies the slot.
 *   PRTC_OPT_WARP_KERNEL   (default 1) batches small enough to give every character a resident warp run on the
 *                          warp-cooperative kernel (0: the lock-step kernel with thinned-out warps).
 *   PRTC_OPT_STREAMED_COPY (default 1) prtc_step_characters on a big static batch (>= 262144 characters) runs ONE
 *                          persistent launch that is fed by the host->device copy while it runs and hands finished
 *                          chunks to the device->host copy while it is still computing (0: the batch is cut into
 *                          chunks that are copied in, stepped and copied out on a few streams).
 *   PRTC_OPT_TOP_STAGING   (default 1) the kernel that refills the per-slot leaf lists walks the top levels of the tree
 *                          (<= 1024 nodes) out of shared memory, staged per block with one TMA bulk copy (0: every
 *                          node comes from global memory).  CANONICAL order, trees of >= 4096 nodes.
 *   PRTC_OPT_POOL_KERNEL   (default 1) persistent launches run the pooled kernel: a warp owns 32 characters and all its lanes
 *                          work through the pooled candidate triangles of those characters (dense cull, dense exact
 *                          test, speculative ordered commit); 0: the lane-per-character kernel of round 1.
 *   PRTC_OPT_TIGHT_REJECT  (default 1) pooled kernel only: a second conservative cull (plane and edge-plane distances of
 *                          the capsule segment) behind the box cull.
 *   PRTC_OPT_CHAIN_FRAMES  (default 0) prtc_step_characters_device, static pass, batches of ~20 000 characters and more: consecutive
 *                          calls with the same arguments may OVERLAP on the device -- the next frame's thread blocks are
 *                          scheduled as the current frame's finish (programmatic dependent launch) and a character's next
 *                          frame starts when its own current frame has been stored, so the idle tail of one frame is filled
 *                          with the head of the next.  Results are the same bytes; every call still completes in stream
 *                          order for anything the caller enqueues AFTER it.  Contract: between two chained calls the caller
 *                          must not enqueue work on the context's stream that WRITES the records (reads are fine, and so is
 *                          any host-side synchronisation); any other libprtc call breaks the chain by itself. */
enum prtc_option { PRTC_OPT_FRAME_CACHE = 1, PRTC_OPT_QUICK_REJECT = 2, PRTC_OPT_PERSISTENT = 3, PRTC_OPT_CROSS_FRAME_CACHE = 4,
                   PRTC_OPT_WARP_KERNEL = 5, PRTC_OPT_STREAMED_COPY = 6, PRTC_OPT_TOP_STAGING = 7, PRTC_OPT_POOL_KERNEL = 8,
                   PRTC_OPT_TIGHT_REJECT = 9, PRTC_OPT_CHAIN_FRAMES = 10 };
int prtc_set_option(prtc_ctx * ctx, int option, int value);

/* Introspection for parity tests: the flattened tree.  leaf_mesh_ids receives the mesh-collider ids of all
 * leaves in traversal (= reference DFS) order. */
int prtc_get_tree_info(prtc_ctx * ctx, uint32_t * n_nodes, uint32_t * n_leaves, uint32_t * n_triangles);
int prtc_get_leaf_order(prtc_ctx * ctx, uint32_t * leaf_mesh_ids, uint32_t capacity);
/* How the commits since prtc_create were done: `incremental` = moved models only (their transforms, vertices, boxes and
 * triangle records; the tree is left alone or, in BY_ID order, put together again on the device), `full` = the tree was
 * re-flattened / rebuilt and uploaded. */
int prtc_get_commit_stats(prtc_ctx * ctx, uint32_t * incremental, uint32_t * full);
/* Diagnostic, only in the diagnostic build of the library (make timeline) with PRTC_TIMELINE=1 in the environment at prtc_create: per claim (a warp's group of characters) of the
 * last k_step_pool launch three 64-bit words {start ns, end ns, SM} (%globaltimer), zero where no claim was made.
 * `*n_words` = words available; `out` may be null to ask for the size. */
int prtc_debug_timeline(prtc_ctx * ctx, unsigned long long * out, size_t capacity_words, size_t * n_words);
/* world-space vertex cache as the device computed it (physics_system.cpp:86), 3 floats per vertex */
int prtc_get_world_vertices(prtc_ctx * ctx, float * xyz, uint32_t capacity_vertices, uint32_t * n_vertices);
/* candidate mesh colliders of one box, in test order (DynamicAABBTree::query, aabb_tree.cpp:34-68), on the device */
int prtc_query_box(prtc_ctx * ctx, const float aabb6[6], uint32_t * mesh_ids, uint32_t capacity, uint32_t * n_found);


/* Device self-test (tests only): the kernels compute some IEEE divisions and square roots through written-out copies of
 * the fast paths of div.rn.f32 / sqrt.rn.f32 that share a reciprocal or interleave independent chains (csrc/prt_math.cuh).
 * Runs n_threads x 16 pseudo-random operand sets through them and through __fdiv_rn / __fsqrt_rn and returns how many
 * results differ in their bits (NaN == NaN).  mode 0: random bit patterns; 1: magnitudes around 1; 2 / 3: magnitudes at the
 * edges of the fast paths' operand ranges (2^+-60 for the divisions, 2^+-100 for the square roots).  Must return 0. */
int prtc_selftest_math(prtc_ctx * ctx, uint32_t n_threads, uint64_t seed, int mode, uint64_t * mismatches);

/* Host-only utility (needs no device, performs no collision arithmetic): builds the reference's dynamic AABB
 * tree over n boxes inserted in order (DynamicAABBTree::insert, aabb_tree.cpp:95-100) and returns it in the
 * flattened device layout: nodes8 = 8 x 32-bit words per node {lo.xyz, link, hi.xyz, count} in right-first
 * preorder (internal: link = skip index, count = -(second child index); leaf: link = box index, count = 1), leaf_order = box
 * indices in traversal order.  Lets the build / flatten logic be checked on a CPU-only box by the tests. */
int prtc_host_build_tree(const float * aabbs6, uint32_t n, float margin, uint32_t * leaf_order, float * nodes8,
                         uint32_t nodes_capacity, uint32_t * n_nodes);
/* Host-only utility: the tree PRTC_ORDER_BY_ID builds on the device, computed by the same index arithmetic
 * (csrc/balanced_tree.cuh) in plain loops: the left-complete binary tree over n boxes (taken as they are, no margin added)
 * in the flattened layout above, leaf: link = box index, count = 1.  A front-to-back walk emits overlapping boxes in
 * ascending index. */
int prtc_host_balanced_tree(const float * aabbs6, uint32_t n, float * nodes8, uint32_t nodes_capacity, uint32_t * n_nodes);
/* Host-only utility: the top-of-tree table the device stages in shared memory (PRTC_OPT_TOP_STAGING) for a flattened
 * tree in the layout above: all nodes of depth <= D (D the largest depth whose levels together fit `cap` entries), same
 * preorder, 8 words per entry {lo.xyz, link, hi.xyz, count} with
 *   internal above the cut: link = table index to go on at when the box is missed, count = -1;
 *   internal at the cut:    link = index g of the node in the full array, count = -(skip index of g);
 *   leaf:                   link = index g of the node in the full array, count = 0.
 * n_top = 0 when the tree has fewer than 4 * cap nodes (no staging). */
int prtc_host_top_table(const float * nodes8, uint32_t n_nodes, uint32_t cap, float * top8, uint32_t top_capacity, uint32_t * n_top);

#ifdef __cplusplus
}
#endif
#endif /* PRTC_H */
