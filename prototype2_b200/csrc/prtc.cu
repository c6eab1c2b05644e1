// prtc.cu -- host side of the C-ABI declared in include/prtc.h, part 1: context, collider registry, commit (vertex
// transform on the device, reference-algorithm tree build on the host, flatten, upload), options and introspection.
// The per-frame entry points are in prtc_step.cu.  There is no CPU implementation of the compute path in this
// library: without a compute-capability-10 device prtc_create fails and nothing else can be called.
#include "prtc_internal.h"
#include "balanced_tree.cuh"

using namespace prt;
using namespace prt::host;

namespace prt {
namespace host {

namespace { thread_local std::string g_last_error; }

int fail(int code, const std::string & msg) {
    g_last_error = msg;
    return code;
}

int build_tree_on_device(prtc_ctx * c, cudaStream_t s);

int commit_impl(prtc_ctx * c) {
    c->chain.armed = false;
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

    // ---- moved models only (prtc_update_models since the last commit, nothing added or removed): INCREMENTAL -------------
    // Only those models' transforms go up, only their meshes are re-transformed (K8) and only their boxes come back.  If every
    // new box still lies inside its leaf's stored fat box (DynamicAABBTree::update keeps the leaf then, aabb_tree.cpp:102-111)
    // the tree is untouched: the triangle / cull records of just those meshes are rebuilt and that is all -- no re-flatten, no
    // upload of the level, the per-slot leaf lists stay valid.  Otherwise: CANONICAL order re-inserts the leaves in the
    // host tree and re-flattens (the shape -- hence the candidate order -- changed; the vertices are not uploaded again);
    // BY_ID order uploads the changed leaf records and lets the device put the tree together again (same topology: the staged
    // top-of-tree table only refreshes its boxes).
    if (c->geometry_dirty && !c->registry_dirty && !c->literal() && !c->moved_models.empty() && c->meshes_on_device == n_meshes &&
        c->h_aabbs.size() == size_t(n_meshes) * 6 && !c->tree_dirty) {
        std::sort(c->moved_models.begin(), c->moved_models.end());
        c->moved_models.erase(std::unique(c->moved_models.begin(), c->moved_models.end()), c->moved_models.end());
        size_t moved_meshes = 0;
        for (uint32_t model : c->moved_models) moved_meshes += c->models[model].n_meshes;
        if (moved_meshes <= 8192) {
            std::vector<ModelXform> xf(c->moved_models.size());
            for (size_t k = 0; k < c->moved_models.size(); ++k) {
                const uint32_t model = c->moved_models[k];
                ModelHost const & mh = c->models[model];
                xf[k].m = affine_from(mh.transform);
                PRTC_CUDA(cudaMemcpyAsync(c->d_xforms.ptr + model, &xf[k], sizeof(ModelXform), cudaMemcpyHostToDevice, s));
                if (!mh.n_meshes) continue;
                launch_transform_refit(c->d_raw.ptr, c->d_world.ptr, c->d_meshes.ptr, mh.first_mesh, mh.n_meshes, c->d_xforms.ptr, c->d_aabbs.ptr, s);
                PRTC_CUDA(cudaMemcpyAsync(c->h_aabbs.data() + 6 * size_t(mh.first_mesh), c->d_aabbs.ptr + 6 * size_t(mh.first_mesh),
                                          size_t(mh.n_meshes) * 6 * sizeof(float), cudaMemcpyDeviceToHost, s));
            }
            PRTC_CUDA(cudaGetLastError());
            PRTC_CUDA(cudaStreamSynchronize(s));
            bool changed = false;
            std::vector<uint32_t> refit;                           // BY_ID: meshes whose stored fat box changed
            for (uint32_t model : c->moved_models) {
                ModelHost const & mh = c->models[model];
                for (uint32_t m = mh.first_mesh; m < mh.first_mesh + mh.n_meshes; ++m) {
                    const float * a = c->h_aabbs.data() + 6 * size_t(m);
                    Box b; b.lo = v3(a[0], a[1], a[2]); b.hi = v3(a[3], a[4], a[5]);
                    if (c->meshes[m].n_vertices) {
                        for (int k = 0; k < 6; ++k) { const float v = std::fabs(a[k]); if (v > c->world_max_abs) c->world_max_abs = v; }
                        for (int k = 0; k < 3; ++k) { if (a[k] < c->world_lo[k]) c->world_lo[k] = a[k]; if (a[3 + k] > c->world_hi[k]) c->world_hi[k] = a[3 + k]; }
                    }
                    if (c->by_id()) {
                        float * f = c->leaf_fat.data() + 6 * size_t(m);
                        Box stored; stored.lo = v3(f[0], f[1], f[2]); stored.hi = v3(f[3], f[4], f[5]);
                        if (!box_contains(stored, b)) {
                            const V3 lo = b.lo - c->cfg.leaf_margin, hi = b.hi + c->cfg.leaf_margin;
                            f[0] = lo.x; f[1] = lo.y; f[2] = lo.z; f[3] = hi.x; f[4] = hi.y; f[5] = hi.z;
                            refit.push_back(m);
                            changed = true;
                        }
                    } else if (c->tree.update(&c->meshes[m].tree_index, &b, 1) != 0) {
                        changed = true;
                    }
                }
            }
            // triangle + cull records of the moved meshes (their layout does not change unless the CANONICAL tree does)
            auto rebuild_records = [&]() -> int {
                for (uint32_t model : c->moved_models) {
                    ModelHost const & mh = c->models[model];
                    for (uint32_t m = mh.first_mesh; m < mh.first_mesh + mh.n_meshes; ++m) {
                        const uint32_t cnt = c->meshes[m].n_vertices / 3, first = c->mesh_tri_first[m];
                        if (cnt) launch_tri_setup(c->d_world.ptr, c->d_tri_src.ptr + first, cnt, c->d_tris.ptr + 3 * size_t(first), c->d_cull.ptr + 4 * size_t(first), s);
                    }
                }
                PRTC_CUDA(cudaGetLastError());
                return PRTC_OK;
            };
            if (!changed) {
                int rc = rebuild_records();
                if (rc) return rc;
                c->moved_models.clear();
                c->geometry_dirty = false;
                ++c->incremental_commits;
                return PRTC_OK;
            }
            if (c->by_id() && c->leaf_slot.size() == n_meshes) {
                for (uint32_t m : refit) {
                    const float * f = c->leaf_fat.data() + 6 * size_t(m);
                    FlatNode l;
                    l.lo[0] = f[0]; l.lo[1] = f[1]; l.lo[2] = f[2]; l.hi[0] = f[3]; l.hi[1] = f[4]; l.hi[2] = f[5];
                    l.link = int32_t(c->mesh_tri_first[m]);
                    l.count = int32_t(c->meshes[m].n_vertices / 3);
                    PRTC_CUDA(cudaMemcpyAsync(c->d_leaves.ptr + 2 * size_t(c->leaf_slot[m]), &l, sizeof(FlatNode), cudaMemcpyHostToDevice, s));
                }
                launch_tree_build(c->d_leaves.ptr, uint32_t(c->flat.leaf_order.size()), c->d_nodes.ptr, s);
                if (!c->top.empty()) launch_top_refresh(c->d_nodes.ptr, c->d_top.ptr, c->d_top_src.ptr, uint32_t(c->top.size()), s);
                int rc = rebuild_records();
                if (rc) return rc;
                PRTC_CUDA(cudaStreamSynchronize(s));               // the staged leaf records are stack-owned
                ++c->tree_version;                                 // leaf boxes changed: every cached leaf list is stale
                c->moved_models.clear();
                c->geometry_dirty = false;
                ++c->incremental_commits;
                return PRTC_OK;
            }
            // CANONICAL: the host tree changed shape -> re-flatten below; the triangle records keep their (registration-order)
            // places, only the moved meshes' records are rebuilt -- unless the leaf-ordered layout is in use (A/B switch)
            static const bool leaf_layout = [] { const char * e = std::getenv("PRTC_LEAF_ORDER"); return e && std::atoi(e) != 0; }();
            if (leaf_layout) c->tri_dirty = true;
            else { int rc = rebuild_records(); if (rc) return rc; }
            c->moved_models.clear();
            c->geometry_dirty = false;
            c->tree_dirty = true;
            ++c->incremental_commits;
        }
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

        c->h_aabbs = aabbs;
        c->registry_dirty = false;
        auto box_of = [&](uint32_t m) {
            Box b;
            b.lo = v3(aabbs[6 * size_t(m)], aabbs[6 * size_t(m) + 1], aabbs[6 * size_t(m) + 2]);
            b.hi = v3(aabbs[6 * size_t(m) + 3], aabbs[6 * size_t(m) + 4], aabbs[6 * size_t(m) + 5]);
            return b;
        };
        // moved models first (they were registered earlier): DynamicAABBTree::update per model, physics_system.cpp:199-200
        // (LITERAL order has done this in prtc_update_models already, where the reference does it)
        // BY_ID keeps no host tree, only every leaf's stored fat box, with the reference's rules: a new leaf stores its box
        // grown by the margin (insertLeaf, aabb_tree.cpp:140-141), a moved one keeps the stored box while it still
        // contains the new one (update, aabb_tree.cpp:102-111) -- the candidate SET of a query depends on nothing else
        auto fatten = [&](uint32_t m) {
            const Box b = box_of(m);
            const V3 lo = b.lo - c->cfg.leaf_margin, hi = b.hi + c->cfg.leaf_margin;
            float * f = c->leaf_fat.data() + 6 * size_t(m);
            f[0] = lo.x; f[1] = lo.y; f[2] = lo.z; f[3] = hi.x; f[4] = hi.y; f[5] = hi.z;
        };
        if (c->by_id()) c->leaf_fat.resize(size_t(n_meshes) * 6);
        for (uint32_t model : c->moved_models) {
            if (c->literal()) break;
            ModelHost const & mh = c->models[model];
            for (uint32_t m = mh.first_mesh; m < mh.first_mesh + mh.n_meshes && m < c->meshes_in_tree; ++m) {
                Box b = box_of(m);
                if (c->by_id()) {
                    const float * f = c->leaf_fat.data() + 6 * size_t(m);
                    Box stored; stored.lo = v3(f[0], f[1], f[2]); stored.hi = v3(f[3], f[4], f[5]);
                    if (!box_contains(stored, b)) fatten(m);
                } else {
                    c->tree.update(&c->meshes[m].tree_index, &b, 1);
                }
            }
        }
        c->moved_models.clear();
        // extent of the world: meshes without vertices (an empty mesh, a removed model) have the fold's start values
        // +-FLT_MAX as their box (physics_system.cpp:80-82) and say nothing about where the geometry is
        c->world_max_abs = 0.0f;
        for (int k = 0; k < 3; ++k) { c->world_lo[k] = 3.0e38f; c->world_hi[k] = -3.0e38f; }
        for (uint32_t m = 0; m < n_meshes; ++m) {
            if (c->meshes[m].n_vertices == 0 || c->models[c->meshes[m].model].n_meshes == 0) continue;
            for (int k = 0; k < 6; ++k) { const float a = std::fabs(aabbs[6 * size_t(m) + k]); if (a > c->world_max_abs) c->world_max_abs = a; }   // NaN never wins
            for (int k = 0; k < 3; ++k) {
                if (aabbs[6 * size_t(m) + k] < c->world_lo[k]) c->world_lo[k] = aabbs[6 * size_t(m) + k];
                if (aabbs[6 * size_t(m) + 3 + k] > c->world_hi[k]) c->world_hi[k] = aabbs[6 * size_t(m) + 3 + k];
            }
        }
        // new meshes: DynamicAABBTree::insert in registration order, physics_system.cpp:102
        for (uint32_t m = c->meshes_in_tree; m < n_meshes; ++m) {
            if (c->by_id()) fatten(m);
            else if (c->models[c->meshes[m].model].n_meshes != 0)      // not removed before its first commit
                c->meshes[m].tree_index = c->tree.insert_leaf(m, LEAF_MESH, box_of(m));
        }
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
        // (round 2: CANONICAL order keeps the triangle records in REGISTRATION order as well -- a moved model that re-inserts a
        //  leaf then re-flattens the nodes only; measured frame time at C4 is the same, the level is registered block by block.
        //  PRTC_LEAF_ORDER=1 restores the leaf-ordered layout for A/B.)
        static const bool leaf_layout = [] { const char * e = std::getenv("PRTC_LEAF_ORDER"); return e && std::atoi(e) != 0; }();
        const bool by_leaf_order = leaf_layout && !c->literal() && !c->by_id();
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
        ++c->full_commits;
        if (by_leaf_order) c->mesh_tri_first.assign(n_meshes, 0u);
        if (c->by_id()) {
            int rc = build_tree_on_device(c, s);
            if (rc) return rc;
        } else {
            auto leaf_range = [&](uint32_t mesh_id, uint32_t & first, uint32_t & count) {
                MeshHost const & mh = c->meshes[mesh_id];
                count = mh.n_vertices / 3;
                if (by_leaf_order) {
                    first = uint32_t(c->tri_src.size());
                    c->mesh_tri_first[mesh_id] = first;
                    for (uint32_t k = 0; k < count; ++k) c->tri_src.push_back(mh.first_vertex + 3 * k);
                } else {
                    first = c->mesh_tri_first[mesh_id];
                }
            };
            // a big mesh-only tree is laid out by several host threads (a moved model that re-inserts a leaf re-flattens the level)
            if (!c->literal() && !by_leaf_order) c->tree.flatten_parallel(c->flat, leaf_range, std::min(16u, std::max(1u, std::thread::hardware_concurrency())));
            else c->tree.flatten(c->flat, leaf_range);
        }
        const size_t nn = c->flat.nodes.size();
        const size_t nt = c->tri_src.size();
        PRTC_CUDA(c->d_nodes.reserve(nn * 2));
        PRTC_CUDA(c->d_node_leaf_id.reserve(nn));
        static_assert(sizeof(FlatNode) == 32, "FlatNode must be two float4");
        if (nn) {
            // pageable sources: the runtime stages them before it returns, so the vectors may be rewritten afterwards
            if (!c->by_id()) {                   // (BY_ID: the nodes were built where they are)
                // a big node array is page-locked where it lies (once, as long as the vector keeps its storage): the upload
                // then runs at the link's speed instead of through the runtime's staging buffer
                void * base = c->flat.nodes.data();
                const size_t bytes = c->flat.nodes.capacity() * sizeof(FlatNode);
                if (bytes >= (size_t(8) << 20) && (base != c->flat_pinned || bytes != c->flat_pinned_bytes)) {
                    if (c->flat_pinned) cudaHostUnregister(c->flat_pinned);
                    c->flat_pinned = nullptr; c->flat_pinned_bytes = 0;
                    if (cudaHostRegister(base, bytes, cudaHostRegisterDefault) == cudaSuccess) { c->flat_pinned = base; c->flat_pinned_bytes = bytes; }
                    (void)cudaGetLastError();
                }
                PRTC_CUDA(cudaMemcpyAsync(c->d_nodes.ptr, c->flat.nodes.data(), nn * sizeof(FlatNode), cudaMemcpyHostToDevice, s));
                if (c->flat_pinned) PRTC_CUDA(cudaStreamSynchronize(s));         // (a pinned source must not be rewritten under the copy)
            }
            c->leaf_ids_stale = true;            // only traces and prtc_query_box read them: uploaded on demand
        }
        // top levels as a compact table for k_xc_refresh (the LITERAL tree changes shape every frame and never gets there)
        c->top.clear();
        if (!c->literal()) build_top_table(c->flat.nodes, TOP_CAP, c->top, &c->top_src);
        if (!c->top.empty()) {
            PRTC_CUDA(c->d_top_src.reserve(c->top_src.size()));
            PRTC_CUDA(cudaMemcpyAsync(c->d_top_src.ptr, c->top_src.data(), c->top_src.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
            PRTC_CUDA(c->d_top.reserve(c->top.size() * 2));
            PRTC_CUDA(cudaMemcpyAsync(c->d_top.ptr, c->top.data(), c->top.size() * sizeof(FlatNode), cudaMemcpyHostToDevice, s));
        }
        if (rebuild_tris) {
            PRTC_CUDA(c->d_tri_src.reserve(nt));
            PRTC_CUDA(c->d_tris.reserve(nt * 3));
            PRTC_CUDA(c->d_cull.reserve(nt * 4));
            if (nt) {
                PRTC_CUDA(cudaMemcpyAsync(c->d_tri_src.ptr, c->tri_src.data(), nt * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
                launch_tri_setup(c->d_world.ptr, c->d_tri_src.ptr, uint32_t(nt), c->d_tris.ptr, c->d_cull.ptr, s);
                PRTC_CUDA(cudaGetLastError());
            }
            PRTC_CUDA(cudaStreamSynchronize(s));
        }
        c->tree_dirty = false;
        c->tri_dirty = false;
    }
    return PRTC_OK;
}

// PRTC_ORDER_BY_ID: leaf records (stored fat box, triangle range) of the live meshes in ascending id go up, the kernels of
// k_tree_build.cu put the tree together in d_nodes, and the host takes a copy for what it derives from the node array
// (top-of-tree table, collider id per node, introspection)
int build_tree_on_device(prtc_ctx * c, cudaStream_t s) {
    const uint32_t n_meshes = uint32_t(c->meshes.size());
    std::vector<FlatNode> leaves;
    c->flat.leaf_order.clear(); c->flat.capsule_order.clear();
    c->leaf_slot.assign(n_meshes, 0xFFFFFFFFu);
    for (uint32_t m = 0; m < n_meshes; ++m) {
        if (c->models[c->meshes[m].model].n_meshes == 0) continue;     // removed
        c->leaf_slot[m] = uint32_t(leaves.size());
        const float * f = c->leaf_fat.data() + 6 * size_t(m);
        FlatNode l;
        l.lo[0] = f[0]; l.lo[1] = f[1]; l.lo[2] = f[2]; l.hi[0] = f[3]; l.hi[1] = f[4]; l.hi[2] = f[5];
        l.link = int32_t(c->mesh_tri_first[m]);
        l.count = int32_t(c->meshes[m].n_vertices / 3);
        leaves.push_back(l);
        c->flat.leaf_order.push_back(m);
    }
    const uint32_t n_leaves = uint32_t(leaves.size());
    const size_t nn = n_leaves ? 2 * size_t(n_leaves) - 1 : 0;
    c->flat.nodes.resize(nn);
    c->flat.node_leaf_id.assign(nn, 0xFFFFFFFFu);
    c->flat.node_leaf_kind.assign(nn, uint8_t(LEAF_NONE));
    c->flat.max_depth = bt_levels(n_leaves);
    if (!n_leaves) return PRTC_OK;
    PRTC_CUDA(c->d_leaves.reserve(size_t(n_leaves) * 2));
    PRTC_CUDA(c->d_nodes.reserve(nn * 2));
    PRTC_CUDA(cudaMemcpyAsync(c->d_leaves.ptr, leaves.data(), size_t(n_leaves) * sizeof(FlatNode), cudaMemcpyHostToDevice, s));
    launch_tree_build(c->d_leaves.ptr, n_leaves, c->d_nodes.ptr, s);
    PRTC_CUDA(cudaGetLastError());
    PRTC_CUDA(cudaMemcpyAsync(c->flat.nodes.data(), c->d_nodes.ptr, nn * sizeof(FlatNode), cudaMemcpyDeviceToHost, s));
    PRTC_CUDA(cudaStreamSynchronize(s));
    uint32_t j = 0;
    for (size_t i = 0; i < nn; ++i)
        if (c->flat.nodes[i].count >= 0) { c->flat.node_leaf_id[i] = c->flat.leaf_order[j++]; c->flat.node_leaf_kind[i] = uint8_t(LEAF_MESH); }
    return PRTC_OK;
}

int fill_params(prtc_ctx * c, float dt, uint32_t n, prtc_transform * d_tf, prtc_char_physics * d_ph, StepParams & p) {
    std::memset(&p, 0, sizeof(p));
    c->chain.armed = false;                                       // (prtc_step_characters_device re-arms it for its own launches)
    p.nodes = c->d_nodes.ptr;
    p.n_nodes = uint32_t(c->flat.nodes.size());
    p.tris = c->d_tris.ptr;
    p.cull = c->d_cull.ptr;
    p.push_slack = c->push_slack;
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
        p.qr_enable = (m < 1.0e6f) ? ((c->quick_reject ? 1u : 0u) | (c->tight_reject ? 2u : 0u)) : 0u;
        p.tune = (c->frame_cache ? 0u : 1u) | (c->warp_kernel ? 0u : 2u) | c->tune_hi;
        p.work_counter = c->persistent ? c->d_work.ptr : nullptr;
        p.work_pair = c->d_work.ptr + 8;
        p.work_flip = &c->work_flip;
        if (c->timeline && n) {                                    // diagnostic: room for one entry per claim, whatever the claim size
            if (c->d_timeline.reserve(3 * (size_t(n) + 1)) == cudaSuccess) {
                cudaMemsetAsync(c->d_timeline.ptr, 0, 3 * (size_t(n) + 1) * sizeof(unsigned long long), c->stream);
                p.timeline = c->d_timeline.ptr;
            }
        }
        p.persistent_blocks = uint32_t(c->sm_count) * 20u;
        p.pool_blocks = uint32_t(c->sm_count) * pool_blocks_per_sm();
        p.n_tris = uint32_t(c->tri_src.size());
        p.pool = (c->pool_kernel && p.n_tris < (1u << 27)) ? 1u : 0u;    // a job word holds 27 bits of triangle index
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
            if (c->top_staging && !c->top.empty()) { p.top = c->d_top.ptr; p.n_top = uint32_t(c->top.size()); }
        }
    }
    return PRTC_OK;
}

// characters whose collider id names no registered capsule are skipped by the kernels and counted in slot 6
int check_bad_colliders(prtc_ctx * c, cudaStream_t on) {
    cudaStream_t s = on ? on : c->stream;
    unsigned long long both[2] = {0, 0};
    PRTC_CUDA(cudaMemcpyAsync(both, c->d_counters.ptr + 6, sizeof(both), cudaMemcpyDeviceToHost, s));
    PRTC_CUDA(cudaStreamSynchronize(s));
    if (both[1]) {
        PRTC_CUDA(cudaMemsetAsync(c->d_counters.ptr + 7, 0, sizeof(unsigned long long), s));
        return fail(PRTC_ERR_CUDA, "the record exchange of the dynamic pass timed out waiting for a peer rank (prtc_dyn_ipc_*)");
    }
    const unsigned long long bad = both[0];
    if (bad) {
        PRTC_CUDA(cudaMemsetAsync(c->d_counters.ptr + 6, 0, sizeof(bad), s));
        return fail(PRTC_ERR_INVALID, "prtc_step_characters: " + std::to_string(bad) + " character(s) name a capsule collider that was never added; their results are undefined");
    }
    return PRTC_OK;
}

int check_mode(prtc_ctx * c, bool device_pointers) {
    if (c->cfg.order_mode != PRTC_ORDER_CANONICAL && c->cfg.order_mode != PRTC_ORDER_LITERAL && c->cfg.order_mode != PRTC_ORDER_BY_ID)
        return fail(PRTC_ERR_INVALID, "unknown order_mode");
    if (c->literal() && device_pointers)
        return fail(PRTC_ERR_INVALID, "PRTC_ORDER_LITERAL maintains the unified tree on the host every frame and needs the "
                                      "host-buffer entry points (prtc_step_characters / _traced)");
    if (c->cfg.dyn_mode != PRTC_DYN_OFF && c->cfg.dyn_mode != PRTC_DYN_JACOBI)
        return fail(PRTC_ERR_INVALID, "unknown dyn_mode");
    return PRTC_OK;
}

// collider id per node of the flattened tree: what traced kernels and prtc_query_box report
int upload_leaf_ids(prtc_ctx * c) {
    if (!c->leaf_ids_stale) return PRTC_OK;
    const size_t nn = c->flat.node_leaf_id.size();
    PRTC_CUDA(c->d_node_leaf_id.reserve(nn));
    if (nn) PRTC_CUDA(cudaMemcpyAsync(c->d_node_leaf_id.ptr, c->flat.node_leaf_id.data(), nn * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
    c->leaf_ids_stale = false;
    return PRTC_OK;
}

} // namespace host
} // namespace prt

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
    if (const char * e = std::getenv("PRTC_TOP_STAGING")) ctx->top_staging = std::atoi(e);
    if (const char * e = std::getenv("PRTC_POOL_KERNEL")) ctx->pool_kernel = std::atoi(e);
    if (const char * e = std::getenv("PRTC_TIGHT_REJECT")) ctx->tight_reject = std::atoi(e);
    if (const char * e = std::getenv("PRTC_CHAIN_FRAMES")) ctx->chain_frames = std::atoi(e) != 0;
    if (const char * e = std::getenv("PRTC_STAGED")) ctx->staged = std::atoi(e);
    if (const char * e = std::getenv("PRTC_XC_MARGIN")) ctx->xc_margin = float(std::atof(e));
    if (const char * e = std::getenv("PRTC_TUNE")) ctx->tune_hi = uint32_t(std::strtoul(e, nullptr, 0)) & 0xFFFFFF00u;
    if (const char * e = std::getenv("PRTC_DOORBELL")) ctx->doorbell = std::atoi(e) != 0;
    if (const char * e = std::getenv("PRTC_TIMELINE")) ctx->timeline = std::atoi(e) != 0;
    if (const char * e = std::getenv("PRTC_PUSH_SLACK")) { const float v = float(std::atof(e)); if (v >= 0.0f && v <= 1.0f) ctx->push_slack = v; }
    ctx->sm_count = prop.multiProcessorCount;
    ctx->tree.set_margin(c.leaf_margin);
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete ctx;
        return fail(PRTC_ERR_CUDA, "prtc_create: cudaStreamCreate failed");
    }
    ctx->own_stream = true;
    if (ctx->d_work.reserve(16) != cudaSuccess || ctx->d_counters.reserve(16) != cudaSuccess || cudaMemsetAsync(ctx->d_counters.ptr, 0, 16 * sizeof(unsigned long long), ctx->stream) != cudaSuccess ||
        cudaMemsetAsync(ctx->d_work.ptr, 0, 16 * sizeof(unsigned), ctx->stream) != cudaSuccess) {
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
    c->d_nodes.release(); c->d_leaves.release(); c->d_top.release(); c->d_top_src.release();
    if (c->flat_pinned) { cudaHostUnregister(c->flat_pinned); c->flat_pinned = nullptr; } c->d_tris.release(); c->d_cull.release(); c->d_node_leaf_id.release(); c->d_tri_src.release();
    c->d_capsules.release(); c->d_counters.release(); c->d_work.release(); c->d_ingest.release(); c->stage.release(); c->resident_io.release();
    c->d_xc_version.release(); c->d_xc_count.release(); c->d_xc_leaf.release(); c->d_xc_region.release(); c->d_tf.release(); c->d_ph.release();
    c->d_tq.release(); c->d_tmesh.release(); c->d_tcaps.release(); c->d_thits.release();
    c->d_scratch_f.release(); c->d_scratch_u.release();
    c->d_cap2char.release(); c->d_seg_start.release(); c->d_seg_prev.release(); c->d_last_box.release();
    for (void * q : c->dyn_ipc.opened) cudaIpcCloseMemHandle(q);
    c->dyn_ipc.opened.clear(); c->dyn_ipc.flags.release();
    c->d_dyn_records.release(); c->d_dyn_count.release(); c->d_dyn_start.release(); c->d_dyn_items.release(); c->d_dyn_reach.release(); c->d_dyn_cursor.release(); c->d_dyn_scan.release(); c->d_timeline.release(); c->d_chain_done.release(); c->d_chain_ring.release();
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    for (int k = 0; k < prtc_ctx::PIPE_STREAMS; ++k) if (c->pipe[k]) cudaStreamDestroy(c->pipe[k]);
    if (c->dyn_ipc.push_stream) cudaStreamDestroy(c->dyn_ipc.push_stream);
    if (c->dyn_ipc.prep_done) cudaEventDestroy(c->dyn_ipc.prep_done);
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
    c->registry_dirty = true;
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

int prtc_remove_model(prtc_ctx * c, uint32_t model_id) {
    if (!c || model_id >= c->models.size()) return fail(PRTC_ERR_INVALID, "prtc_remove_model: bad model id");
    ModelHost & mh = c->models[model_id];
    // DynamicAABBTree::remove over the model's mesh leaves in order (physics_system.cpp:136)
    for (uint32_t m = mh.first_mesh; m < mh.first_mesh + mh.n_meshes; ++m) {
        MeshHost & me = c->meshes[m];
        if (m < c->meshes_in_tree && me.tree_index >= 0) c->tree.remove(me.tree_index);     // (BY_ID never has a tree index)
        me.tree_index = -1;
        me.n_vertices = 0;          // the vertices stay in the registry but are never referenced again
    }
    mh.n_meshes = 0;
    c->geometry_dirty = true;
    c->registry_dirty = true;
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

int prtc_set_option(prtc_ctx * c, int option, int value) {
    if (!c) return fail(PRTC_ERR_INVALID, "prtc_set_option: null context");
    switch (option) {
        case PRTC_OPT_FRAME_CACHE: c->frame_cache = value != 0; return PRTC_OK;
        case PRTC_OPT_QUICK_REJECT: c->quick_reject = value != 0; return PRTC_OK;
        case PRTC_OPT_PERSISTENT: c->persistent = value != 0; return PRTC_OK;
        case PRTC_OPT_CROSS_FRAME_CACHE: c->cross_frame = value != 0; return PRTC_OK;
        case PRTC_OPT_WARP_KERNEL: c->warp_kernel = value != 0; return PRTC_OK;
        case PRTC_OPT_STREAMED_COPY: c->streamed_copy = value != 0; return PRTC_OK;
        case PRTC_OPT_TOP_STAGING: c->top_staging = value != 0; return PRTC_OK;
        case PRTC_OPT_POOL_KERNEL: c->pool_kernel = value != 0; return PRTC_OK;
        case PRTC_OPT_TIGHT_REJECT: c->tight_reject = value != 0; return PRTC_OK;
        case PRTC_OPT_CHAIN_FRAMES: c->chain_frames = value != 0; c->chain.armed = false; return PRTC_OK;
        default: return fail(PRTC_ERR_INVALID, "prtc_set_option: unknown option");
    }
}

int prtc_get_commit_stats(prtc_ctx * c, uint32_t * incremental, uint32_t * full) {
    if (!c) return fail(PRTC_ERR_INVALID, "prtc_get_commit_stats: null context");
    if (incremental) *incremental = c->incremental_commits;
    if (full) *full = c->full_commits;
    return PRTC_OK;
}

int prtc_debug_timeline(prtc_ctx * c, unsigned long long * out, size_t capacity_words, size_t * n_words) {
    if (!c || !n_words) return fail(PRTC_ERR_INVALID, "prtc_debug_timeline: null argument");
    *n_words = c->d_timeline.ptr ? c->d_timeline.cap : 0;
    if (!out || !c->d_timeline.ptr) return PRTC_OK;
    PRTC_CUDA(cudaSetDevice(c->cfg.device));
    PRTC_CUDA(cudaStreamSynchronize(c->stream));
    PRTC_CUDA(cudaMemcpy(out, c->d_timeline.ptr, std::min(capacity_words, c->d_timeline.cap) * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    return PRTC_OK;
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
    // (PRTC_HOST_FLATTEN_THREADS: lay the tree out with that many host threads, as a commit of a big level does -- for the tests)
    const char * ft = std::getenv("PRTC_HOST_FLATTEN_THREADS");
    if (ft && std::atoi(ft) > 1) tree.flatten_parallel(flat, [&](uint32_t id, uint32_t & first, uint32_t & count) { first = id; count = 1; }, unsigned(std::atoi(ft)));
    else tree.flatten(flat, [&](uint32_t id, uint32_t & first, uint32_t & count) { first = id; count = 1; });
    *n_nodes = uint32_t(flat.nodes.size());
    if (leaf_order) std::copy(flat.leaf_order.begin(), flat.leaf_order.end(), leaf_order);
    if (nodes8) {
        if (nodes_capacity < flat.nodes.size()) return fail(PRTC_ERR_CAPACITY, "prtc_host_build_tree: node buffer too small");
        std::memcpy(nodes8, flat.nodes.data(), flat.nodes.size() * sizeof(FlatNode));
    }
    return PRTC_OK;
}

int prtc_host_balanced_tree(const float * aabbs6, uint32_t n, float * nodes8, uint32_t nodes_capacity, uint32_t * n_nodes) {
    if ((n && !aabbs6) || !n_nodes) return fail(PRTC_ERR_INVALID, "prtc_host_balanced_tree: null argument");
    const uint32_t nn = n ? 2 * n - 1 : 0;
    *n_nodes = nn;
    if (!nodes8) return PRTC_OK;
    if (nodes_capacity < nn) return fail(PRTC_ERR_CAPACITY, "prtc_host_balanced_tree: node buffer too small");
    FlatNode * nodes = reinterpret_cast<FlatNode *>(nodes8);
    const uint32_t k = bt_levels(n);
    // the same passes as k_tree_build.cu, as plain loops
    for (uint32_t j = 0; j < n; ++j) {
        FlatNode & l = nodes[bt_index(0, j, k, n)];
        for (int a = 0; a < 3; ++a) { l.lo[a] = aabbs6[6 * size_t(j) + a]; l.hi[a] = aabbs6[6 * size_t(j) + 3 + a]; }
        l.link = int32_t(j); l.count = 1;
    }
    auto lo_of = [](float a, float b) { return (a != a || b != b) ? NAN : std::fmin(a, b); };
    auto hi_of = [](float a, float b) { return (a != a || b != b) ? NAN : std::fmax(a, b); };
    for (uint32_t level = 1; level <= k; ++level) {
        const uint32_t n_blocks = uint32_t((uint64_t(n) + (1ull << level) - 1) >> level);
        for (uint32_t j = 0; j < n_blocks; ++j) {
            if (!bt_is_node(level, j, n)) continue;
            const uint32_t idx = bt_index(level, j, k, n), c1 = idx + 1, c2 = idx + (1u << level);
            FlatNode & x = nodes[idx];
            for (int a = 0; a < 3; ++a) { x.lo[a] = lo_of(nodes[c1].lo[a], nodes[c2].lo[a]); x.hi[a] = hi_of(nodes[c1].hi[a], nodes[c2].hi[a]); }
            x.link = int32_t(idx + 2u * bt_leaves_of(level, j, n) - 1u);
            x.count = -int32_t(c2);
        }
    }
    return PRTC_OK;
}

int prtc_host_top_table(const float * nodes8, uint32_t n_nodes, uint32_t cap, float * top8, uint32_t top_capacity, uint32_t * n_top) {
    if ((n_nodes && !nodes8) || !n_top) return fail(PRTC_ERR_INVALID, "prtc_host_top_table: null argument");
    std::vector<FlatNode> nodes(n_nodes), top;
    if (n_nodes) std::memcpy(nodes.data(), nodes8, size_t(n_nodes) * sizeof(FlatNode));
    build_top_table(nodes, cap, top);
    *n_top = uint32_t(top.size());
    if (top8) {
        if (top_capacity < top.size()) return fail(PRTC_ERR_CAPACITY, "prtc_host_top_table: table buffer too small");
        if (!top.empty()) std::memcpy(top8, top.data(), top.size() * sizeof(FlatNode));
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

int prtc_selftest_math(prtc_ctx * c, uint32_t n_threads, uint64_t seed, int mode, uint64_t * mismatches) {
    if (!c || !mismatches || mode < 0 || mode > 3) return fail(PRTC_ERR_INVALID, "prtc_selftest_math: bad argument");
    PRTC_CUDA(cudaSetDevice(c->cfg.device));
    PRTC_CUDA(c->d_scratch_u.reserve(4));
    unsigned long long * bad = reinterpret_cast<unsigned long long *>(c->d_scratch_u.ptr);
    PRTC_CUDA(cudaMemsetAsync(bad, 0, sizeof(unsigned long long), c->stream));
    launch_selftest_math(seed, n_threads, mode, bad, c->stream);
    PRTC_CUDA(cudaGetLastError());
    unsigned long long h = 0;
    PRTC_CUDA(cudaMemcpyAsync(&h, bad, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    PRTC_CUDA(cudaStreamSynchronize(c->stream));
    *mismatches = h;
    return PRTC_OK;
}

} // extern "C"
