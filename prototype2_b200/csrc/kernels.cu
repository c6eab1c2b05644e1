// kernels.cu -- the auxiliary sm_100a kernels of the character-collision path and the step dispatcher.
//
//   launch_step        picks the step kernel for a batch: k_step_warp (one warp per character, small batches;
//                      k_step_warp.cu), k_step_stream (persistent, static pass at scale; k_step_stream.cu, preceded by
//                      k_xc_refresh) or k_step (lock-step, every mode, traced or not; k_step.cu).  Together they are
//                      K1+K2+K3+K4+K6 of SURVEY.md 2.2 fused: PhysicsSystem::updateCharacters for one frame
//                      (reference physics_system.cpp:245-318,330-438).
//   k_transform_refit  K8: world-space vertex cache and per-mesh AABB (physics_system.cpp:68-93,161-202)
//   k_tri_setup        leaf-ordered triangle records with precomputed unit normals
//   k_dyn_prep/_grid   K1/K5 of the JACOBI dynamic pass: per-character records and the uniform grid over stored boxes
//   k_raycast          K7: PhysicsSystem::raycast (physics_system.cpp:205-242)
//   k_query_box        one box against the flattened tree (introspection)
//
// Compile with -fmad=false: results must be bit-identical to the CPU reference (see prt_math.cuh).
#include <cstring>
#include <cstdlib>

#include "step_common.cuh"


namespace prt {

namespace {

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
                            uint32_t n_tris, float4 * __restrict__ tris, float4 * __restrict__ cull) {
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
    if (cull) {                                                   // cull record of the pooled kernel (collide.cuh make_cull_record)
        const CullRec r = make_cull_record(a, b, c, n, !ok);
        cull[4 * size_t(t)] = make_float4(r.pl[0], r.pl[1], r.pl[2], r.pl[3]);
        cull[4 * size_t(t) + 1] = make_float4(r.e0[0], r.e0[1], r.e0[2], r.e0[3]);
        cull[4 * size_t(t) + 2] = make_float4(r.e1[0], r.e1[1], r.e1[2], r.e1[3]);
        cull[4 * size_t(t) + 3] = make_float4(r.e2[0], r.e2[1], r.e2[2], r.e2[3]);
    }
}

// the staged top-of-tree table after the node boxes changed but the topology did not (PRTC_ORDER_BY_ID, a moved model):
// every entry takes its box from the node it mirrors, the two integer words (table links) stay
__global__ void k_top_refresh(const float4 * __restrict__ nodes, float4 * __restrict__ top, const uint32_t * __restrict__ src, uint32_t n_top) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_top) return;
    const float4 lo = nodes[2 * size_t(src[k])], hi = nodes[2 * size_t(src[k]) + 1];
    float4 tlo = top[2 * size_t(k)], thi = top[2 * size_t(k) + 1];
    tlo.x = lo.x; tlo.y = lo.y; tlo.z = lo.z; thi.x = hi.x; thi.y = hi.y; thi.z = hi.z;
    top[2 * size_t(k)] = tlo; top[2 * size_t(k) + 1] = thi;
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

// Engine-sized trees (the shipped level has 345 nodes): ONE WARP PER RAY.  The result is the minimum t over every triangle of
// every mesh leaf the reference's traversal reaches, and a leaf is reached exactly when its own box passes the slab test --
// its ancestors' boxes contain it and the test is monotone in the box (correctly rounded subtractions and multiplications
// are monotone) -- so the hierarchy is not needed for the answer: the lanes test all nodes 32 at a time (a handful of
// independent round trips instead of one dependent load per tree level and per triangle), the triangles of the hit leaves
// are spread over the lanes, and the minimum is reduced with shuffles (min of floats is order independent).
__global__ void __launch_bounds__(32) k_raycast_warp(const float4 * __restrict__ nodes, uint32_t n_nodes, const float4 * __restrict__ tris, uint32_t n_rays,
                                                     const float * __restrict__ origins, const float * __restrict__ dirs, const float * __restrict__ maxd,
                                                     float * __restrict__ hit_xyz, unsigned char * __restrict__ hit_mask, uint32_t * done_flag,
                                                     unsigned * done_count, uint32_t done_value) {
    constexpr unsigned FULL = 0xffffffffu;
    const uint32_t r = blockIdx.x, lane = threadIdx.x;
    if (r < n_rays) {
        const V3 o = v3(origins[3 * r], origins[3 * r + 1], origins[3 * r + 2]);
        const V3 d = v3(dirs[3 * r], dirs[3 * r + 1], dirs[3 * r + 2]);
        const float md = maxd[r];
        const V3 end = o + d * md;
        float best = 3.402823466e+38f;                            // std::numeric_limits<float>::max()
        bool any = false;
        for (uint32_t base = 0; base < n_nodes; base += 32) {
            const uint32_t k = base + lane;
            bool hit_leaf = false;
            float4 lo = make_float4(0, 0, 0, 0), hi = lo;
            if (k < n_nodes) {
                lo = __ldg(nodes + 2 * k); hi = __ldg(nodes + 2 * k + 1);
                const int cnt = __float_as_int(hi.w);
                hit_leaf = cnt >= 0 && !(cnt & PRT_CAPSULE_LEAF) && ray_box(lo, hi, o, d, md);
            }
            unsigned m = __ballot_sync(FULL, hit_leaf);
            while (m) {
                const int src = __ffs(m) - 1;
                m &= m - 1;
                const uint32_t first = uint32_t(__float_as_int(__shfl_sync(FULL, lo.w, src)));
                const uint32_t count = uint32_t(__float_as_int(__shfl_sync(FULL, hi.w, src)));
                for (uint32_t t = first + lane; t < first + count; t += 32) {
                    const float4 t0 = __ldg(tris + 3 * size_t(t)), t1 = __ldg(tris + 3 * size_t(t) + 1), t2 = __ldg(tris + 3 * size_t(t) + 2);
                    float tt;
                    if (segment_triangle(o, end, v3(t0.x, t0.y, t0.z), v3(t1.x, t1.y, t1.z), v3(t2.x, t2.y, t2.z), tt)) {
                        any = true;
                        best = (tt < best) ? tt : best;           // std::min(intersectionTime, t)
                    }
                }
            }
        }
#pragma unroll
        for (int sft = 16; sft > 0; sft >>= 1) { const float other = __shfl_xor_sync(FULL, best, sft); best = (other < best) ? other : best; }
        any = __any_sync(FULL, any);
        if (lane == 0) {
            hit_mask[r] = any ? 1 : 0;
            if (any) {
                const V3 h = o + (d * best) * md;                 // origin + direction * intersectionTime * maxDistance
                hit_xyz[3 * r] = h.x; hit_xyz[3 * r + 1] = h.y; hit_xyz[3 * r + 2] = h.z;
            }
        }
    }
    if (done_flag) {                                               // completion doorbell (StepParams::done_flag)
        __threadfence_system();
        __syncwarp(FULL);
        if (lane == 0 && atomicAdd(done_count, 1u) + 1u == gridDim.x) {
            *done_count = 0u;
            __threadfence_system();
            *reinterpret_cast<volatile uint32_t *>(done_flag) = done_value;
        }
    }
}

void launch_raycast(const float4 * nodes, uint32_t n_nodes, const float4 * tris, uint32_t n_rays, const float * origins,
                    const float * dirs, const float * maxd, float * hit_xyz, unsigned char * hit_mask, cudaStream_t stream,
                    uint32_t * done_flag, unsigned * done_count, uint32_t done_value) {
    if (n_rays == 0) return;
    if (n_nodes <= 4096u && n_rays <= 65535u) {
        k_raycast_warp<<<n_rays, 32, 0, stream>>>(nodes, n_nodes, tris, n_rays, origins, dirs, maxd, hit_xyz, hit_mask, done_flag, done_count, done_value);
        return;
    }
    const uint32_t block = 64;
    k_raycast<<<(n_rays + block - 1) / block, block, 0, stream>>>(nodes, n_nodes, tris, n_rays, origins, dirs, maxd, hit_xyz, hit_mask);
}
bool raycast_rings_doorbell(uint32_t n_nodes, uint32_t n_rays) { return n_nodes <= 4096u && n_rays <= 65535u && n_rays > 0; }

// resident state, step kernels that do not read / write the I/O arrays themselves: movement in before, results out after
__global__ void k_resident_in(uint32_t n, const float * __restrict__ io_move, prtc_char_physics * __restrict__ ph) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    ph[i].movement[0] = io_move[3 * size_t(i)]; ph[i].movement[1] = io_move[3 * size_t(i) + 1]; ph[i].movement[2] = io_move[3 * size_t(i) + 2];
}
__global__ void k_resident_out(uint32_t n, const prtc_transform * __restrict__ tf, const prtc_char_physics * __restrict__ ph,
                               float4 * __restrict__ io_result, uint32_t * done_flag, unsigned * done_count, uint32_t done_value) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) io_result[i] = make_float4(tf[i].position[0], tf[i].position[1], tf[i].position[2], ph[i].grounded ? 1.0f : 0.0f);
    if (done_flag) {                                               // doorbell, see StepParams::done_flag
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0 && atomicAdd(done_count, 1u) + 1u == gridDim.x) {
            *done_count = 0u;
            __threadfence_system();
            *reinterpret_cast<volatile uint32_t *>(done_flag) = done_value;
        }
    }
}

static int launch_step_inner(StepParams const & p_in, bool traced, int mode, cudaStream_t stream);

// only the pooled path reads slots that a caller may have refreshed ahead of time
bool launch_slot_refresh(StepParams const & p, int mode, cudaStream_t stream) {
    if (p.n == 0 || !(p.pool && p.work_counter && mode != MODE_LITERAL) || p.xc_version == nullptr || (p.tune & 1u)) return false;
    launch_k_xc_refresh(p, stream);
    return true;
}


int launch_step(StepParams const & p_in, bool traced, int mode, cudaStream_t stream) {
    if (p_in.n == 0) return 0;
    // which kernel will run is decided inside; the pooled kernel reads / writes the resident I/O arrays itself, the others
    // get the two small kernels around them
    StepParams p = p_in;
    const bool pooled_first = p.pool && p.work_counter && !traced && mode != MODE_LITERAL;
    const uint32_t resident_warps = pooled_first ? p.pool_blocks : (p.persistent_blocks ? p.persistent_blocks : 2368u);
    const bool wpc = !traced && (p.n + resident_warps - 1) / resident_warps <= 1u && mode != MODE_JACOBI && !(p.tune & 2u);
    const bool native_io = pooled_first && !wpc;
    int launches = 0;
    if (p.io_move && !native_io) { k_resident_in<<<(p.n + 255) / 256, 256, 0, stream>>>(p.n, p.io_move, p.physics); ++launches; }
    float4 * io_result = p.io_result;
    uint32_t * const bell = p.done_flag;
    if (!native_io) { p.io_move = nullptr; p.io_result = nullptr; }
    if (io_result && !native_io) p.done_flag = nullptr;     // the kernel behind the step writes the results: it rings the doorbell
    launches += launch_step_inner(p, traced, mode, stream);
    if (io_result && !native_io) { k_resident_out<<<(p.n + 255) / 256, 256, 0, stream>>>(p.n, p.transforms, p.physics, io_result, bell, p.done_count, p.done_value); ++launches; }
    return launches;
}

static int launch_step_inner(StepParams const & p_in, bool traced, int mode, cudaStream_t stream) {
    int launches = 1;
    StepParams p = p_in;
    const uint32_t block = STEP_BLOCK;
    const bool pooled = p.pool && p.work_counter && !traced && mode != MODE_LITERAL;
    // fewer characters than resident lanes: thin the warps out (1..32 characters per warp)
    const uint32_t resident_warps = pooled ? p.pool_blocks : (p.persistent_blocks ? p.persistent_blocks : 2368u);
    uint32_t lpw = (p.n + resident_warps - 1) / resident_warps;
    const bool warp_per_character = !traced && lpw <= 1u && mode != MODE_JACOBI && !(p.tune & 2u);
    if (pooled && !warp_per_character) {
        // the pooled kernel: 32 characters per warp when the batch fills the machine, else 4 / 8 / 16 so that every
        // resident warp slot gets work (the pooled cull and exact stages use all 32 lanes whatever the number of owners)
        uint32_t opw = 4u;
        while (opw < 32u && (p.n + opw - 1) / opw > p.pool_blocks) opw <<= 1;
        if (opw == 32u) {
            // a batch of a few rounds of claims: size the claims so that the rounds come out whole (125 000 characters on
            // 2368 resident warps: 27 per claim = 2 rounds, instead of 32 per claim = 1.65 rounds with a half-empty tail)
            const uint32_t rounds = (p.n + 32u * p.pool_blocks - 1) / (32u * p.pool_blocks);
            if (rounds <= 6u) opw = (p.n + rounds * p.pool_blocks - 1) / (rounds * p.pool_blocks);
            if (opw < 16u) opw = 16u;
            if (opw > 32u) opw = 32u;
        }
        static const uint32_t force_opw = [] { const char * e = std::getenv("PRTC_OPW"); return e ? uint32_t(std::atoi(e)) : 0u; }();
        if (force_opw >= 1u && force_opw <= 32u) opw = force_opw;   // experiments
        p.lanes_per_warp = opw;
        const uint32_t want = (p.n + opw - 1) / opw;
        // chained frames: the static pass of a batch of at least a quarter of the machine's warp slots (below that, frames pile up as
        // spinning blocks faster than they retire), nothing of the resident I/O in it
        bool chained = false;
        // ... and of few rounds of claims: the chained kernel costs 6 - 7 % per claim when frames overlap (slot refresh, claim words and
        // the acquire inside it), which pays when the tail it fills is a larger share of the frame than that -- measured: 2 rounds
        // (a 125 k shard) -10 %, 3.3 rounds -2.5 %, 6.6 rounds (500 k) -1.7 %, 13 rounds (the full 1 M batch) +1.5 %
        static const uint32_t chain_max_rounds = [] { const char * e = std::getenv("PRTC_CHAIN_MAX_ROUNDS"); return e ? uint32_t(std::atoi(e)) : 8u; }();
        if (p.chain && p.chain_done && p.chain_ring && mode == MODE_STATIC && 4u * want >= p.pool_blocks && want <= chain_max_rounds * p.pool_blocks && p.xc_version != nullptr &&
            !(p.tune & 1u) && !p.io_move && !p.io_result && !p.done_flag && !p.host_flags) {
            ChainState & cs = *p.chain;
            const unsigned long long sig[6] = {p.n, opw, (unsigned long long)(uintptr_t)p.transforms, (unsigned long long)(uintptr_t)p.physics,
                                               (unsigned long long)(uintptr_t)stream, p.tree_version};
            static const bool no_follow = std::getenv("PRTC_CHAIN_NOFOLLOW") != nullptr;     // experiments: the chained kernel, launch after launch
            const bool follow = cs.armed && std::memcmp(sig, cs.sig, sizeof(sig)) == 0 && !no_follow;
            if (!follow) { cudaMemsetAsync(p.chain_ring, 0, CHAIN_RING * sizeof(unsigned), stream); cs.ring = 0; }
            const uint32_t k = cs.ring % CHAIN_RING;
            ++cs.ring;
            p.work_counter = p.chain_ring + k;
            p.work_next = p.chain_ring + ((k + 1u) % CHAIN_RING);
            p.chain_prev = follow ? cs.serial : 0u;
            if (++cs.serial == 0u) cs.serial = 1u;
            p.chain_serial = cs.serial;
            std::memcpy(cs.sig, sig, sizeof(sig));
            cs.armed = true;
            chained = true;
        } else {
            if (p.chain) p.chain->armed = false;
            p.chain_done = nullptr;
        }
        if (p.xc_version != nullptr && !(p.tune & 1u)) { if (!p.xc_fresh && !chained) { launch_k_xc_refresh(p, stream); ++launches; } }
        else p.xc_version = nullptr;
        launch_k_step_pool(p, want < p.pool_blocks ? want : p.pool_blocks, false, mode, stream);
        return launches;
    }
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
    if (slots) { launch_k_xc_refresh(p, stream); ++launches; }
    if (warp_per_character) {
        launch_k_step_warp(p, mode, stream);
    } else if (mode == MODE_STATIC && !traced && p.work_counter) {
        // persistent: one warp per block, as many blocks as fit (20 per SM), work handed out by a counter
        const uint32_t want = (p.n + 31) / 32;
        launch_k_step_stream(p, want < p.persistent_blocks ? want : p.persistent_blocks, false, stream);
    } else {
        launch_k_step(p, traced, mode, grid, stream);
    }
    return launches;
}

// Self-test of the written-out division / square-root fast paths (prt_math.cuh operator/(V3, float), div2, len_inv3)
// against the library's correctly rounded operations on pseudo-random operands: mode 0 = random bit patterns (every
// exponent, denormals, infinities, NaNs, zeros), mode 1 = magnitudes around 1 (the range the collision code lives in),
// mode 2 / 3 = magnitudes at the edges of the fast paths' ranges (2^+-60 for the divisions, 2^+-100 for the square roots).  Counts results whose bits differ (two NaNs count as equal).
__device__ __forceinline__ uint32_t st_hash(uint64_t & s) {
    s ^= s << 13; s ^= s >> 7; s ^= s << 17;
    return uint32_t((s * 0x2545F4914F6CDD1Dull) >> 32);
}
__device__ __forceinline__ float st_operand(uint64_t & s, int mode) {
    const uint32_t h = st_hash(s);
    if (mode == 0) return __uint_as_float(h);
    const uint32_t sign = h & 0x80000000u, mant = h & 0x007FFFFFu;
    int e;
    if (mode == 1) e = 127 + int((h >> 23) & 15u) - 8;                                   // 2^-8 .. 2^7
    else { const int edge = mode == 2 ? 60 : 100, k = int((h >> 23) & 7u) - 4; e = ((h >> 27) & 1u) ? 127 + edge + k : 127 - edge + k; }   // around 2^+-60 / 2^+-100
    return __uint_as_float(sign | (uint32_t(e) << 23) | mant);
}
__device__ __forceinline__ bool st_same(float a, float b) {
    return __float_as_uint(a) == __float_as_uint(b) || (a != a && b != b);
}
__global__ void k_selftest_math(uint64_t seed, uint32_t n, int mode, unsigned long long * bad) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t s = seed * 0x9E3779B97F4A7C15ull + (uint64_t(i) + 1) * 0xD1B54A32D192ED03ull;
    st_hash(s); st_hash(s);
    uint32_t wrong = 0;
    for (int round = 0; round < 16; ++round) {
        const V3 a = v3(st_operand(s, mode), st_operand(s, mode), st_operand(s, mode));
        const float b = st_operand(s, mode);
        const V3 q = a / b;
        wrong += !st_same(q.x, __fdiv_rn(a.x, b)) + !st_same(q.y, __fdiv_rn(a.y, b)) + !st_same(q.z, __fdiv_rn(a.z, b));
        float u, v;
        div2(a.x, a.y, b, u, v);
        wrong += !st_same(u, __fdiv_rn(a.x, b)) + !st_same(v, __fdiv_rn(a.y, b));
        // squared lengths are sums of squares: non-negative (or NaN / inf)
        const float d1 = fabsf(a.x), d2 = fabsf(a.y), d3 = fabsf(b);
        float s1, s2, s3, i1, i2, i3;
        len_inv3(d1, d2, d3, s1, s2, s3, i1, i2, i3);
        const float r1 = __fsqrt_rn(d1), r2 = __fsqrt_rn(d2), r3 = __fsqrt_rn(d3);
        wrong += !st_same(s1, r1) + !st_same(s2, r2) + !st_same(s3, r3);
        wrong += !st_same(i1, __fdiv_rn(1.0f, r1)) + !st_same(i2, __fdiv_rn(1.0f, r2)) + !st_same(i3, __fdiv_rn(1.0f, r3));
    }
    if (wrong) atomicAdd(bad, (unsigned long long)wrong);
}
void launch_selftest_math(uint64_t seed, uint32_t n, int mode, unsigned long long * bad, cudaStream_t stream) {
    k_selftest_math<<<(n + 255) / 256, 256, 0, stream>>>(seed, n, mode, bad);
}

__global__ void k_fold_counters(unsigned long long * main_counters, unsigned long long * scratch) {
    const unsigned k = threadIdx.x;
    if (k < 6) { if (scratch[k]) atomicAdd(main_counters + k, scratch[k]); scratch[k] = 0ull; }
}
void launch_fold_counters(unsigned long long * main_counters, unsigned long long * scratch, cudaStream_t stream) {
    k_fold_counters<<<1, 32, 0, stream>>>(main_counters, scratch);
}

int launch_step_ingest(StepParams const & p_in, cudaStream_t stream) {
    if (p_in.n == 0) return 0;
    StepParams p = p_in;
    p.lanes_per_warp = 32;
    if (p.tune & 1u) p.xc_version = nullptr;
    const uint32_t want = (p.n + 31) / 32;
    if (p.pool) launch_k_step_pool(p, want < p.pool_blocks ? want : p.pool_blocks, true, MODE_STATIC, stream);
    else launch_k_step_stream(p, want < p.persistent_blocks ? want : p.persistent_blocks, true, stream);
    if (!p.xc_version) return 1;
    // the slots cannot be refreshed ahead of the kernel (the input is not there yet): refresh them behind it from the
    // state this frame leaves, which is the next frame's input unless the caller edits it -- and then the per-substep
    // "query box inside region" check of k_step_stream still decides, so a wrong guess costs a tree walk, never a result
    StepParams r = p;
    r.gate = nullptr;
    launch_k_xc_refresh(r, stream);
    return 2;
}

void launch_transform_refit(const float * raw_xyz, float * world_xyz, const MeshDev * meshes, uint32_t first_mesh,
                            uint32_t n_meshes, const ModelXform * xforms, float * mesh_aabbs, cudaStream_t stream) {
    if (n_meshes == 0) return;
    const uint32_t block = 128;
    k_transform_refit<<<(n_meshes + block - 1) / block, block, 0, stream>>>(raw_xyz, world_xyz, meshes, first_mesh,
                                                                           n_meshes, xforms, mesh_aabbs);
}

void launch_tri_setup(const float * world_xyz, const uint32_t * tri_src_vertex, uint32_t n_tris, float4 * tris, float4 * cull,
                      cudaStream_t stream) {
    if (n_tris == 0) return;
    const uint32_t block = 256;
    k_tri_setup<<<(n_tris + block - 1) / block, block, 0, stream>>>(world_xyz, tri_src_vertex, n_tris, tris, cull);
}

void launch_top_refresh(const float4 * nodes, float4 * top, const uint32_t * src, uint32_t n_top, cudaStream_t stream) {
    if (n_top) k_top_refresh<<<(n_top + 255) / 256, 256, 0, stream>>>(nodes, top, src, n_top);
}

void launch_query_box(const float4 * nodes, uint32_t n_nodes, const uint32_t * node_leaf_id, const float * aabb6,
                      uint32_t * out_ids, uint32_t capacity, uint32_t * out_count, cudaStream_t stream) {
    k_query_box<<<1, 32, 0, stream>>>(nodes, n_nodes, node_leaf_id, aabb6, out_ids, capacity, out_count);
}

} // namespace prt
