// k_sweep.cu -- TIER B: swept ellipsoid vs triangles with collide-and-slide, one warp per ellipsoid.
// Compile with -fmad=false: results must be bit-identical to the CPU checker (see prt_math.cuh).
//
// This is the narrowphase as BASELINE.json's north_star words it (ellipsoid-scaled space, earliest time of impact, slide
// planes, iterated resolve).  The reference snapshot no longer contains it: PhysicsSystem::collisionResponse is declared
// and never defined (physics_system.h:131-136); what survives are the helpers of the older implementation, "from
// Generic Collision Detection for Games Using Ellipsoids by Paul Nettle" (src/util/physics_util.cpp:51-178).  The
// helpers are restated below instruction for instruction; the loop around them is Nettle's algorithm as their
// signatures imply it.  PARITY UNPINNED -- the only checker is the same loop on the CPU around the reference's own helper
// functions (test infrastructure; see DESIGN.md "Tier B").
#include "step_common.cuh"

namespace prt {

namespace {

// ---- physics_util.cpp helpers ---------------------------------------------------------------------------------------
// :96-108  intersectRayPlane
__device__ __forceinline__ bool ray_plane(V3 planeOrigin, V3 planeNormal, V3 rayOrigin, V3 rayVector, float & t) {
    const float d = -dot(planeNormal, planeOrigin);
    const float numer = dot(planeNormal, rayOrigin) + d;
    const float denom = dot(planeNormal, rayVector);
    if (denom == 0.0f) return false;
    t = -fdiv(numer, denom);
    return true;
}
// :113-133  intersectRaySphere
__device__ __forceinline__ bool ray_sphere(V3 rayOrigin, V3 rayDirection, V3 sphereOrigin, float sphereRadius, float & t) {
    const V3 m = rayOrigin - sphereOrigin;
    const float b = dot(m, rayDirection);
    const float c = dot(m, m) - (sphereRadius * sphereRadius);
    if (c > 0.0f && b > 0.0f) return false;
    const float discr = b * b - c;
    if (discr < 0.0f) return false;
    t = -b - fsqrt(discr);
    if (t < 0.0f) t = 0.0f;
    return true;
}
// :51-94  checkPointInTriangle ("Craig's solution")
__device__ __forceinline__ bool point_in_triangle(V3 point, V3 pa, V3 pb, V3 pc) {
    const V3 u = pb - pa;
    const V3 v = pc - pa;
    const V3 n = cross(u, v);
    const V3 w = point - pa;
    const float gamma = fdiv(dot(cross(u, w), n), dot(n, n));
    const float beta = fdiv(dot(cross(w, v), n), dot(n, n));
    const float alpha = (1.0f - gamma) - beta;
    return (0.0f <= alpha) && (alpha <= 1.0f) && (0.0f <= beta) && (beta <= 1.0f) && (0.0f <= gamma) && (gamma <= 1.0f);
}
// :98-120 (first half of the file)  closestPointOnTrianglePerimeter; note the commented-out `minDist = distCA2`
__device__ __forceinline__ V3 closest_on_perimeter(V3 a, V3 b, V3 c, V3 p) {
    const V3 rab = closest_on_segment(a, b, p);
    const V3 rbc = closest_on_segment(b, c, p);
    const V3 rca = closest_on_segment(c, a, p);
    const V3 dab = p - rab, dbc = p - rbc, dca = p - rca;       // glm::distance2(r, p) = length2(p - r)
    const float distAB2 = dot(dab, dab), distBC2 = dot(dbc, dbc), distCA2 = dot(dca, dca);
    float minDist = distAB2;
    V3 minPoint = rab;
    if (distBC2 < minDist) { minDist = distBC2; minPoint = rbc; }
    if (distCA2 < minDist) minPoint = rca;
    return minPoint;
}

constexpr int SW_CAP = 128;          // frontier / leaf-list capacity per warp (overflow => ordered passes)

// One warp per ellipsoid.  Per collide-and-slide iteration: frontier walk of the flattened tree with the box of the whole
// motion -> the candidate triangles form one list -> every lane sweeps the unit sphere against every 32nd of them in
// ellipsoid space -> the earliest time of impact is reduced over the warp with shuffles (ties: lowest triangle id) ->
// the winner's contact point and normal are broadcast -> move, slide plane, project the rest of the motion.
__global__ void __launch_bounds__(32) k_sweep(SweepParams p) {
    constexpr unsigned FULL = 0xffffffffu;
    __shared__ uint32_t s_front[2][SW_CAP];
    __shared__ uint32_t s_leaf[SW_CAP];
    __shared__ uint32_t s_start[SW_CAP + 1];
    __shared__ uint32_t s_tbeg[SW_CAP];
    const uint32_t lane = threadIdx.x;
    const unsigned lt = (1u << lane) - 1u;
    const uint32_t q = blockIdx.x;
    if (q >= p.n) return;
    const V3 r = v3(p.radii[3 * q], p.radii[3 * q + 1], p.radii[3 * q + 2]);
    V3 P = v3(p.centers[3 * q], p.centers[3 * q + 1], p.centers[3 * q + 2]) / r;       // ellipsoid space
    V3 V = v3(p.velocities[3 * q], p.velocities[3 * q + 1], p.velocities[3 * q + 2]) / r;
    uint32_t c_tri = 0, c_node = 0;
    for (uint32_t it = lane; it < p.iterations; it += 32) {
        const size_t h = size_t(q) * p.iterations + it;
        p.hit_tri[h] = -1; p.hit_toi[h] = 0.0f;
        p.hit_normal[3 * h] = 0.0f; p.hit_normal[3 * h + 1] = 0.0f; p.hit_normal[3 * h + 2] = 0.0f;
    }
    __syncwarp(FULL);

    for (uint32_t it = 0; it < p.iterations; ++it) {
        const float len = length(V);
        if (!(len > 0.0f)) break;
        const V3 dir = normalize(V);
        // broadphase box in world space: the ellipsoid at both ends of the motion, plus slack for the rounding of the
        // divisions into ellipsoid space (a candidate list that is too long costs time, one that is too short a hit)
        const V3 w0 = P * r, w1 = (P + V) * r;
        const float slack = 1.0e-3f + 1.0e-5f * fmaxf(fmaxf(fabsf(w0.x) + fabsf(w1.x), fabsf(w0.y) + fabsf(w1.y)), fabsf(w0.z) + fabsf(w1.z));
        Box qb;
        qb.lo = gmin(w0, w1) - r - v3(slack);
        qb.hi = gmax(w0, w1) + r + v3(slack);

        uint32_t nleaf = 0;
        bool overflow = false;
        {
            uint32_t cur = 0, nfront = p.n_nodes ? 1u : 0u;
            if (lane == 0) s_front[0][0] = 0u;
            __syncwarp(FULL);
            while (nfront != 0 && !overflow) {
                uint32_t nnext = 0;
                for (uint32_t base = 0; base < nfront; base += 32) {
                    const uint32_t k = base + lane;
                    bool internal = false, leaf = false;
                    uint32_t idx = 0, second = 0;
                    if (k < nfront) {
                        idx = s_front[cur][k];
                        const float4 lo = __ldg(p.nodes + 2 * idx);
                        const float4 hi = __ldg(p.nodes + 2 * idx + 1);
                        ++c_node;
                        if (node_overlaps(lo, hi, qb)) {
                            const int cnt = __float_as_int(hi.w);
                            if (cnt < 0) { internal = true; second = uint32_t(-cnt); }
                            else if (!(cnt & PRT_CAPSULE_LEAF)) leaf = true;
                        }
                    }
                    const unsigned m_int = __ballot_sync(FULL, internal);
                    const unsigned m_leaf = __ballot_sync(FULL, leaf);
                    if (nnext + 2 * __popc(m_int) > SW_CAP || nleaf + __popc(m_leaf) > SW_CAP) { overflow = true; break; }
                    if (internal) {
                        const uint32_t o = nnext + 2 * __popc(m_int & lt);
                        s_front[cur ^ 1][o] = idx + 1;
                        s_front[cur ^ 1][o + 1] = second;
                    }
                    if (leaf) s_leaf[nleaf + __popc(m_leaf & lt)] = idx;
                    nnext += 2 * __popc(m_int);
                    nleaf += __popc(m_leaf);
                }
                cur ^= 1;
                nfront = nnext;
                __syncwarp(FULL);
            }
        }
        uint32_t onode = p.n_nodes;                               // resume position of the ordered passes (overflow only)
        if (overflow) { onode = 0; nleaf = 0; }

        // per-lane earliest impact
        bool found = false;
        float best_t = 0.0f;
        uint32_t best_tri = 0xFFFFFFFFu;
        V3 best_point = v3(0.0f), best_normal = v3(0.0f);
        for (;;) {
            if (overflow) {                                       // next pass of leaves: stackless walk, whole warp in step
                nleaf = 0;
                while (onode < p.n_nodes && nleaf < SW_CAP) {
                    const float4 lo = __ldg(p.nodes + 2 * onode);
                    const float4 hi = __ldg(p.nodes + 2 * onode + 1);
                    const bool ov = node_overlaps(lo, hi, qb);
                    const int cnt = __float_as_int(hi.w);
                    if (lane == 0) ++c_node;
                    if (cnt < 0) onode = ov ? onode + 1 : uint32_t(__float_as_int(lo.w));
                    else {
                        if (ov && !(cnt & PRT_CAPSULE_LEAF)) { if (lane == 0) s_leaf[nleaf] = onode; ++nleaf; }
                        onode = onode + 1;
                    }
                }
                __syncwarp(FULL);
            }
            for (uint32_t j = lane; j < nleaf; j += 32) s_tbeg[j] = uint32_t(__float_as_int(__ldg(p.nodes + 2 * s_leaf[j]).w));
            __syncwarp(FULL);
            if (lane == 0) {
                uint32_t acc = 0;
                for (uint32_t j = 0; j < nleaf; ++j) { s_start[j] = acc; acc += uint32_t(__float_as_int(__ldg(p.nodes + 2 * s_leaf[j] + 1).w)); }
                s_start[nleaf] = acc;
            }
            __syncwarp(FULL);
            const uint32_t total = s_start[nleaf];
            for (uint32_t seq = lane; seq < total; seq += 32) {
                uint32_t lo_j = 0, hi_j = nleaf;                  // leaf holding triangle #seq
                while (hi_j - lo_j > 1) { const uint32_t mid = (lo_j + hi_j) >> 1; if (s_start[mid] <= seq) lo_j = mid; else hi_j = mid; }
                const size_t t = size_t(s_tbeg[lo_j]) + (seq - s_start[lo_j]);
                const float4 t0 = __ldg(p.tris + 3 * t), t1 = __ldg(p.tris + 3 * t + 1), t2 = __ldg(p.tris + 3 * t + 2);
                ++c_tri;
                const V3 a = v3(t0.x, t0.y, t0.z) / r, b = v3(t1.x, t1.y, t1.z) / r, c = v3(t2.x, t2.y, t2.z) / r;
                const V3 nr = cross(b - a, c - a);
                if (dot(nr, nr) == 0.0f) continue;                                   // degenerate
                const V3 N = normalize(nr);
                if (!(dot(N, dir) < 0.0f)) continue;                                 // not moving towards the front face
                const float dist = dot(N, P - a);
                if (!(dist >= 0.0f)) continue;                                       // centre behind the plane
                const V3 sIP = P - N;                                                // sphere intersection point
                float tt = 0.0f;
                V3 pIP;
                if (dist <= 1.0f) {                                                  // plane cuts the sphere: along the normal
                    if (!ray_plane(a, N, sIP, N, tt)) continue;
                    pIP = sIP + N * tt;
                } else {
                    if (!ray_plane(a, N, sIP, dir, tt)) continue;
                    pIP = sIP + dir * tt;
                }
                const V3 polyIP = point_in_triangle(pIP, a, b, c) ? pIP : closest_on_perimeter(a, b, c, pIP);
                float toi = 0.0f;
                if (!ray_sphere(polyIP, -dir, P, 1.0f, toi)) continue;
                if (!(toi <= len)) continue;
                const uint32_t id = p.tri_src[t] / 3u;                               // triangle id in registration order
                if (!found || __float_as_uint(toi) < __float_as_uint(best_t) ||
                    (__float_as_uint(toi) == __float_as_uint(best_t) && id < best_tri)) {
                    found = true; best_t = toi; best_tri = id; best_point = polyIP; best_normal = N;
                }
            }
            __syncwarp(FULL);
            if (!overflow || onode >= p.n_nodes) break;
        }

        // ---- earliest time of impact over the warp: (toi bits, triangle id) as one 64-bit key, min by shuffles ----
        unsigned long long key = found ? ((unsigned long long)__float_as_uint(best_t) << 32) | best_tri : ~0ull;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long other = __shfl_xor_sync(FULL, key, o);
            key = other < key ? other : key;
        }
        if (key == ~0ull) { P = P + V; V = v3(0.0f); break; }     // free motion
        const unsigned m_win = __ballot_sync(FULL, found && ((((unsigned long long)__float_as_uint(best_t)) << 32) | best_tri) == key);
        const int src = __ffs(m_win) - 1;
        const float toi = __shfl_sync(FULL, best_t, src);
        const uint32_t tri = __shfl_sync(FULL, best_tri, src);
        const V3 point = v3(__shfl_sync(FULL, best_point.x, src), __shfl_sync(FULL, best_point.y, src), __shfl_sync(FULL, best_point.z, src));
        const V3 tri_normal = v3(__shfl_sync(FULL, best_normal.x, src), __shfl_sync(FULL, best_normal.y, src), __shfl_sync(FULL, best_normal.z, src));

        // ---- response: stop just short of the contact, slide along the tangent plane ----
        float move = toi - p.very_close;
        if (move < 0.0f) move = 0.0f;
        const V3 newP = P + dir * move;
        V3 slideN = newP - point;
        slideN = dot(slideN, slideN) == 0.0f ? tri_normal : normalize(slideN);
        const V3 dest = P + V;
        float t3 = 0.0f;
        V3 newDest = dest;
        if (ray_plane(point, slideN, dest, slideN, t3)) newDest = dest + slideN * t3;
        if (lane == 0) {
            const size_t h = size_t(q) * p.iterations + it;
            p.hit_tri[h] = int32_t(tri);
            p.hit_toi[h] = fdiv(toi, len);
            p.hit_normal[3 * h] = slideN.x; p.hit_normal[3 * h + 1] = slideN.y; p.hit_normal[3 * h + 2] = slideN.z;
        }
        V = newDest - point;
        P = newP;
    }

    const V3 w = P * r, wv = V * r;
    if (lane == 0) {
        p.out_pos[3 * q] = w.x; p.out_pos[3 * q + 1] = w.y; p.out_pos[3 * q + 2] = w.z;
        p.out_vel[3 * q] = wv.x; p.out_vel[3 * q + 1] = wv.y; p.out_vel[3 * q + 2] = wv.z;
    }
    c_tri = __reduce_add_sync(FULL, c_tri);
    c_node = __reduce_add_sync(FULL, c_node);
    if (lane == 0 && p.counters) {
        atomicAdd(p.counters + 1, (unsigned long long)c_tri);
        atomicAdd(p.counters + 2, (unsigned long long)c_node);
    }
}

} // namespace

void launch_sweep(SweepParams const & p, cudaStream_t stream) {
    if (p.n == 0) return;
    k_sweep<<<p.n, 32, 0, stream>>>(p);
}

} // namespace prt
