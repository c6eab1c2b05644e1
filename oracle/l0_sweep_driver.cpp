// ORACLE / TEST INFRASTRUCTURE ONLY -- never linked into, imported by or executed from the product path.
//
// Checker for the Tier-B kernel (prtc_sweep_ellipsoids, SURVEY.md section 0 finding 1 / 8f row f4): the swept
// ellipsoid-vs-triangle collide-and-slide that BASELINE.json's north_star words.  PARITY UNPINNED: this snapshot of the
// reference has no caller for it -- PhysicsSystem::collisionResponse(intersectionPoint, collisionNormal,
// intersectionTime, ...) is declared and never defined (physics_system.h:131-136) -- only the helper functions of the
// older implementation survive ("from Generic Collision Detection for Games Using Ellipsoids by Paul Nettle",
// src/util/physics_util.cpp:51-178).  So:
//   * the HELPERS are the reference's own: physics_util::intersectRayPlane / intersectRaySphere / checkPointInTriangle /
//     closestPointOnTrianglePerimeter, compiled UNMODIFIED from /root/reference (oracle/Makefile -> _ref/libprt_l0_sweep.so)
//   * the LOOP around them below is our restatement of Nettle's algorithm as those helpers' signatures imply it, and
//     is the definition the CUDA kernel is held to, bit for bit.  It brute-forces every triangle (no tree).
#include <cstdint>
#include <cstring>
#include <cmath>

#include <glm/glm.hpp>
#include <glm/gtx/norm.hpp>

#include "src/util/physics_util.h"

namespace {
glm::vec3 ld3(const float * p) { return glm::vec3(p[0], p[1], p[2]); }
glm::vec3 div3(glm::vec3 const & a, glm::vec3 const & b) { return glm::vec3(a.x / b.x, a.y / b.y, a.z / b.z); }   // component-wise, like glm's vec3 / vec3
glm::vec3 neg3(glm::vec3 const & a) { return glm::vec3(-a.x, -a.y, -a.z); }
uint32_t bits(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
}

// tri_xyz: world-space triangles, 9 floats each, in REGISTRATION order (their index is the hit id).
// hit_* hold one entry per (query, iteration): triangle id or -1, time of impact in [0,1] of that iteration's motion,
// slide-plane normal in ellipsoid space.
extern "C" void l0_sweep(const float * tri_xyz, uint32_t n_tris, uint32_t n, const float * centers, const float * radii,
                         const float * velocities, uint32_t iterations, float very_close, float * out_pos, float * out_vel,
                         int32_t * hit_tri, float * hit_toi, float * hit_normal, uint64_t * tri_tests) {
    uint64_t tests = 0;
    for (uint32_t q = 0; q < n; ++q) {
        glm::vec3 const r = ld3(radii + 3 * q);
        glm::vec3 P = div3(ld3(centers + 3 * q), r);                 // ellipsoid space: the ellipsoid is the unit sphere
        glm::vec3 V = div3(ld3(velocities + 3 * q), r);
        for (uint32_t it = 0; it < iterations; ++it) { hit_tri[q * iterations + it] = -1; hit_toi[q * iterations + it] = 0.0f;
            hit_normal[3 * (q * iterations + it)] = hit_normal[3 * (q * iterations + it) + 1] = hit_normal[3 * (q * iterations + it) + 2] = 0.0f; }
        for (uint32_t it = 0; it < iterations; ++it) {
            float const len = glm::length(V);
            if (!(len > 0.0f)) break;
            glm::vec3 const dir = glm::normalize(V);
            bool found = false;
            float best_t = 0.0f;
            uint32_t best_tri = 0;
            glm::vec3 best_point(0.0f), best_normal(0.0f);
            for (uint32_t t = 0; t < n_tris; ++t) {
                ++tests;
                glm::vec3 const a = div3(ld3(tri_xyz + 9 * size_t(t)), r), b = div3(ld3(tri_xyz + 9 * size_t(t) + 3), r), c = div3(ld3(tri_xyz + 9 * size_t(t) + 6), r);
                glm::vec3 const nr = glm::cross(b - a, c - a);
                if (glm::dot(nr, nr) == 0.0f) continue;                              // degenerate
                glm::vec3 const N = glm::normalize(nr);
                if (!(glm::dot(N, dir) < 0.0f)) continue;                            // not moving towards the front face
                float const dist = glm::dot(N, P - a);
                if (!(dist >= 0.0f)) continue;                                       // centre behind the plane
                glm::vec3 const sIP = P - N;                                         // sphere intersection point
                float t1 = 0.0f;
                glm::vec3 pIP;
                if (dist <= 1.0f) {                                                  // plane cuts the sphere: along the normal
                    if (!physics_util::intersectRayPlane(a, N, sIP, N, t1)) continue;
                    pIP = sIP + N * t1;
                } else {
                    if (!physics_util::intersectRayPlane(a, N, sIP, dir, t1)) continue;
                    pIP = sIP + dir * t1;
                }
                glm::vec3 const polyIP = physics_util::checkPointInTriangle(pIP, a, b, c) ? pIP
                                         : physics_util::closestPointOnTrianglePerimeter(a, b, c, pIP);
                float t2 = 0.0f;
                if (!physics_util::intersectRaySphere(polyIP, neg3(dir), P, 1.0f, t2)) continue;
                if (!(t2 <= len)) continue;
                // earliest impact; equal times -> lowest triangle id
                if (!found || bits(t2) < bits(best_t)) { found = true; best_t = t2; best_tri = t; best_point = polyIP; best_normal = N; }
            }
            if (!found) { P = P + V; V = glm::vec3(0.0f); break; }
            float move = best_t - very_close;
            if (move < 0.0f) move = 0.0f;
            glm::vec3 const newP = P + dir * move;
            glm::vec3 slideN = newP - best_point;
            slideN = glm::dot(slideN, slideN) == 0.0f ? best_normal : glm::normalize(slideN);
            glm::vec3 const dest = P + V;
            float t3 = 0.0f;
            glm::vec3 newDest = dest;
            if (physics_util::intersectRayPlane(best_point, slideN, dest, slideN, t3)) newDest = dest + slideN * t3;
            hit_tri[q * iterations + it] = int32_t(best_tri);
            hit_toi[q * iterations + it] = best_t / len;
            hit_normal[3 * (q * iterations + it)] = slideN.x; hit_normal[3 * (q * iterations + it) + 1] = slideN.y; hit_normal[3 * (q * iterations + it) + 2] = slideN.z;
            V = newDest - best_point;
            P = newP;
        }
        glm::vec3 const w = P * r, wv = V * r;
        out_pos[3 * q] = w.x; out_pos[3 * q + 1] = w.y; out_pos[3 * q + 2] = w.z;
        out_vel[3 * q] = wv.x; out_vel[3 * q + 1] = wv.y; out_vel[3 * q + 2] = wv.z;
    }
    if (tri_tests) *tri_tests = tests;
}
