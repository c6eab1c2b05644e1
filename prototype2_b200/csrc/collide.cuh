// collide.cuh -- device-side narrowphase: one capsule against one triangle / one capsule, with the reference's
// exact arithmetic, branch structure and quirks.
//
// Restates (for the GPU) reference src/game/system/physics/collision_system.cpp:19-156 (collideCapsuleMesh),
// :158-234 (collideCapsuleCapsule), :236-252 (handleCollision, COLLIDE branch) and
// src/util/physics_util.cpp:5-49 (closestPointOnTriangle), :125-139 (closestPointOnLineSegment).
// What is hoisted out of the reference's per-triangle loop is hoisted only where the hoisted value is a pure
// function of unchanged inputs, so every float that reaches a branch or an output is bit-identical:
//   * the capsule segment (A, B, CapsuleNormal, pointA, segment direction and length) depends only on the
//     capsule and the current position -> kept in `Pose`, recomputed after every non-zero push
//   * the unit plane normal of a triangle depends only on the triangle -> precomputed once per commit by
//     k_tri_setup with the same instruction sequence (stored in the .w lanes of the triangle record)
#ifndef PRT_COLLIDE_CUH
#define PRT_COLLIDE_CUH

#include "prt_math.cuh"

namespace prt {

#define PRT_CAPSULE_LEAF 0x40000000      /* leaf `count` flag: the leaf is a capsule of the unified (LITERAL) tree, link = capsule id */
#define PRT_DEGENERATE_BITS 0x7FC0DEADu   /* NaN payload the FPU never produces: marks length(n)==0 triangles */

struct CapsuleDev {          // CapsuleCollider (colliders.h:14-18), padded to 32 bytes
    float height, radius;
    float ox, oy, oz;
    float pad0, pad1, pad2;
};

// Per-capsule constants of `tform * a` / `tform * b` (collision_system.cpp:37-43) with tform = [R | pos]:
//   A = (R0*a.x + R1*a.y) + (R2*a.z + pos)      -> A = sa1 + (sa2 + pos)   (+ pos*0, see segment_ends)
struct CapsuleFrame {
    V3 sa1, sa2, sb1, sb2;
    float radius;
};
PRT_HD CapsuleFrame capsule_frame(CapsuleDev const & c, Rot3 const & r) {
    CapsuleFrame f;
    V3 a = v3(c.ox, c.oy, c.oz);
    V3 b = a + v3(0.0f, c.height, 0.0f);             // capsule.offset + vec3{0, height, 0}
    f.sa1 = r.c0 * a.x + r.c1 * a.y; f.sa2 = r.c2 * a.z;
    f.sb1 = r.c0 * b.x + r.c1 * b.y; f.sb2 = r.c2 * b.z;
    f.radius = c.radius;
    return f;
}

// CapsuleCollider::getAABB (colliders.h:19-27) for already transformed segment ends
PRT_HD Box capsule_box(V3 tA, V3 tB, float radius) {
    Box b;
    b.lo = gmin(tA, tB) - v3(radius);
    b.hi = gmax(tA, tB) + v3(radius);
    return b;
}

struct Pose {
    V3 A, B;                 // segment end points in world space
    V3 cn;                   // CapsuleNormal = normalize(A - B)                 :45
    V3 pointA;               // A - cn * radius                                   :46-48
    V3 segv; float segd;     // normalize(B - A), distance(A, B): closestPointOnLineSegment(A, B, .)   :67
};
// `cols` is the position tform was built from.  translate(mat4(1), cols) computes its translation column as
// ((m0*x + m1*y) + m2*z) + m3 with identity columns, i.e. every component picks up 0*x, 0*y, 0*z terms, and the 4x4
// product with toMat4(q) then adds translation*0 to every rotation entry.  For finite positions all of that is +-0;
// for a position with ANY inf/NaN component it is NaN everywhere -- which is how a character at y = inf gets an
// all-NaN box (and therefore every mesh as candidate) in the reference.  `zs` carries exactly that.  `t` is the
// matrix's translation column (== cols except for glm::translate(tform, v), physics_system.cpp:271).
PRT_HD void segment_ends(CapsuleFrame const & f, V3 cols, V3 t, V3 & A, V3 & B) {
    const float zs = (cols.x * 0.0f + cols.y * 0.0f) + cols.z * 0.0f;
    A = (f.sa1 + (f.sa2 + t)) + zs;
    B = (f.sb1 + (f.sb2 + t)) + zs;
}
PRT_HD void segment_ends(CapsuleFrame const & f, V3 pos, V3 & A, V3 & B) { segment_ends(f, pos, pos, A, B); }
PRT_HD Pose make_pose(CapsuleFrame const & f, V3 pos) {
    Pose p;
    segment_ends(f, pos, p.A, p.B);
    // collision_system.cpp:47 normalises A - B, closestPointOnLineSegment (physics_util.cpp:125-139) normalises B - A and
    // takes distance(A, B): the squared length is the same float either way ((-x)(-x) == x x, same order of the sums), so
    // one square root and one reciprocal serve both -- normalize(v) = v * (1 / sqrt(dot(v, v))) (glm::inversesqrt)
    V3 ab = p.A - p.B;
    V3 ba = p.B - p.A;
    float dd = dot(ba, ba);
    float s = fsqrt(dd);
    float inv = fdiv(1.0f, s);
    p.cn = ab * inv;
    p.pointA = p.A - p.cn * f.radius;
    p.segv = ba * inv;
    p.segd = s;
    return p;
}

// physics_util.cpp:125-139 with v = normalize(b - a) and d = distance(a, b) supplied
PRT_HD V3 closest_on_segment(V3 a, V3 b, V3 v, float d, V3 p) {
    V3 c = p - a;
    float t = dot(v, c);
    if (t < 0) return a;
    if (t > d) return b;
    return a + v * t;
}
PRT_HD V3 closest_on_segment(V3 a, V3 b, V3 p) {
    V3 ba = b - a;
    float s = fsqrt(dot(ba, ba));
    return closest_on_segment(a, b, ba * fdiv(1.0f, s), s, p);
}

// physics_util.cpp:5-49
PRT_HD V3 closest_on_triangle(V3 a, V3 b, V3 c, V3 p) {
    V3 ab = b - a;
    V3 ac = c - a;
    V3 bc = c - b;
    float snom = dot(p - a, ab), sdenom = dot(p - b, a - b);
    float tnom = dot(p - a, ac), tdenom = dot(p - c, a - c);
    if (snom <= 0.0f && tnom <= 0.0f) return a;
    float unom = dot(p - b, bc), udenom = dot(p - c, b - c);
    if (sdenom <= 0.0f && unom <= 0.0f) return b;
    if (tdenom <= 0.0f && udenom <= 0.0f) return c;
    V3 n = cross(b - a, c - a);
    float vc = dot(n, cross(a - p, b - p));
    if (vc <= 0.0f && snom >= 0.0f && sdenom >= 0.0f) return a + fdiv(snom, snom + sdenom) * ab;
    float va = dot(n, cross(b - p, c - p));
    if (va <= 0.0f && unom >= 0.0f && udenom >= 0.0f) return b + fdiv(unom, unom + udenom) * bc;
    float vb = dot(n, cross(c - p, a - p));
    if (vb <= 0.0f && tnom >= 0.0f && tdenom >= 0.0f) return a + fdiv(tnom, tnom + tdenom) * ac;
    float u, v;
    div2(va, vb, (va + vb) + vc, u, v);
    float w = (1.0f - u) - v;
    return (u * a + v * b) + w * c;
}

// unit plane normal as collision_system.cpp:52-54 computes it; returns false for the degenerate skip
PRT_HD bool triangle_normal(V3 pa, V3 pb, V3 pc, V3 & n) {
    V3 nr = cross(pb - pa, pc - pa);
    float dd = dot(nr, nr);
    if (fsqrt(dd) == 0.0f) return false;
    n = nr * fdiv(1.0f, fsqrt(dd));
    return true;
}

struct Contact {
    V3 impulse;
    V3 normal;
};

// collision_system.cpp:56-142 for one triangle with unit normal n.  Returns true when the reference would call
// handleCollision (`inside || intersects`), filling the contact.
PRT_HD bool capsule_triangle(Pose const & ps, float radius, uint32_t abs_mode, V3 pa, V3 pb, V3 pc, V3 n, Contact & out) {
    float nDotCn = dot(n, ps.cn);
    V3 reference_point;
    if (fabsf(nDotCn) > 0.0001f) {
        float t = dot(n, (pa - ps.pointA) / fabsf(nDotCn));
        V3 line_plane_intersection = ps.pointA + ps.cn * t;
        reference_point = closest_on_triangle(pa, pb, pc, line_plane_intersection);
    } else {
        reference_point = pa;
    }
    V3 center = closest_on_segment(ps.A, ps.B, ps.segv, ps.segd, reference_point);

    float dist = dot(center - pa, n);
    if (dist < 0.0f) return false;                              // back-face cull                  :71
    if (abs_dist(dist, abs_mode) > radius) return false;       // range cull (see prtc_abs_mode)  :72

    V3 point0 = center - n * dist;
    V3 c0 = cross(point0 - pa, pb - pa);
    V3 c1 = cross(point0 - pb, pc - pb);
    V3 c2 = cross(point0 - pc, pa - pc);
    bool inside = dot(c0, n) <= 0 && dot(c1, n) <= 0 && dot(c2, n) <= 0;

    float radiussq = radius * radius;
    // the three closestPointOnLineSegment calls (:82-99) with their edge lengths and inverse lengths computed together
    // (same values as closest_on_segment(a, b, p) computes one edge at a time; see len_inv3)
    const V3 e1 = pb - pa, e2 = pc - pb, e3 = pa - pc;
    float s1, s2, s3, i1, i2, i3;
    len_inv3(dot(e1, e1), dot(e2, e2), dot(e3, e3), s1, s2, s3, i1, i2, i3);
    V3 point1 = closest_on_segment(pa, pb, e1 * i1, s1, center);
    V3 v1 = center - point1;
    float distsq1 = dot(v1, v1);
    bool intersects = distsq1 < radiussq;
    V3 point2 = closest_on_segment(pb, pc, e2 * i2, s2, center);
    V3 v2 = center - point2;
    float distsq2 = dot(v2, v2);
    intersects |= distsq2 < radiussq;
    V3 point3 = closest_on_segment(pc, pa, e3 * i3, s3, center);
    V3 v3_ = center - point3;
    float distsq3 = dot(v3_, v3_);
    intersects |= distsq3 < radiussq;

    if (!(inside || intersects)) return false;

    V3 intersection_vec;
    if (inside) {
        intersection_vec = center - point0;
    } else {
        // :108-127 -- `distsq = best_distsq;` (sic, twice): best_distsq stays at edge 1's value, so edge 3
        // wins over edge 2 whenever both beat edge 1
        float best_distsq = distsq1;
        intersection_vec = v1;
        if (distsq2 < best_distsq) intersection_vec = v2;
        if (distsq3 < best_distsq) intersection_vec = v3_;
    }
    float len = length(intersection_vec);
    if (len > 0.0f && len < radius) {
        float penetration_depth = radius - len;
        V3 penetration_normal = intersection_vec / len;
        out.impulse = penetration_depth * (penetration_normal + 0.0001f);   // scalar added to every component :136
        out.normal = penetration_normal;
    } else {
        out.impulse = v3(0.0f);
        out.normal = n;
    }
    return true;
}

// Conservative plane / edge-plane pre-cull of the pooled kernel (k_step_pool), on a per-triangle CULL RECORD that
// k_tri_setup precomputes next to the triangle record (4 float4 = 64 bytes):
//     pl = (n.xyz, d0 = n . pa)        e_k = (m_k.xyz, d_k = m_k . v_k)   for the three edges, m_k = (e_k x n) / |e_k|
// (the in-plane unit normal of edge k pointing OUT of the triangle, v_k a vertex on that edge).
//
// capsule_triangle() reports a contact only through `inside || intersects` (collision_system.cpp:101), and both put
// `center` -- a point of the capsule segment AB (closestPointOnLineSegment, :67) -- close to the triangle:
//   intersects -> a point of an edge lies within `radius` of center                                   (:82-99)
//   inside     -> the projection of center onto the plane lies inside all three edge planes and the plane distance
//                 is in [0, dmax)  (back-face and range culls :71-72; dmax = radius, or floor(radius) + 1 in int-abs mode)
// Every point Q of the triangle has (Q - v_k) . m_k <= 0 and (Q - pa) . n == 0, and both forms are linear along AB, so
// for every centre P on AB
//   plane distance      dlo <= (P - pa) . n <= dhi          (dlo/dhi = min/max over the two end points)
//   |P - Q|^2           >=  max(0, dlo)^2 + max(0, s)^2     (s = max over the edges of min over the end points of
//                                                            (. - v_k) . m_k)
// Hence no contact is possible at this pose when
//   dlo > dmax + eps,   or   dhi < -eps,   or   (s > eps  and  max(0, dlo)^2 + s^2 > (radius + eps)^2)
// `eps` absorbs the rounding of the exact test's intermediate points and of this test (>= 64 ulp of the largest world
// coordinate, StepParams::qr_eps).  `slack` makes the answer hold for EVERY pose whose translation differs from the
// tested one by at most `slack` (in any norm <= the 1-norm): both linear forms move by at most that much, so dlo, dhi
// and s are widened by it.  The pooled kernel culls a window of triangles once, at the pose the window starts with, and
// keeps the answers while the pushes applied since then stay inside that budget.
// The edge-plane argument needs a stored normal that really is perpendicular to the edges, which rounding cannot
// guarantee for slivers: the record carries edge planes only for triangles whose squared area is at least
// 2^-4 |ab|^2 |ac|^2 (sine of the angle at pa >= 1/4; the normal of such a triangle is tilted by < 1e-6 rad), all-zero
// edge entries otherwise (s == 0: never rejected by the edges).  A degenerate triangle (length(n) == 0, skipped by
// the reference :53) is marked in pl.x.  NaNs anywhere make dA or dB NaN => "not rejected".  MUFU.RSQ / fused
// multiply-adds are fine here: nothing computed in these functions reaches an output.  Soundness is checked on the CPU
// by tests/test_precull.py (host build of this header, random poses displaced within the slack).
PRT_HD float rsq_approx(float x) {
#ifdef __CUDA_ARCH__
    return rsqrtf(x);
#else
    return 1.0f / sqrtf(x);
#endif
}
PRT_HD float fdot(V3 a, V3 b) { return fmaf(a.x, b.x, fmaf(a.y, b.y, a.z * b.z)); }
PRT_HD V3 fcross(V3 a, V3 b) {
    return v3(fmaf(a.y, b.z, -(a.z * b.y)), fmaf(a.z, b.x, -(a.x * b.z)), fmaf(a.x, b.y, -(a.y * b.x)));
}
struct CullRec { float pl[4], e0[4], e1[4], e2[4]; };
PRT_HD CullRec make_cull_record(V3 pa, V3 pb, V3 pc, V3 n, bool degenerate) {
    CullRec r;
    for (int k = 0; k < 4; ++k) { r.pl[k] = 0.0f; r.e0[k] = 0.0f; r.e1[k] = 0.0f; r.e2[k] = 0.0f; }
    if (degenerate) {
        union { uint32_t u; float f; } mark; mark.u = PRT_DEGENERATE_BITS;
        r.pl[0] = mark.f;
        return r;
    }
    r.pl[0] = n.x; r.pl[1] = n.y; r.pl[2] = n.z; r.pl[3] = fdot(n, pa);
    const V3 eab = pb - pa, ebc = pc - pb, eca = pa - pc;
    const float lab = fdot(eab, eab), lbc = fdot(ebc, ebc), lca = fdot(eca, eca);
    const V3 mab = fcross(eab, n);
    // conditioning: |ab x ac|^2 >= 2^-4 |ab|^2 |ac|^2, with ab x ac reconstructed as (mab . ca) n-wards: |mab . eca| is
    // twice the area when n is the true unit normal, so this also fails for a normal that is not
    const float twice_area = fdot(mab, eca);
    if (!(twice_area * twice_area >= 0.0625f * (lab * lca)) || !(twice_area > 0.0f)) return r;
    const V3 mbc = fcross(ebc, n), mca = fcross(eca, n);
    const float iab = rsq_approx(lab), ibc = rsq_approx(lbc), ica = rsq_approx(lca);
    const V3 uab = mab * iab, ubc = mbc * ibc, uca = mca * ica;
    const float dab = fdot(uab, pa), dbc = fdot(ubc, pb), dca = fdot(uca, pa);
    const float all = ((uab.x + uab.y) + (uab.z + dab)) + ((ubc.x + ubc.y) + (ubc.z + dbc)) + ((uca.x + uca.y) + (uca.z + dca));
    if (!(all - all == 0.0f)) return r;                          // something overflowed / is not finite: no edge planes
    r.e0[0] = uab.x; r.e0[1] = uab.y; r.e0[2] = uab.z; r.e0[3] = dab;
    r.e1[0] = ubc.x; r.e1[1] = ubc.y; r.e1[2] = ubc.z; r.e1[3] = dbc;
    r.e2[0] = uca.x; r.e2[1] = uca.y; r.e2[2] = uca.z; r.e2[3] = dca;
    return r;
}
// the record's four rows as they are loaded: (x, y, z, w)
PRT_HD bool tight_reject_rec(V3 A, V3 B, float radius, float dmax, float eps, float slack,
                             V3 n, float d0, V3 m0, float d0e, V3 m1, float d1e, V3 m2, float d2e) {
    const float dA = fdot(A, n) - d0, dB = fdot(B, n) - d0;
    if (!(dA == dA) || !(dB == dB)) return false;
    const float dlo = fminf(dA, dB) - slack, dhi = fmaxf(dA, dB) + slack;
    if (dlo > dmax + eps || dhi < -eps) return true;
    const float s0 = fminf(fdot(A, m0), fdot(B, m0)) - d0e;
    const float s1 = fminf(fdot(A, m1), fdot(B, m1)) - d1e;
    const float s2 = fminf(fdot(A, m2), fdot(B, m2)) - d2e;
    const float s = fmaxf(fmaxf(s0, s1), s2) - slack;
    if (!(s > eps)) return false;
    const float dp = fmaxf(dlo, 0.0f);
    const float reach = radius + eps;
    return fmaf(dp, dp, s * s) > reach * reach;
}

// collision_system.cpp:164-221: capsule A (pose endpoints a_A, a_B) against capsule B (b_A, b_B).
PRT_HD bool capsule_capsule(V3 a_A, V3 a_B, float radiusA, V3 b_A, V3 b_B, float radiusB, Contact & out) {
    V3 v0 = b_A - a_A;
    V3 v1 = b_B - a_A;
    V3 v2 = b_A - a_B;
    V3 v3_ = b_B - a_B;
    float d0 = dot(v0, v0);
    float d1 = dot(v1, v1);
    float d2 = dot(v2, v2);
    float d3 = dot(v3_, v3_);
    V3 bestA;
    if (d2 < d0 || d2 < d1 || d3 < d0 || d3 < d1) bestA = a_B;
    else bestA = a_A;
    V3 bestB = closest_on_segment(b_A, b_B, bestA);
    bestA = closest_on_segment(a_A, a_B, bestB);
    V3 penetration_normal = bestA - bestB;
    float len = length(penetration_normal);
    penetration_normal = penetration_normal / len;            // len == 0 -> NaN, unguarded in the reference
    float penetration_depth = (radiusA + radiusB) - len;
    if (!(penetration_depth > 0)) return false;
    out.impulse = penetration_depth * (penetration_normal + 0.0001f);
    out.normal = penetration_normal;
    return true;
}

} // namespace prt
#endif
