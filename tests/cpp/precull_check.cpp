// precull_check.cpp -- TEST program (host build of prototype2_b200/csrc/collide.cuh, no oracle, no GPU).
//
// Soundness of the two conservative pre-culls the step kernels run in front of the exact capsule-triangle test:
// whenever capsule_triangle() -- the restatement of collision_system.cpp:56-142 the kernels execute -- reports a
// contact, neither the box cull (same formula as quick_reject in step_common.cuh) nor tight_reject_rec on the
// triangle's precomputed cull record (collide.cuh; what the pooled kernel runs, at a pose displaced within its push
// budget) may have answered "skip".  Two generators:
//   sim   a crowd of capsules falling onto / sliding over the synthetic terrain of SURVEY 8d, brute-force candidates
//         from the cell grid, poses exactly as the step kernels chain them (pose rebuilt after every non-zero push);
//         also reports how many exact tests each stage leaves
//   fuzz  random triangles (fat, thin, tiny, huge, far from the origin), random capsules around them at grazing
//         distances, non-finite values sprinkled in
// usage: precull_check sim <terrain_n> <capsules> <frames> <origin_shift> | fuzz <cases> <seed>
// exit code 0 = no false reject.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>
#include "../../prototype2_b200/csrc/collide.cuh"

using namespace prt;

static bool box_reject(V3 smin, V3 smax, float radius, float dmax, float eps, V3 pa, V3 pb, V3 pc, V3 n) {
    const float ex = fmaxf(radius, dmax * fabsf(n.x)) + eps;
    const float ey = fmaxf(radius, dmax * fabsf(n.y)) + eps;
    const float ez = fmaxf(radius, dmax * fabsf(n.z)) + eps;
    const float lox = fminf(fminf(pa.x, pb.x), pc.x), hix = fmaxf(fmaxf(pa.x, pb.x), pc.x);
    const float loy = fminf(fminf(pa.y, pb.y), pc.y), hiy = fmaxf(fmaxf(pa.y, pb.y), pc.y);
    const float loz = fminf(fminf(pa.z, pb.z), pc.z), hiz = fmaxf(fmaxf(pa.z, pb.z), pc.z);
    return (lox - ex > smax.x) || (hix + ex < smin.x) || (loy - ey > smax.y) || (hiy + ey < smin.y) ||
           (loz - ez > smax.z) || (hiz + ez < smin.z);
}

static float margin_for(float world_max_abs) {            // prtc.cu make_params: 64 ulp of the largest coordinate, >= 1e-3
    float ulp = std::nextafter(world_max_abs, INFINITY) - world_max_abs;
    return std::max(1e-3f, 64.0f * ulp);
}

struct Stats { unsigned long long tris = 0, after_box = 0, after_tight = 0, contacts = 0, bad_box = 0, bad_tight = 0; };

static void probe(Pose const & ps, float radius, uint32_t abs_mode, float eps, V3 pa, V3 pb, V3 pc, Stats & st, Contact & ct, bool & hit) {
    V3 n;
    hit = false;
    if (!triangle_normal(pa, pb, pc, n)) return;
    const float dmax = abs_mode == 1u ? radius : floorf(radius) + 1.0f;
    ++st.tris;
    const bool rb = box_reject(gmin(ps.A, ps.B), gmax(ps.A, ps.B), radius, dmax, eps, pa, pb, pc, n);
    static const float mult = getenv("PRECULL_MULT") ? float(atof(getenv("PRECULL_MULT"))) : 1.0f;   // the kernels pass the first stage's margin
    // the pooled kernel culls at the pose a window STARTED with and keeps the answer while the pushes since then stay
    // within the budget `slack` (1-norm): test the record at a pose displaced by up to that much from the one the exact
    // test runs at
    static std::mt19937 drng(777);
    std::uniform_real_distribution<float> du(-1.0f, 1.0f);
    const float slack = (drng() % 3 == 0) ? 0.0f : 0.1f * radius * float(drng() % 4);
    V3 d = v3(du(drng), du(drng), du(drng));
    const float l1 = fabsf(d.x) + fabsf(d.y) + fabsf(d.z);
    d = l1 > 0 ? d * (slack * (0.999f * (drng() % 2 ? 1.0f : du(drng) * du(drng))) / l1) : v3(0.0f);
    const CullRec rec = make_cull_record(pa, pb, pc, n, false);
    const bool rt = tight_reject_rec(ps.A + d, ps.B + d, radius, dmax, mult * eps, mult * slack, v3(rec.pl[0], rec.pl[1], rec.pl[2]), rec.pl[3],
                                     v3(rec.e0[0], rec.e0[1], rec.e0[2]), rec.e0[3], v3(rec.e1[0], rec.e1[1], rec.e1[2]), rec.e1[3],
                                     v3(rec.e2[0], rec.e2[1], rec.e2[2]), rec.e2[3]);
    if (!rb) ++st.after_box;
    if (!rb && !rt) ++st.after_tight;
    hit = capsule_triangle(ps, radius, abs_mode, pa, pb, pc, n, ct);
    if (hit) {
        ++st.contacts;
        if (rb) ++st.bad_box;
        if (rt) {
            ++st.bad_tight;
            if (st.bad_tight < 5)
                fprintf(stderr, "FALSE REJECT A=(%g %g %g) B=(%g %g %g) r=%g tri=(%g %g %g)(%g %g %g)(%g %g %g)\n", ps.A.x, ps.A.y,
                        ps.A.z, ps.B.x, ps.B.y, ps.B.z, radius, pa.x, pa.y, pa.z, pb.x, pb.y, pb.z, pc.x, pc.y, pc.z);
        }
    }
}

static float terrain_h(float x, float z, float shift) { return 0.75f * sinf(0.07f * (x - shift)) * cosf(0.05f * (z - shift)); }

static int run_sim(int N, int ncaps, int frames, float shift, uint32_t abs_mode) {
    std::vector<float> h(size_t(N + 1) * (N + 1));
    for (int x = 0; x <= N; ++x) for (int z = 0; z <= N; ++z) h[size_t(x) * (N + 1) + z] = terrain_h(float(x) + shift, float(z) + shift, shift);
    auto vert = [&](int x, int z) { return v3(float(x) + shift, h[size_t(x) * (N + 1) + z], float(z) + shift); };
    CapsuleDev cap{0.5f, 0.5f, 0.0f, 0.5f, 0.0f, 0, 0, 0};
    Rot3 R; R.c0 = v3(1, 0, 0); R.c1 = v3(0, 1, 0); R.c2 = v3(0, 0, 1);
    const CapsuleFrame fr = capsule_frame(cap, R);
    const float eps = margin_for(float(N) + shift) + 1e-5f * 2.0f;
    std::mt19937 rng(12345);
    std::uniform_real_distribution<float> u(0.0f, 1.0f);
    std::vector<V3> pos(ncaps), vel(ncaps), gn(ncaps, v3(0.0f));
    std::vector<int> grounded(ncaps, 0);
    for (int i = 0; i < ncaps; ++i) {
        float x = 2 + u(rng) * (N - 4), z = 2 + u(rng) * (N - 4);
        float y = 0.75f * sinf(0.07f * x) * cosf(0.05f * z) - 0.1f + 0.4f * u(rng);
        pos[i] = v3(x + shift, y, z + shift);
        vel[i] = v3(-0.2f + 0.4f * u(rng), -0.2f * u(rng), -0.2f + 0.4f * u(rng));
    }
    const float dt = 1.0f / 30.0f, gravity = 1.0f, stick = 0.05f, friction = 10.0f;
    Stats st;
    unsigned long long substeps = 0;
    for (int f = 0; f < frames; ++f) {
        if (f == 5) { st = Stats(); substeps = 0; }       // like bench.py: skip the landing frames
        for (int i = 0; i < ncaps; ++i) {
            V3 prevVel = vel[i], v = vel[i], p = pos[i];
            if (grounded[i]) v = v + ((-stick * gravity) * dt) * gn[i];
            int g = 0;
            V3 prevA, prevB;
            segment_ends(fr, p, prevA, prevB);
            const V3 stepv = v / 4.0f;
            int left = 4;
            while (left) {
                --left;
                p = p + stepv;
                Pose ps = make_pose(fr, p);
                Box qb = box_union(capsule_box(prevA, prevB, cap.radius), capsule_box(ps.A, ps.B, cap.radius));
                prevA = ps.A; prevB = ps.B;
                ++substeps;
                int x0 = std::max(0, int(floorf(qb.lo.x - shift)) - 1), x1 = std::min(N - 1, int(floorf(qb.hi.x - shift)) + 1);
                int z0 = std::max(0, int(floorf(qb.lo.z - shift)) - 1), z1 = std::min(N - 1, int(floorf(qb.hi.z - shift)) + 1);
                bool any = false;
                for (int x = x0; x <= x1; ++x) for (int z = z0; z <= z1; ++z) {
                    V3 a = vert(x, z), b = vert(x, z + 1), c = vert(x + 1, z + 1), d = vert(x + 1, z);
                    Box cb; cb.lo = gmin(gmin(a, b), gmin(c, d)) - v3(0.05f); cb.hi = gmax(gmax(a, b), gmax(c, d)) + v3(0.05f);
                    if (cb.lo.x > qb.hi.x || cb.hi.x < qb.lo.x || cb.lo.y > qb.hi.y || cb.hi.y < qb.lo.y || cb.lo.z > qb.hi.z || cb.hi.z < qb.lo.z) continue;
                    any = true;
                    for (int t = 0; t < 2; ++t) {
                        Contact ct; bool hit;
                        probe(ps, cap.radius, abs_mode, eps, a, t ? c : b, t ? d : c, st, ct, hit);
                        if (hit) {
                            p = p + ct.impulse;
                            if (dot(ct.normal, v3(0, 1, 0)) > 0.3f) { g = 1; gn[i] = ct.normal; }
                            if (ct.impulse.x != 0 || ct.impulse.y != 0 || ct.impulse.z != 0) ps = make_pose(fr, p);
                        }
                    }
                }
                if (!any) break;
            }
            p = p + (float(left) * v) / 4.0f;
            v = prevVel;
            const float fr_ratio = 1.0f / (1.0f + dt * friction);
            v.x *= fr_ratio; v.z *= fr_ratio;
            if (g) v.y = gmax(0.0f, v.y); else v.y = v.y + (-gravity * dt);
            pos[i] = p; vel[i] = v; grounded[i] = g;
        }
    }
    printf("sim N=%d caps=%d frames=%d shift=%g abs=%u eps=%g: per substep: triangles %.2f, after box cull %.2f, after tight cull %.2f, contacts %.3f; "
           "false rejects: box %llu tight %llu\n", N, ncaps, frames, shift, abs_mode, eps, double(st.tris) / substeps,
           double(st.after_box) / substeps, double(st.after_tight) / substeps, double(st.contacts) / substeps, st.bad_box, st.bad_tight);
    return (st.bad_box || st.bad_tight) ? 1 : 0;
}

static int run_fuzz(long cases, unsigned seed) {
    std::mt19937 rng(seed);
    std::uniform_real_distribution<float> u(-1.0f, 1.0f), u01(0.0f, 1.0f);
    Stats st;
    for (long k = 0; k < cases; ++k) {
        // scene scale and offset from the origin
        const float scale = powf(10.0f, -2.0f + 5.0f * u01(rng));          // 0.01 .. 1000
        const float off = (rng() % 3 == 0) ? 0.0f : powf(10.0f, 5.0f * u01(rng)) * (rng() % 2 ? 1.f : -1.f);
        V3 o = v3(off * u01(rng), off * u01(rng), off * u01(rng));
        V3 pa = o + v3(u(rng), u(rng), u(rng)) * scale, pb = o + v3(u(rng), u(rng), u(rng)) * scale, pc = o + v3(u(rng), u(rng), u(rng)) * scale;
        switch (rng() % 8) {
            case 0: pc = pa + (pb - pa) * u01(rng) + v3(u(rng), u(rng), u(rng)) * (scale * 1e-4f); break;     // sliver
            case 1: pb = pa + v3(u(rng), u(rng), u(rng)) * (scale * 1e-3f); break;                             // short edge at pa
            case 2: pa.y = pb.y = pc.y = o.y; break;                                                           // axis aligned
            default: break;
        }
        float world_max = 0.0f;
        for (V3 q : {pa, pb, pc}) world_max = fmaxf(world_max, fmaxf(fabsf(q.x), fmaxf(fabsf(q.y), fabsf(q.z))));
        CapsuleDev cap{scale * (0.05f + 2.0f * u01(rng)), scale * (0.02f + 1.5f * u01(rng)), 0.0f, 0.0f, 0.0f, 0, 0, 0};
        if (rng() % 2) { cap.oy = cap.radius; }
        Quat q; q.w = u(rng); q.x = u(rng) * 0.5f; q.y = u(rng); q.z = u(rng) * 0.5f;
        if (rng() % 3 == 0) { q.w = 1; q.x = q.y = q.z = 0; }
        const CapsuleFrame fr = capsule_frame(cap, rotation_of(q));
        const float eps = margin_for(world_max) + 1e-5f * (fabsf(cap.radius) + fabsf(cap.height) + fabsf(cap.ox) + fabsf(cap.oy) + fabsf(cap.oz));
        // a point of the triangle (or slightly outside), lifted along a random direction by about the radius
        float b0 = u01(rng), b1 = u01(rng) * (1 - b0);
        if (rng() % 4 == 0) { b0 = 1.3f * u(rng); b1 = 1.3f * u(rng); }
        V3 base = pa + (pb - pa) * b0 + (pc - pa) * b1;
        V3 nr = cross(pb - pa, pc - pa);
        float nl = sqrtf(dot(nr, nr));
        V3 dir = nl > 0 ? nr * (1.0f / nl) : v3(0, 1, 0);
        dir = dir + v3(u(rng), u(rng), u(rng)) * (rng() % 2 ? 0.05f : 1.0f);
        const float lift[6] = {0.0f, 0.5f, 0.98f, 1.0f, 1.02f, 1.6f};
        V3 p = base + dir * (cap.radius * lift[rng() % 6] * (1.0f + 0.01f * u(rng)));
        // put the capsule's lower (or upper) sphere centre near p
        V3 A0, B0;
        segment_ends(fr, v3(0.0f), A0, B0);
        V3 pos = p - ((rng() % 2) ? A0 : B0) + (B0 - A0) * ((rng() % 3 == 0) ? u01(rng) : 0.0f) * ((rng() % 2) ? 1.f : -1.f);
        if (rng() % 500 == 0) pos.y = NAN;
        if (rng() % 500 == 0) pos.x = INFINITY;
        if (rng() % 500 == 0) pb.z = NAN;
        if (rng() % 500 == 0) pa.x = -INFINITY;
        Pose ps = make_pose(fr, pos);
        Contact ct; bool hit;
        probe(ps, cap.radius, rng() % 2, eps, pa, pb, pc, st, ct, hit);
    }
    printf("fuzz cases=%ld seed=%u: non-degenerate %llu, after box cull %llu, after tight cull %llu, contacts %llu; false rejects: box %llu tight %llu\n",
           cases, seed, st.tris, st.after_box, st.after_tight, st.contacts, st.bad_box, st.bad_tight);
    return (st.bad_box || st.bad_tight) ? 1 : 0;
}

int main(int argc, char ** argv) {
    if (argc >= 6 && !strcmp(argv[1], "sim"))
        return run_sim(atoi(argv[2]), atoi(argv[3]), atoi(argv[4]), float(atof(argv[5])), argc > 6 ? uint32_t(atoi(argv[6])) : 0u);
    if (argc >= 4 && !strcmp(argv[1], "fuzz")) return run_fuzz(atol(argv[2]), unsigned(atoi(argv[3])));
    fprintf(stderr, "usage: precull_check sim <terrain_n> <capsules> <frames> <origin_shift> [abs_mode] | fuzz <cases> <seed>\n");
    return 2;
}
