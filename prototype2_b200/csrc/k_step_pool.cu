// k_step_pool.cu -- the static pass (and the JACOBI dynamic pass) at scale: 32 characters per warp, their candidate
// triangles POOLED and tested by all 32 lanes.
// Compile with -fmad=false: results must be bit-identical to the CPU reference (see prt_math.cuh).
//
// Why.  k_step_stream gives every character one lane; a lane filters its own triangle stream and the warp converges for
// the exact tests.  ncu (profiles/r01_final_*): 13.9 of 32 lanes active per instruction -- the per-lane filter loop runs
// with 8.8 lanes (runs of rejected triangles are geometric and end with the slowest lane), the exact test with ~20, and a
// character's frame has a latency of ~0.2 ms, which is what limits a 125 k-character shard on one of eight GPUs.
//
// Here a lane OWNS a character (state, substep bookkeeping) but the triangle WORK of the warp's 32 characters is one
// pool that all lanes execute, in two dense stages:
//   jobs     every owner with candidates left contributes a window of its next triangles, in the reference's test order;
//            the windows are laid end to end (owner-major), job j is found from its index alone: the owner by counting
//            window starts (one redux + popc), the triangle from the owner's <= 4 leaf pieces -- nothing is materialised;
//   CULL     lane k takes job k, k+32, ...: the triangle's 64-byte cull record (plane + three edge planes, precomputed by
//            k_tri_setup) arrives through a cp.async ring two trips ahead; tight_reject_rec (collide.cuh) against the pose
//            the owner's window STARTED with, widened by the push budget; ballots compact the survivors into a ring, order kept;
//   EXACT    whenever 32 survivors wait: lane k takes survivor k -- the reference's capsule-triangle test at the owner's
//            CURRENT pose (published in shared memory);
//   COMMIT   speculative and ordered (the rule k_step_warp already uses): a character's survivors sit in contiguous lanes;
//            contacts up to and including its first non-zero push are committed in order (position += impulse, last ground
//            contact wins, collision_system.cpp:242-252); the owner moves, rebuilds and publishes its pose at once; its
//            later survivors of that trip saw a stale pose and stay at the head of the ring for the next trip.  A
//            zero-impulse contact does not move the capsule (:132-141), so this is exactly the sequential Gauss-Seidel
//            loop of collision_system.cpp:29-155.
// Push budget: a cull answer holds for every pose within `push_slack` of the one it was computed at (collide.cuh), so a
// window is culled ONCE; only an owner whose pushes add up to more than the budget has its window cut behind the pushing
// triangle and the rest culled again at the new pose (next round).
// The window grows when few owners are still busy (JOB_CAP / busy owners), which keeps the tail of a substep short.
// Substeps run in lock step across the warp's characters (LOAD, SUBSTEP set-up and STORE are dense, lane = character).
// Results, counters (substeps, triangle tests, contacts) and candidate order are those of k_step / k_step_stream.
#include "step_common.cuh"

#ifndef PRTC_WITH_TIMELINE
#define PRTC_WITH_TIMELINE 0          // 1: k_step_pool records a claim timeline when StepParams::timeline is set (tools/claim_timeline.py)
#endif

namespace prt {

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int PL_LEAF_CAP = LEAF_CAP;   // candidate leaves buffered per owner and pass (more => another pass of the ordered walk)
constexpr int PL_JOB_CAP = 1024;        // jobs per round and warp (what bounds a window when few owners are busy)
constexpr int PL_STAGES = 2;            // cull trips whose records are in flight
constexpr int PL_PIECES = 4;            // leaf pieces a window may span
constexpr uint32_t PL_TRI_MASK = (1u << 27) - 1u;      // job word: owner lane << 27 | triangle index

enum PoseField { PF_A = 0, PF_B = 3, PF_CN = 6, PF_PA = 9, PF_SV = 12, PF_SEGD = 15, PF_RADIUS = 16, PF_DMAX = 17, PF_EPS = 18,
                 PF_A0 = 19, PF_B0 = 22, PF_COUNT = 25 };

template <bool JACOBI>
struct PoolShared {
    float4 ring[PL_STAGES][4][32];         // cull records in flight (cp.async), one column per lane
    uint32_t r2_e[64], r2_q[64];           // survivors of the cull waiting for the exact test: job word, position in the owner's window
    uint32_t off[32];                      // the owners' offsets in the round's job sequence (exclusive prefix sum of the windows)
    uint32_t rk2o[32];                     // k-th owner with a window this round
    uint32_t bnd[PL_PIECES - 1][32];       // window positions where the owner's 2nd, 3rd, 4th leaf piece starts
    uint32_t fst[PL_PIECES][32];           // first triangle of each piece
    float pose[PF_COUNT][32];              // the owners' current poses (+ the segment their window was culled at), [field][owner]
    uint32_t lfirst[PL_LEAF_CAP][32];      // candidate leaves of the substep: first triangle / number of triangles
    uint32_t lcount[PL_LEAF_CAP][32];
    uint32_t mb_flags[32];                 // commit mailbox per owner: 1 zero-impulse contact, 2 ground contact, 4 push
    uint32_t mb_q[32];                     // position of the pushing triangle in the owner's window
    uint32_t inval[32];                    // the owner's window was cut (push budget spent): its later jobs and survivors are void
    float mb_imp[3][32], mb_gn[3][32];
    uint32_t dyn[JACOBI ? DYN_CAP : 1][32];   // JACOBI: neighbour capsules of the pass, ascending
};

__device__ __forceinline__ V3 ld3(const float (*a)[32], int f, uint32_t o) { return v3(a[f][o], a[f + 1][o], a[f + 2][o]); }
__device__ __forceinline__ void st3(float (*a)[32], int f, uint32_t o, V3 v) { a[f][o] = v.x; a[f + 1][o] = v.y; a[f + 2][o] = v.z; }

// CHAINED (static pass, PRTC_OPT_CHAIN_FRAMES): consecutive frames of the same batch overlap on the device.  The launch
// carries the programmatic-stream-serialization attribute and every block lets the next launch in as soon as it runs (a
// launch starts when every block of the one before it HAS STARTED -- not finished), so later frames' blocks take the warp
// slots that earlier frames leave empty or never needed; several frames can be on the machine at once.  What orders the
// frames is a word per claim: claim g of frame f+1 starts when claim g of frame f (the same characters) has stored its
// results (StepParams::chain_done).  Nothing may sit between the two step kernels, so this variant brings the per-slot leaf
// lists up to date itself at LOAD (the check and, rarely, the walk of k_xc_refresh).  What another frame wrote -- the records,
// the slot arrays -- is read with plain loads behind an acquire of the claim's word (the other side: release store behind
// the warp barrier), which is what the memory model asks for; reading them past the L1 instead (ld.global.cg) cost 6 %.
template <bool GATED, int MODE, bool CHAINED>
__global__ void __launch_bounds__(32, 16) k_step_pool(StepParams p) {
    constexpr bool JACOBI = MODE == MODE_JACOBI;
    __shared__ PoolShared<JACOBI> sh;
    const uint32_t lane = threadIdx.x;
    const unsigned lt = (1u << lane) - 1u;
    uint32_t c_sub = 0, c_tri = 0, c_node = 0, c_contact = 0, c_exact = 0, c_cc = 0;
    const bool use_qr = (p.qr_enable & 1u) != 0u, use_tq = (p.qr_enable & 2u) != 0u;
    const bool use_cache = !(p.tune & 1u) && p.xc_version != nullptr;
    const uint32_t wmax = ((p.tune >> 8) & 0xFFu) ? ((p.tune >> 8) & 0xFFu) : 32u;   // largest window (jobs per owner and round)
    if (p.work_next && blockIdx.x == 0 && lane == 0) *p.work_next = 0u;      // the NEXT launch's work counter (see StepParams::work_pair)
    if (CHAINED) {
        // (the next frame's counter is zero before any of its blocks can run: they start only after EVERY block of this
        //  launch has passed this point)
        __threadfence();
        __syncwarp(FULL);
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    }
    const float fsub = float(p.substeps);
    // characters per warp: 32 when the batch fills the machine; fewer owners (lanes [0, opw)) spread a smaller batch over
    // more warps -- the pooled work is executed by all 32 lanes either way
    const uint32_t opw = GATED ? 32u : (p.lanes_per_warp ? p.lanes_per_warp : 32u);

    for (;;) {
        // ---------------- claim `opw` consecutive characters ----------------
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(p.work_counter, opw);
        base = __shfl_sync(FULL, base, 0);
        if (base >= p.n) break;
#if PRTC_WITH_TIMELINE
        if (p.timeline && lane == 0) {                            // diagnostic (PRTC_TIMELINE): when and where this claim started
            unsigned long long t; uint32_t sm;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
            p.timeline[3 * size_t(base / opw) + 0] = t;
            p.timeline[3 * size_t(base / opw) + 2] = sm;
        }
#endif
        if (CHAINED && p.chain_prev) {
            // the same characters' claim of the previous frame must have stored its results: poll, then one acquire load -- what
            // it orders (through the warp barrier, for all lanes) are the plain loads of the records and the slot arrays below
            if (lane == 0) {
                const uint32_t * done = p.chain_done + base / opw;
                while (*reinterpret_cast<const volatile uint32_t *>(done) != p.chain_prev) __nanosleep(64);
                uint32_t seen;
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(done) : "memory");
                (void)seen;
            }
            __syncwarp(FULL);
        }
        if (GATED) {
            // the batch is still arriving over PCIe: wait until this warp's records have landed (chunk boundaries are
            // multiples of 32 characters = whole 128-byte lines of both record arrays)
            const uint32_t need = base + 32u < p.n ? base + 32u : p.n;
            if (lane == 0) while (*reinterpret_cast<const volatile uint32_t *>(p.gate) < need) __nanosleep(200);
            __syncwarp(FULL);
            __threadfence();
        }
        const uint32_t i = base + lane;
        const bool valid = lane < opw && i < p.n;
        prtc_transform * tf = p.transforms + (valid ? i : base);
        prtc_char_physics * ph = p.physics + (valid ? i : base);

        // ---------------- LOAD + phase 1 (physics_system.cpp:277-285) ----------------
        V3 pos = v3(tf->position[0], tf->position[1], tf->position[2]);
        Quat q; q.x = tf->rotation[0]; q.y = tf->rotation[1]; q.z = tf->rotation[2]; q.w = tf->rotation[3];
        V3 vel = v3(ph->velocity[0], ph->velocity[1], ph->velocity[2]);
        V3 gn = v3(ph->ground_normal[0], ph->ground_normal[1], ph->ground_normal[2]);
        uint32_t cid = ph->collider;
        if (cid >= p.n_capsules) {                                // unknown capsule id: never dereferenced; reported, result undefined
            if (valid && p.counters) atomicAdd(p.counters + 6, 1ull);
            if (valid && p.host_flags) p.host_flags[1] = 1u;
            cid = 0;
        }
        const CapsuleDev cap = p.capsules[cid];
        const CapsuleFrame fr = capsule_frame(cap, rotation_of(q));
        const float radius = cap.radius;
        if (ph->grounded) vel = vel + ((-p.stick * p.gravity) * p.dt) * gn;
        float mvx = ph->movement[0], mvz = ph->movement[2];
        if (p.io_move && valid) {                                 // resident state: this frame's movement input comes from the I/O array
            const float * mv = p.io_move + 3 * size_t(i);
            mvx = mv[0]; mvz = mv[2];
            ph->movement[0] = mvx; ph->movement[1] = mv[1]; ph->movement[2] = mvz;
        }
        vel.x = vel.x + mvx;
        vel.z = vel.z + mvz;
        uint32_t grounded = 0;
        const V3 stepv = vel / fsub;
        V3 prevA, prevB;
        segment_ends(fr, pos, prevA, prevB);                      // prevTform                      :344
        uint32_t stepsLeft = p.substeps;
        sh.pose[PF_RADIUS][lane] = radius;
        sh.pose[PF_DMAX][lane] = (p.abs_mode == 1u) ? radius : floorf(radius) + 1.0f;   // largest plane distance surviving the range cull
        sh.pose[PF_EPS][lane] = p.qr_eps + 1e-5f * (fabsf(cap.radius) + fabsf(cap.height) + fabsf(cap.ox) + fabsf(cap.oy) + fabsf(cap.oz));
        // per-slot leaf list kept valid by k_xc_refresh: read per substep straight from device memory (coalesced: SoA)
        const size_t slot = size_t(p.xc_first) + i;
        bool cache_ok = valid && use_cache && p.xc_version[slot] == p.tree_version;
        Box region; region.lo = v3(0.0f); region.hi = v3(0.0f);
        uint32_t ncache = 0;
        if (cache_ok) {
            ncache = p.xc_count[slot];
            region.lo = v3(p.xc_region[0 * size_t(p.xc_stride) + slot], p.xc_region[1 * size_t(p.xc_stride) + slot], p.xc_region[2 * size_t(p.xc_stride) + slot]);
            region.hi = v3(p.xc_region[3 * size_t(p.xc_stride) + slot], p.xc_region[4 * size_t(p.xc_stride) + slot], p.xc_region[5 * size_t(p.xc_stride) + slot]);
        }
        if (CHAINED && use_cache) {
            // k_xc_refresh's job, done here: does the slot's region cover what this frame can need?  If not, one walk with the
            // need grown by the margin and stretched along the coasting motion (see k_xc_refresh) rewrites the slot.
            V3 eA, eB;
            segment_ends(fr, pos + vel, eA, eB);
            Box R = box_union(capsule_box(prevA, prevB, radius), capsule_box(eA, eB, radius));
            R.lo = R.lo - v3(0.25f); R.hi = R.hi + v3(0.25f);
            if (valid && !(cache_ok && box_contains(region, R))) {
                if (p.xc_reuse) {
                    const float r = fdiv(1.0f, 1.0f + p.dt * p.friction);
                    const float coast = r < 0.999f ? fdiv(r, 1.0f - r) : 0.0f;
                    const float dx = coast * ph->velocity[0] + 2.0f * mvx, dz = coast * ph->velocity[2] + 2.0f * mvz;
                    if (dx == dx && dz == dz && fabsf(dx) < 4.0f && fabsf(dz) < 4.0f) {
                        R.lo.x = R.lo.x + fminf(dx, 0.0f); R.hi.x = R.hi.x + fmaxf(dx, 0.0f);
                        R.lo.z = R.lo.z + fminf(dz, 0.0f); R.hi.z = R.hi.z + fmaxf(dz, 0.0f);
                    }
                    R.lo = R.lo - v3(p.xc_margin); R.hi = R.hi + v3(p.xc_margin);
                }
                uint32_t nd = 0, nl = 0;
                while (nd < p.n_nodes) {
                    const float4 lo = __ldg(p.nodes + 2 * nd);
                    const float4 hi = __ldg(p.nodes + 2 * nd + 1);
                    const bool ov = node_overlaps(lo, hi, R);
                    const int cnt = __float_as_int(hi.w);
                    if (cnt < 0) {
                        nd = ov ? nd + 1 : uint32_t(__float_as_int(lo.w));
                    } else {
                        if (ov) {
                            if (nl == uint32_t(XC_CAP)) { nl = 0xFFFFFFFFu; break; }
                            p.xc_leaf[size_t(nl) * p.xc_stride + slot] = nd;
                            ++nl;
                        }
                        nd = nd + 1;
                    }
                }
                if (nl == 0xFFFFFFFFu) { p.xc_version[slot] = 0u; cache_ok = false; }     // too many leaves for a slot: walk per substep
                else {
                    p.xc_region[0 * size_t(p.xc_stride) + slot] = R.lo.x; p.xc_region[1 * size_t(p.xc_stride) + slot] = R.lo.y;
                    p.xc_region[2 * size_t(p.xc_stride) + slot] = R.lo.z; p.xc_region[3 * size_t(p.xc_stride) + slot] = R.hi.x;
                    p.xc_region[4 * size_t(p.xc_stride) + slot] = R.hi.y; p.xc_region[5 * size_t(p.xc_stride) + slot] = R.hi.z;
                    p.xc_count[slot] = nl;
                    p.xc_version[slot] = p.tree_version;
                    region = R; ncache = nl; cache_ok = true;
                }
            }
            __syncwarp(FULL);
        }

        // JACOBI: the neighbours of the whole frame -- every other capsule whose stored box overlaps the region the frame can
        // reach (pose now, pose after the pre-stepped velocity, 0.25 of slack for push-outs) -- collected once, ascending
        bool dyn_cached = false;
        uint32_t n_dyn = 0;
        Box dyn_region; dyn_region.lo = v3(0.0f); dyn_region.hi = v3(0.0f);
        if (JACOBI && valid && !(p.tune & 1u)) {
            V3 eA, eB;
            segment_ends(fr, pos + vel, eA, eB);
            dyn_region = box_union(capsule_box(prevA, prevB, radius), capsule_box(eA, eB, radius));
            dyn_region.lo = dyn_region.lo - v3(0.25f); dyn_region.hi = dyn_region.hi + v3(0.25f);
            bool more = false;
            n_dyn = dyn_collect(dyn_grid_of(p), dyn_region, p.dyn_first + i, -1, &sh.dyn[0][lane], more);
            dyn_cached = !more;
        }
        bool running = valid;
        Box qb; qb.lo = v3(0.0f); qb.hi = v3(0.0f);
        uint32_t node = 0, nleaf = 0, remaining = 0, cur_k = 0, cur_off = 0;

        // ordered stackless walk (DynamicAABBTree::query order) from `node`, buffering up to PL_LEAF_CAP leaves
        auto ordered_walk = [&]() {
            nleaf = 0; remaining = 0; cur_k = 0; cur_off = 0;
            while (node < p.n_nodes && nleaf < uint32_t(PL_LEAF_CAP)) {
                const float4 lo = __ldg(p.nodes + 2 * node);
                const float4 hi = __ldg(p.nodes + 2 * node + 1);
                const bool ov = node_overlaps(lo, hi, qb);
                const int cnt = __float_as_int(hi.w);
                ++c_node;
                if (cnt < 0) {
                    node = ov ? node + 1 : uint32_t(__float_as_int(lo.w));
                } else {
                    if (ov) {
                        sh.lfirst[nleaf][lane] = uint32_t(__float_as_int(lo.w));
                        sh.lcount[nleaf][lane] = uint32_t(cnt);
                        remaining += uint32_t(cnt);
                        ++nleaf;
                    }
                    node = node + 1;
                }
            }
            c_tri += remaining;
        };
        auto publish_pose = [&](Pose const & ps) {
            st3(sh.pose, PF_A, lane, ps.A); st3(sh.pose, PF_B, lane, ps.B); st3(sh.pose, PF_CN, lane, ps.cn);
            st3(sh.pose, PF_PA, lane, ps.pointA); st3(sh.pose, PF_SV, lane, ps.segv);
            sh.pose[PF_SEGD][lane] = ps.segd;
        };

        // ---------------- phase 2: collideCharacterWithWorld (physics_system.cpp:330-438), substeps in lock step ----------------
        for (;;) {
            running = running && stepsLeft != 0;
            if (!__any_sync(FULL, running)) break;
            if (running) {
                --stepsLeft;
                pos = pos + stepv;                                //                                 :350
                Pose ps = make_pose(fr, pos);
                qb = box_union(capsule_box(prevA, prevB, radius), capsule_box(ps.A, ps.B, radius));   // :355-356
                prevA = ps.A; prevB = ps.B;                       // prevTform = tform               :358
                ++c_sub;
                if (JACOBI) {
                    // ---- capsule-capsule pass, CANONICAL order: every other capsule whose STORED fat
                    // box overlaps the query box, ascending global index, seen at its start-of-frame pose; stored boxes
                    // are binned in a uniform xz grid and a pair is reported from the one cell that holds the lower
                    // corner of the common cell range (same rule as k_step<JACOBI>)
                    // (the frame's neighbours were collected once, at LOAD, for a region covering the whole frame's motion; a
                    //  substep whose query box lies inside that region re-tests just those stored boxes -- same set, same
                    //  ascending order -- instead of searching the grid again)
                    const uint32_t self = p.dyn_first + i;
                    auto cc_test = [&](uint32_t j) {
                        const float4 sA = __ldg(p.dyn_records + 4 * size_t(j)), sB = __ldg(p.dyn_records + 4 * size_t(j) + 1);
                        V3 aA, aB;
                        segment_ends(fr, pos, aA, aB);
                        Contact ct;
                        ++c_cc;
                        if (capsule_capsule(aA, aB, radius, v3(sA.x, sA.y, sA.z), v3(sB.x, sB.y, sB.z), sA.w, ct)) {
                            pos = pos + ct.impulse;               // handleCollision (collision_system.cpp:242-252)
                            ++c_contact;
                            if (dot(ct.normal, v3(0.0f, 1.0f, 0.0f)) > p.ground_dot) { grounded = 1; gn = ct.normal; }
                        }
                    };
                    // (one loop for both cases, one copy of the capsule-capsule test: the instruction cache is precious here)
                    const bool from_cache = dyn_cached && box_contains(dyn_region, qb);
                    if (!from_cache) dyn_cached = false;          // the passes below reuse the buffer
                    long long last = -1;
                    for (;;) {
                        bool more = false;
                        const uint32_t ncand = from_cache ? n_dyn : dyn_collect(dyn_grid_of(p), qb, self, last, &sh.dyn[0][lane], more);
                        // the frame's neighbours against this substep's box, four stored boxes in flight at a time; the tests
                        // themselves stay in ascending order (every one sees the position the previous one left)
                        uint32_t live = from_cache ? 0u : ((ncand >= 32u) ? 0xFFFFFFFFu : ((1u << ncand) - 1u));
                        if (from_cache) {
                            for (uint32_t r4 = 0; r4 < ncand; r4 += 4) {
                                float4 blo[4], bhi[4];
#pragma unroll
                                for (int u = 0; u < 4; ++u) {
                                    const uint32_t j = sh.dyn[r4 + u < ncand ? r4 + u : r4][lane];
                                    blo[u] = __ldg(p.dyn_records + 4 * size_t(j) + 2); bhi[u] = __ldg(p.dyn_records + 4 * size_t(j) + 3);
                                }
#pragma unroll
                                for (int u = 0; u < 4; ++u) if (r4 + u < ncand && node_overlaps(blo[u], bhi[u], qb)) live |= 1u << (r4 + u);
                            }
                        }
                        while (live) {
                            const uint32_t r = uint32_t(__ffs(live) - 1);
                            live &= live - 1u;
                            cc_test(sh.dyn[r][lane]);
                        }
                        if (!more) break;
                        last = (long long)sh.dyn[ncand - 1][lane];
                    }
                    ps = make_pose(fr, pos);
                }
                publish_pose(ps);
                // ---- candidates: the slot's cached leaves that overlap the query box, or an ordered walk ----
                if (cache_ok && box_contains(region, qb)) {
                    nleaf = 0; remaining = 0; cur_k = 0; cur_off = 0;
                    bool fits = true;
                    // four leaves per trip: their indices, then their eight node words, are independent loads in flight
                    for (uint32_t j4 = 0; j4 < ncache && fits; j4 += 4) {
                        uint32_t leaf[4];
                        float4 lo[4], hi[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) leaf[u] = *(p.xc_leaf + size_t(j4 + u < ncache ? j4 + u : j4) * p.xc_stride + slot);
#pragma unroll
                        for (int u = 0; u < 4; ++u) { lo[u] = __ldg(p.nodes + 2 * leaf[u]); hi[u] = __ldg(p.nodes + 2 * leaf[u] + 1); }
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            if (j4 + u < ncache && fits) {
                                ++c_node;
                                if (node_overlaps(lo[u], hi[u], qb)) {
                                    if (nleaf == uint32_t(PL_LEAF_CAP)) fits = false;
                                    else {
                                        const uint32_t cnt = uint32_t(__float_as_int(hi[u].w));
                                        sh.lfirst[nleaf][lane] = uint32_t(__float_as_int(lo[u].w));
                                        sh.lcount[nleaf][lane] = cnt;
                                        remaining += cnt;
                                        ++nleaf;
                                    }
                                }
                            }
                        }
                    }
                    node = p.n_nodes;
                    if (fits) c_tri += remaining;
                    else { node = 0; ordered_walk(); }            // more candidates than one pass holds: walk in passes
                } else {
                    node = 0;
                    ordered_walk();
                }
                if (nleaf == 0) running = false;                  // `if (meshColIDs.empty()) break;`  :391-393
            }
            __syncwarp(FULL);

            // ---------------- rounds: pooled cull + exact test + ordered commit ----------------
            for (;;) {
                // more candidate leaves than one pass holds: next pass of the ordered walk
                if (running && remaining == 0 && node < p.n_nodes) ordered_walk();
                const bool has = running && remaining != 0;
                const unsigned m_has = __ballot_sync(FULL, has);
                if (!m_has) break;
                uint32_t w = uint32_t(PL_JOB_CAP) / uint32_t(__popc(m_has));
                w = w < wmax ? w : wmax;
                // the owner's window: its next triangles, at most `w`, at most PL_PIECES leaf pieces
                uint32_t n_o = 0;
                if (has) {
                    const uint32_t cap = remaining < w ? remaining : w;
                    uint32_t k = cur_k, of = cur_off;
                    int np = 0;
                    while (np < PL_PIECES && n_o < cap) {
                        const uint32_t cnt = sh.lcount[k][lane] - of;
                        if (cnt) {
                            sh.fst[np][lane] = sh.lfirst[k][lane] + of;
                            if (np) sh.bnd[np - 1][lane] = n_o;
                            n_o += cnt < cap - n_o ? cnt : cap - n_o;
                            ++np;
                        }
                        ++k; of = 0;
                    }
                    for (; np < PL_PIECES; ++np) if (np) sh.bnd[np - 1][lane] = 0xFFFFFFFFu;
                    // the segment this window is culled at; the pushes applied from here on are charged to `used`
                    st3(sh.pose, PF_A0, lane, ld3(sh.pose, PF_A, lane));
                    st3(sh.pose, PF_B0, lane, ld3(sh.pose, PF_B, lane));
                    sh.rk2o[__popc(m_has & lt)] = lane;
                }
                uint32_t incl = n_o;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(FULL, incl, d); if (int(lane) >= d) incl += t; }
                const uint32_t J = __shfl_sync(FULL, incl, 31);
                const uint32_t off_o = incl - n_o;
                sh.off[lane] = off_o;
                sh.mb_flags[lane] = 0u;
                sh.inval[lane] = 0u;
                float used = 0.0f;
                uint32_t cut = 0;
                __syncwarp(FULL);

                // Fetch the cull records of the trip that starts at job jf: the owner of job j is the one with the
                // (number of windows starting at or before j)-th window, counted with one redux over the owners' start bits;
                // the triangle follows from the owner's leaf pieces.
                uint32_t starts_before = 0;                       // windows starting before the trip being fetched
                auto fetch = [&](int stage, uint32_t jf, uint32_t & e, bool & ok) {
                    const uint32_t d = off_o - jf;
                    const unsigned M = __reduce_or_sync(FULL, (n_o != 0u && d < 32u) ? (1u << d) : 0u);
                    const uint32_t j = jf + lane;
                    ok = j < J;
                    if (ok) {
                        const uint32_t o = sh.rk2o[starts_before + __popc(M & (0xFFFFFFFFu >> (31u - lane))) - 1u];
                        const uint32_t q = j - sh.off[o];
                        const uint32_t b1 = sh.bnd[0][o], b2 = sh.bnd[1][o], b3 = sh.bnd[2][o];
                        const uint32_t piece = (q >= b1 ? 1u : 0u) + (q >= b2 ? 1u : 0u) + (q >= b3 ? 1u : 0u);
                        const uint32_t pbase = q >= b3 ? b3 : (q >= b2 ? b2 : (q >= b1 ? b1 : 0u));
                        e = (o << 27) | (sh.fst[piece][o] + (q - pbase));
                        const float4 * src = p.cull + 4 * size_t(e & PL_TRI_MASK);
                        cp_async16(&sh.ring[stage][0][lane], src);
                        cp_async16(&sh.ring[stage][1][lane], src + 1);
                        cp_async16(&sh.ring[stage][2][lane], src + 2);
                        cp_async16(&sh.ring[stage][3][lane], src + 3);
                    }
                    starts_before += __popc(M);
                    cp_async_commit();
                };
                // One exact trip: the first min(n2, 32) survivors of the ring, one per lane, against their owners' current
                // poses; ordered commit through the owners' mailboxes (see the file header)
                uint32_t n2 = 0, h2 = 0;
                auto exact_trip = [&]() {
                    const uint32_t take = n2 < 32u ? n2 : 32u;
                    uint32_t o = 32u + lane, qpos = 0, e = 0;
                    bool contact = false, ground = false, nonzero = false;
                    Contact ct; ct.impulse = v3(0.0f); ct.normal = v3(0.0f);
                    if (lane < take) {
                        e = sh.r2_e[(h2 + lane) & 63u];
                        qpos = sh.r2_q[(h2 + lane) & 63u];
                        o = e >> 27;
                        if (!sh.inval[o]) {
                            const float4 * src = p.tris + 3 * size_t(e & PL_TRI_MASK);
                            const float4 t0 = __ldg(src), t1 = __ldg(src + 1), t2 = __ldg(src + 2);
                            Pose ps;
                            ps.A = ld3(sh.pose, PF_A, o); ps.B = ld3(sh.pose, PF_B, o); ps.cn = ld3(sh.pose, PF_CN, o);
                            ps.pointA = ld3(sh.pose, PF_PA, o); ps.segv = ld3(sh.pose, PF_SV, o); ps.segd = sh.pose[PF_SEGD][o];
                            ++c_exact;
                            contact = capsule_triangle(ps, sh.pose[PF_RADIUS][o], p.abs_mode, v3(t0.x, t0.y, t0.z), v3(t1.x, t1.y, t1.z),
                                                       v3(t2.x, t2.y, t2.z), v3(t0.w, t1.w, t2.w), ct);
                            if (contact) {
                                ground = dot(ct.normal, v3(0.0f, 1.0f, 0.0f)) > p.ground_dot;
                                nonzero = ct.impulse.x != 0.0f || ct.impulse.y != 0.0f || ct.impulse.z != 0.0f;
                            }
                        } else {
                            o = 32u + lane;                       // void: its owner's window was cut earlier in this round
                        }
                    }
                    const unsigned seg = __match_any_sync(FULL, o);           // the lanes holding this owner's survivors (contiguous)
                    const unsigned m_contact = __ballot_sync(FULL, contact);
                    const unsigned m_push = __ballot_sync(FULL, nonzero);
                    const unsigned m_ground = __ballot_sync(FULL, ground);
                    bool retain = false;
                    if (o < 32u) {
                        // commit everything up to and including the owner's first pushing triangle; later lanes saw a stale pose
                        const unsigned push_seg = m_push & seg;
                        const int fp = push_seg ? __ffs(push_seg) - 1 : 31;
                        const unsigned upto = fp == 31 ? FULL : ((2u << fp) - 1u);
                        const unsigned v_contact = m_contact & seg & upto;
                        const unsigned v_ground = m_ground & seg & upto;
                        if (int(lane) == __ffs(seg) - 1 && v_contact) {
                            sh.mb_flags[o] = ((v_contact & ~m_push) ? 1u : 0u) | (v_ground ? 2u : 0u) | (push_seg ? 4u : 0u);
                            c_contact += __popc(v_contact);
                        }
                        if (push_seg && int(lane) == fp) { st3(sh.mb_imp, 0, o, ct.impulse); sh.mb_q[o] = qpos; }
                        if (v_ground && int(lane) == 31 - __clz(v_ground)) st3(sh.mb_gn, 0, o, ct.normal);   // the LAST ground contact in order wins (:247-252)
                        retain = push_seg != 0u && int(lane) > fp;
                    }
                    __syncwarp(FULL);
                    // the owners take their mailboxes at once: the next trip tests against the moved pose
                    {
                        const uint32_t flags = sh.mb_flags[lane];
                        if (flags) {
                            sh.mb_flags[lane] = 0u;
                            if (flags & 1u) pos = pos + v3(0.0f);         // `position += impulse` with a zero impulse (:242)
                            if (flags & 2u) { grounded = 1; gn = ld3(sh.mb_gn, 0, lane); }
                            if (flags & 4u) {
                                const V3 imp = ld3(sh.mb_imp, 0, lane);
                                pos = pos + imp;
                                publish_pose(make_pose(fr, pos));         // tform is rebuilt per triangle     :40
                                used += (fabsf(imp.x) + fabsf(imp.y)) + fabsf(imp.z);
                                if (!(used <= p.push_slack)) {            // budget spent (or NaN): cut the window behind this triangle
                                    sh.inval[lane] = 1u;
                                    cut = sh.mb_q[lane] + 1u;
                                }
                            }
                        }
                    }
                    __syncwarp(FULL);
                    // survivors that saw the stale pose stay at the head of the ring, order kept (unless their window was cut)
                    if (retain && sh.inval[o]) retain = false;
                    const unsigned m_ret = __ballot_sync(FULL, retain);
                    const uint32_t n_ret = uint32_t(__popc(m_ret));
                    if (retain) {
                        const uint32_t at = (h2 + take - n_ret + uint32_t(__popc(m_ret & lt))) & 63u;
                        sh.r2_e[at] = e;
                        sh.r2_q[at] = qpos;
                    }
                    h2 = (h2 + take - n_ret) & 63u;
                    n2 -= take - n_ret;
                    __syncwarp(FULL);
                };

                // ---- CULL: one job per lane and trip; the survivors go to the ring and are tested as soon as 32 of them wait.
                // (one call site for the exact trip and one cull body: unrolled copies of either thrash the instruction cache)
                uint32_t e_st[PL_STAGES];
                bool ok_st[PL_STAGES];
#pragma unroll
                for (int st = 0; st < PL_STAGES; ++st) fetch(st, uint32_t(st) * 32u, e_st[st], ok_st[st]);
                uint32_t j0 = 0;
                int ring_at = 0;
                const float slack = p.push_slack;
                for (;;) {
                    if (n2 >= 32u || (j0 >= J && n2 != 0u)) { exact_trip(); continue; }
                    if (j0 >= J) break;
                    cp_async_wait<PL_STAGES - 1>();
                    bool keep = false;
                    const uint32_t e = e_st[0];
                    const uint32_t o = e >> 27;
                    if (ok_st[0] && !sh.inval[o]) {
                        const float4 pl = sh.ring[ring_at][0][lane];
                        if (__float_as_uint(pl.x) != PRT_DEGENERATE_BITS) {               // length(n) == 0   :53
                            keep = true;
                            if (use_tq) {
                                const float4 c0 = sh.ring[ring_at][1][lane], c1 = sh.ring[ring_at][2][lane], c2 = sh.ring[ring_at][3][lane];
                                keep = !tight_reject_rec(ld3(sh.pose, PF_A0, o), ld3(sh.pose, PF_B0, o), sh.pose[PF_RADIUS][o], sh.pose[PF_DMAX][o],
                                                         sh.pose[PF_EPS][o], slack, v3(pl.x, pl.y, pl.z), pl.w, v3(c0.x, c0.y, c0.z), c0.w,
                                                         v3(c1.x, c1.y, c1.z), c1.w, v3(c2.x, c2.y, c2.z), c2.w);
                            }
                        }
                    }
                    const unsigned m = __ballot_sync(FULL, keep);
                    if (keep) {
                        const uint32_t at = (h2 + n2 + __popc(m & lt)) & 63u;
                        sh.r2_e[at] = e;
                        sh.r2_q[at] = (j0 + lane) - sh.off[o];
                        // the exact trip will want the triangle's vertices: start bringing its 48 bytes (one or two lines) to L1
                        const float4 * v = p.tris + 3 * size_t(e & PL_TRI_MASK);
                        asm volatile("prefetch.global.L1 [%0];" :: "l"(v));
                        asm volatile("prefetch.global.L1 [%0];" :: "l"(v + 2));
                    }
                    n2 += __popc(m);
#pragma unroll
                    for (int st = 0; st + 1 < PL_STAGES; ++st) { e_st[st] = e_st[st + 1]; ok_st[st] = ok_st[st + 1]; }
                    fetch(ring_at, j0 + 32u * PL_STAGES, e_st[PL_STAGES - 1], ok_st[PL_STAGES - 1]);
                    ring_at = ring_at + 1 == PL_STAGES ? 0 : ring_at + 1;
                    j0 += 32u;
                    __syncwarp(FULL);
                }
                cp_async_wait<0>();

                // ---- the owners advance their cursors: the whole window, or up to the cut ----
                if (has) {
                    uint32_t consumed = sh.inval[lane] ? cut : n_o;
                    remaining -= consumed;
                    while (consumed) {
                        const uint32_t left = sh.lcount[cur_k][lane] - cur_off;
                        if (consumed < left) { cur_off += consumed; consumed = 0; }
                        else { consumed -= left; ++cur_k; cur_off = 0; }
                    }
                }
                __syncwarp(FULL);
            }
        }

        // ---------------- remaining motion (:437), phase 3 (:298-317), STORE ----------------
        if (valid) {
            pos = pos + (float(stepsLeft) * vel) / fsub;
            V3 v = v3(ph->velocity[0], ph->velocity[1], ph->velocity[2]);   // prevVelocities[i]: still untouched in memory (:277,304)
            const float frictionRatio = fdiv(1.0f, 1.0f + (p.dt * p.friction));
            v.x = v.x * frictionRatio;
            v.z = v.z * frictionRatio;
            if (grounded) v.y = gmax((0.0f * p.gravity) * p.dt, v.y);
            else v.y = v.y + ((-1.0f * p.gravity) * p.dt);
            tf->position[0] = pos.x; tf->position[1] = pos.y; tf->position[2] = pos.z;
            ph->velocity[0] = v.x; ph->velocity[1] = v.y; ph->velocity[2] = v.z;
            ph->ground_normal[0] = gn.x; ph->ground_normal[1] = gn.y; ph->ground_normal[2] = gn.z;
            ph->grounded = grounded;
            if (p.io_result) p.io_result[i] = make_float4(pos.x, pos.y, pos.z, grounded ? 1.0f : 0.0f);
        }
        if (CHAINED) {
            // this claim's results are out: the next frame's claim of the same characters may start
            // (the lanes' stores happen before the warp barrier, the release store after it is cumulative)
            __syncwarp(FULL);
            if (lane == 0) asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(p.chain_done + base / opw), "r"(p.chain_serial) : "memory");
        }
        if (GATED) {
            // every finished character is counted (per 1024, then per 65536) so that the copy-out stream can start on a
            // chunk the moment its last character is written
            __threadfence();
            __syncwarp(FULL);
            if (lane == 0) {
                const uint32_t cnt = base + 32u <= p.n ? 32u : p.n - base;
                const uint32_t sb = base >> INGEST_SUB_SHIFT;
                const uint32_t left = p.n - (sb << INGEST_SUB_SHIFT);
                const uint32_t sb_size = left < (1u << INGEST_SUB_SHIFT) ? left : (1u << INGEST_SUB_SHIFT);
                if (atomicAdd(p.done_sub + sb, cnt) + cnt == sb_size) {
                    __threadfence_system();
                    atomicAdd(p.done_chunk + (sb >> INGEST_CHUNK_SHIFT), 1u);
                }
            }
        }
        __syncwarp(FULL);
#if PRTC_WITH_TIMELINE
        if (p.timeline && lane == 0) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            p.timeline[3 * size_t(base / opw) + 1] = t;
        }
#endif
    }

    c_sub = __reduce_add_sync(FULL, c_sub);
    c_tri = __reduce_add_sync(FULL, c_tri);
    c_node = __reduce_add_sync(FULL, c_node);
    c_contact = __reduce_add_sync(FULL, c_contact);
    c_exact = __reduce_add_sync(FULL, c_exact);
    c_cc = __reduce_add_sync(FULL, c_cc);
    if (p.done_flag && p.io_result) {
        // resident step: every result this warp wrote is visible to the host before the doorbell can ring
        __threadfence_system();
        __syncwarp(FULL);
        if (lane == 0 && atomicAdd(p.done_count, 1u) + 1u == gridDim.x) {
            *p.done_count = 0u;
            __threadfence_system();
            *reinterpret_cast<volatile uint32_t *>(p.done_flag) = p.done_value;
        }
    }
    if (lane == 0 && p.counters) {
        atomicAdd(p.counters + 0, (unsigned long long)c_sub);
        atomicAdd(p.counters + 1, (unsigned long long)c_tri);
        atomicAdd(p.counters + 2, (unsigned long long)c_node);
        atomicAdd(p.counters + 3, (unsigned long long)c_contact);
        if (JACOBI) atomicAdd(p.counters + 4, (unsigned long long)c_cc);
        atomicAdd(p.counters + 5, (unsigned long long)c_exact);
    }
}

} // namespace

// occupancy the launch bounds aim for (registers: 128 per thread, shared memory: ~10 KB per one-warp block)
uint32_t pool_blocks_per_sm() { return 16u; }

void launch_k_step_pool(StepParams const & p_in, uint32_t blocks, bool gated, int mode, cudaStream_t stream) {
    StepParams p = p_in;
    if (p.chain_done && !gated && mode == MODE_STATIC) {
        // a chained frame: its work counter (one of a ring of four) was zeroed by the launch before it or by the caller; the
        // attribute lets its blocks in before the previous launch has drained
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(blocks); cfg.blockDim = dim3(32); cfg.dynamicSmemBytes = 0; cfg.stream = stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = p.chain_prev ? 1 : 0;
        cudaLaunchKernelEx(&cfg, k_step_pool<false, MODE_STATIC, true>, p);
        return;
    }
    p.chain_done = nullptr;
    if (p.work_flip && p.work_pair && !gated) {
        const uint32_t k = *p.work_flip & 1u;                      // this launch's counter was zeroed by the previous one (or at creation)
        p.work_counter = p.work_pair + k;
        p.work_next = p.work_pair + (k ^ 1u);
        *p.work_flip = k ^ 1u;
    } else {
        p.work_next = nullptr;
        cudaMemsetAsync(p.work_counter, 0, sizeof(unsigned), stream);
    }
    if (mode == MODE_JACOBI) k_step_pool<false, MODE_JACOBI, false><<<blocks, 32, 0, stream>>>(p);
    else if (gated) k_step_pool<true, MODE_STATIC, false><<<blocks, 32, 0, stream>>>(p);
    else k_step_pool<false, MODE_STATIC, false><<<blocks, 32, 0, stream>>>(p);
}

} // namespace prt
