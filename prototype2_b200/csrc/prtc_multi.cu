// prtc_multi.cu -- several B200s of one node behind ONE handle (include/prtc.h prtc_multi_*): a context per device in one
// process, static geometry and its tree replicated, the characters of every call partitioned into contiguous slices
// (SURVEY.md 8e).  The static pass needs no exchange.  The dynamic-vs-dynamic pass (PRTC_DYN_JACOBI) has one per frame: every
// device's k_dyn_prep stores its characters' 64-byte records straight into all the other devices' record arrays (peer stores
// over NVLink, cudaDeviceEnablePeerAccess), events order "all records written" before any device bins them; without peer
// access the slices are copied with cudaMemcpyPeerAsync instead.
#include <memory>

#include "prtc_internal.h"

using namespace prt;
using namespace prt::host;

struct prtc_multi {
    std::vector<prtc_ctx *> ctx;
    std::vector<cudaEvent_t> ev_prep;            // "this device's records are written (everywhere)"
    bool peer = false;                           // every pair of devices can address each other's memory
    bool peers_wired = false;
    uint32_t dyn_n = 0;
    // partition of the current call / of the resident state
    std::vector<uint32_t> first, count;
    // resident state: one movement array and one result array for all devices (pinned, portable, mapped)
    unsigned char * io = nullptr;
    size_t io_bytes = 0;
    float * io_move = nullptr;
    float * io_result = nullptr;
    uint32_t resident_n = 0;
    bool resident = false;
};

namespace {

void partition(prtc_multi * m, uint32_t n) {
    const uint32_t g = uint32_t(m->ctx.size());
    uint32_t per = (n + g - 1) / g;
    per = (per + 31u) & ~31u;                    // whole 128-byte lines of both record arrays per slice
    m->first.assign(g, 0); m->count.assign(g, 0);
    uint32_t at = 0;
    for (uint32_t k = 0; k < g; ++k) {
        m->first[k] = at;
        m->count[k] = at < n ? std::min(per, n - at) : 0u;
        at += m->count[k];
    }
}

// JACOBI: every context owns the slice [first, first + count) of n characters and a record array for all n; with peer access
// each device's prep kernel writes into all of them
int wire_dynamic(prtc_multi * m, uint32_t n) {
    if (m->peers_wired && m->dyn_n == n) return PRTC_OK;
    const size_t g = m->ctx.size();
    for (size_t k = 0; k < g; ++k) {
        prtc_ctx * c = m->ctx[k];
        PRTC_CUDA(cudaSetDevice(c->cfg.device));
        int rc = prtc_dyn_configure(c, m->first[k], n);
        if (rc) return rc;
        rc = dyn_reserve(c, m->count[k]);
        if (rc) return rc;
    }
    for (size_t k = 0; k < g; ++k) {
        prtc_ctx * c = m->ctx[k];
        c->dyn_peers.n = 0;
        if (!m->peer) continue;
        for (size_t j = 0; j < g; ++j)
            if (j != k) c->dyn_peers.records[c->dyn_peers.n++] = m->ctx[j]->d_dyn_records.ptr;
    }
    m->peers_wired = true;
    m->dyn_n = n;
    return PRTC_OK;
}

// after every device has enqueued its prep: nobody bins the records before all of them are there
int exchange_records(prtc_multi * m) {
    const size_t g = m->ctx.size();
    for (size_t k = 0; k < g; ++k) {
        PRTC_CUDA(cudaSetDevice(m->ctx[k]->cfg.device));
        PRTC_CUDA(cudaEventRecord(m->ev_prep[k], m->ctx[k]->stream));
    }
    for (size_t k = 0; k < g; ++k) {
        prtc_ctx * c = m->ctx[k];
        PRTC_CUDA(cudaSetDevice(c->cfg.device));
        for (size_t j = 0; j < g; ++j) {
            if (j == k) continue;
            PRTC_CUDA(cudaStreamWaitEvent(c->stream, m->ev_prep[j], 0));
            if (!m->peer && m->count[j]) {       // no peer stores: fetch device j's slice
                const size_t off = 4 * size_t(m->first[j]);
                PRTC_CUDA(cudaMemcpyPeerAsync(c->d_dyn_records.ptr + off, c->cfg.device, m->ctx[j]->d_dyn_records.ptr + off, m->ctx[j]->cfg.device,
                                              4 * size_t(m->count[j]) * sizeof(float4), c->stream));
            }
        }
    }
    return PRTC_OK;
}

bool jacobi(prtc_multi * m) { return m->ctx[0]->jacobi(); }

} // namespace

extern "C" {

int prtc_multi_create(const prtc_config * cfg, const int32_t * devices, uint32_t n_devices, prtc_multi ** out) {
    if (!cfg || !devices || !out || n_devices == 0 || n_devices > 16) return fail(PRTC_ERR_INVALID, "prtc_multi_create: null argument or 0 / more than 16 devices");
    if (cfg->order_mode == PRTC_ORDER_LITERAL)
        return fail(PRTC_ERR_INVALID, "prtc_multi_create: PRTC_ORDER_LITERAL keeps the reference's sequential visibility between characters and does not shard");
    std::unique_ptr<prtc_multi> m(new (std::nothrow) prtc_multi);
    if (!m) return fail(PRTC_ERR_ALLOC, "prtc_multi_create: out of memory");
    auto abandon = [&](int rc) { for (prtc_ctx * c : m->ctx) prtc_destroy(c); for (cudaEvent_t e : m->ev_prep) cudaEventDestroy(e); return rc; };
    for (uint32_t k = 0; k < n_devices; ++k) {
        prtc_config c = *cfg;
        c.device = devices[k];
        prtc_ctx * ctx = nullptr;
        const int rc = prtc_create(&c, &ctx);
        if (rc) return abandon(rc);
        m->ctx.push_back(ctx);
        cudaEvent_t e = nullptr;
        if (cudaSetDevice(devices[k]) != cudaSuccess || cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess)
            return abandon(fail(PRTC_ERR_CUDA, "prtc_multi_create: cudaEventCreate failed"));
        m->ev_prep.push_back(e);
    }
    // peer access between every pair of distinct devices (two contexts on one device address the same memory anyway)
    m->peer = true;
    for (uint32_t a = 0; a < n_devices && m->peer; ++a)
        for (uint32_t b = 0; b < n_devices; ++b) {
            if (devices[a] == devices[b]) continue;
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, devices[a], devices[b]) != cudaSuccess || !can) { m->peer = false; break; }
            cudaSetDevice(devices[a]);
            const cudaError_t e = cudaDeviceEnablePeerAccess(devices[b], 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { m->peer = false; break; }
            (void)cudaGetLastError();
        }
    *out = m.release();
    return PRTC_OK;
}

void prtc_multi_destroy(prtc_multi * m) {
    if (!m) return;
    for (prtc_ctx * c : m->ctx) prtc_destroy(c);
    for (cudaEvent_t e : m->ev_prep) cudaEventDestroy(e);
    if (m->io) cudaFreeHost(m->io);
    delete m;
}

uint32_t prtc_multi_device_count(const prtc_multi * m) { return m ? uint32_t(m->ctx.size()) : 0u; }
prtc_ctx * prtc_multi_context(prtc_multi * m, uint32_t k) { return (m && k < m->ctx.size()) ? m->ctx[k] : nullptr; }

// ---- registration: the same calls on every device, the ids come out the same ----
int prtc_multi_add_model(prtc_multi * m, const float * xyz, uint32_t n_vertices, const uint32_t * mesh_first, const uint32_t * mesh_count,
                         uint32_t n_meshes, const prtc_transform * transform, uint32_t * out_model_id, uint32_t * out_first_mesh_id) {
    if (!m) return fail(PRTC_ERR_INVALID, "prtc_multi_add_model: null handle");
    for (prtc_ctx * c : m->ctx) {
        const int rc = prtc_add_model(c, xyz, n_vertices, mesh_first, mesh_count, n_meshes, transform, out_model_id, out_first_mesh_id);
        if (rc) return rc;
    }
    return PRTC_OK;
}
int prtc_multi_add_capsule(prtc_multi * m, float height, float radius, const float offset[3], uint32_t * out_id) {
    if (!m) return fail(PRTC_ERR_INVALID, "prtc_multi_add_capsule: null handle");
    for (prtc_ctx * c : m->ctx) { const int rc = prtc_add_capsule(c, height, radius, offset, out_id); if (rc) return rc; }
    return PRTC_OK;
}
int prtc_multi_update_capsule(prtc_multi * m, uint32_t id, float height, float radius, const float offset[3]) {
    if (!m) return fail(PRTC_ERR_INVALID, "prtc_multi_update_capsule: null handle");
    for (prtc_ctx * c : m->ctx) { const int rc = prtc_update_capsule(c, id, height, radius, offset); if (rc) return rc; }
    return PRTC_OK;
}
int prtc_multi_remove_model(prtc_multi * m, uint32_t model_id) {
    if (!m) return fail(PRTC_ERR_INVALID, "prtc_multi_remove_model: null handle");
    for (prtc_ctx * c : m->ctx) { const int rc = prtc_remove_model(c, model_id); if (rc) return rc; }
    return PRTC_OK;
}
int prtc_multi_update_models(prtc_multi * m, const uint32_t * model_ids, const prtc_transform * transforms, uint32_t n) {
    if (!m) return fail(PRTC_ERR_INVALID, "prtc_multi_update_models: null handle");
    for (prtc_ctx * c : m->ctx) { const int rc = prtc_update_models(c, model_ids, transforms, n); if (rc) return rc; }
    return PRTC_OK;
}
int prtc_multi_commit(prtc_multi * m) {
    if (!m) return fail(PRTC_ERR_INVALID, "prtc_multi_commit: null handle");
    for (prtc_ctx * c : m->ctx) { const int rc = prtc_commit(c); if (rc) return rc; }
    return PRTC_OK;
}
int prtc_multi_set_option(prtc_multi * m, int option, int value) {
    if (!m) return fail(PRTC_ERR_INVALID, "prtc_multi_set_option: null handle");
    for (prtc_ctx * c : m->ctx) { const int rc = prtc_set_option(c, option, value); if (rc) return rc; }
    return PRTC_OK;
}
int prtc_multi_raycast(prtc_multi * m, uint32_t n, const float * origins, const float * directions, const float * max_distances, float * hit_xyz,
                       uint8_t * hit_mask) {
    if (!m) return fail(PRTC_ERR_INVALID, "prtc_multi_raycast: null handle");
    return prtc_raycast(m->ctx[0], n, origins, directions, max_distances, hit_xyz, hit_mask);      // geometry is replicated: any device answers
}
int prtc_multi_get_counters(prtc_multi * m, prtc_counters * out, int reset) {
    if (!m || !out) return fail(PRTC_ERR_INVALID, "prtc_multi_get_counters: null argument");
    std::memset(out, 0, sizeof(*out));
    for (prtc_ctx * c : m->ctx) {
        prtc_counters k;
        const int rc = prtc_get_counters(c, &k, reset);
        if (rc) return rc;
        out->substeps += k.substeps; out->tri_tests += k.tri_tests; out->node_tests += k.node_tests; out->contacts += k.contacts;
        out->cc_tests += k.cc_tests; out->exact_tests += k.exact_tests; out->launches += k.launches;
    }
    return PRTC_OK;
}

// ---- PhysicsSystem::updateCharacters over all devices, HOST buffers ----
// Slice k of the batch goes to device k: copy in, step, copy out, every device on its own stream and PCIe link; the host
// thread only enqueues and then waits for all of them.
int prtc_multi_step_characters(prtc_multi * m, float dt, uint32_t n, prtc_transform * tf, prtc_char_physics * ph) {
    if (!m || (n && (!tf || !ph))) return fail(PRTC_ERR_INVALID, "prtc_multi_step_characters: null argument");
    if (m->resident) return fail(PRTC_ERR_INVALID, "prtc_multi_step_characters: resident state is active (prtc_multi_resident_step)");
    partition(m, n);
    const size_t g = m->ctx.size();
    const bool dyn = jacobi(m);
    if (dyn) { const int rc = wire_dynamic(m, n); if (rc) return rc; }
    for (size_t k = 0; k < g; ++k) {
        prtc_ctx * c = m->ctx[k];
        const uint32_t cnt = m->count[k], f = m->first[k];
        PRTC_CUDA(cudaSetDevice(c->cfg.device));
        if (c->geometry_dirty || c->tree_dirty || c->capsules_dirty) { const int rc = commit_impl(c); if (rc) return rc; }
        PRTC_CUDA(c->d_tf.reserve(cnt));
        PRTC_CUDA(c->d_ph.reserve(cnt));
        if (cnt) {
            PRTC_CUDA(cudaMemcpyAsync(c->d_tf.ptr, tf + f, size_t(cnt) * sizeof(prtc_transform), cudaMemcpyHostToDevice, c->stream));
            PRTC_CUDA(cudaMemcpyAsync(c->d_ph.ptr, ph + f, size_t(cnt) * sizeof(prtc_char_physics), cudaMemcpyHostToDevice, c->stream));
        }
        if (dyn) { const int rc = dyn_prepare(c, dt, cnt, c->d_tf.ptr, c->d_ph.ptr); if (rc) return rc; }
    }
    if (dyn) { const int rc = exchange_records(m); if (rc) return rc; }
    for (size_t k = 0; k < g; ++k) {
        prtc_ctx * c = m->ctx[k];
        const uint32_t cnt = m->count[k], f = m->first[k];
        PRTC_CUDA(cudaSetDevice(c->cfg.device));
        if (cnt || dyn) { const int rc = prtc_step_characters_device(c, dt, cnt, c->d_tf.ptr, c->d_ph.ptr); if (rc) return rc; }
        if (cnt) {
            PRTC_CUDA(cudaMemcpyAsync(tf + f, c->d_tf.ptr, size_t(cnt) * sizeof(prtc_transform), cudaMemcpyDeviceToHost, c->stream));
            PRTC_CUDA(cudaMemcpyAsync(ph + f, c->d_ph.ptr, size_t(cnt) * sizeof(prtc_char_physics), cudaMemcpyDeviceToHost, c->stream));
        }
    }
    int status = PRTC_OK;
    for (size_t k = 0; k < g; ++k) {
        PRTC_CUDA(cudaSetDevice(m->ctx[k]->cfg.device));
        const int rc = check_bad_colliders(m->ctx[k]);            // synchronises device k
        if (rc && !status) status = rc;
    }
    return status;
}

// ---- resident state over all devices: ONE movement array and ONE result array, slice k read / written by device k ----
int prtc_multi_resident_begin(prtc_multi * m, uint32_t n, const prtc_transform * tf, const prtc_char_physics * ph) {
    if (!m || (n && (!tf || !ph))) return fail(PRTC_ERR_INVALID, "prtc_multi_resident_begin: null argument");
    partition(m, n);
    const size_t g = m->ctx.size();
    const size_t need = 64 * g + 64 + size_t(n) * 7 * sizeof(float) + 64;
    if (need > m->io_bytes) {
        if (m->io) cudaFreeHost(m->io);
        m->io = nullptr; m->io_bytes = 0;
        PRTC_CUDA(cudaHostAlloc(reinterpret_cast<void **>(&m->io), need, cudaHostAllocMapped | cudaHostAllocPortable));
        m->io_bytes = need;
    }
    m->io_move = reinterpret_cast<float *>(m->io + 64 * g);
    m->io_result = m->io_move + 3 * size_t(n) + ((4 - (3 * size_t(n)) % 4) % 4);
    if (jacobi(m)) { const int rc = wire_dynamic(m, n); if (rc) return rc; }
    for (size_t k = 0; k < g; ++k) {
        const uint32_t f = m->first[k], cnt = m->count[k];
        const int rc = resident_begin_impl(m->ctx[k], cnt, cnt ? tf + f : tf, cnt ? ph + f : ph, reinterpret_cast<uint32_t *>(m->io + 64 * k),
                                           m->io_move + 3 * size_t(f), m->io_result + 4 * size_t(f));
        if (rc) return rc;
    }
    m->resident = true;
    m->resident_n = n;
    return PRTC_OK;
}
int prtc_multi_resident_io(prtc_multi * m, float ** movement, float ** result) {
    if (!m || !m->resident) return fail(PRTC_ERR_INVALID, "prtc_multi_resident_io: call prtc_multi_resident_begin first");
    if (movement) *movement = m->io_move;
    if (result) *result = m->io_result;
    return PRTC_OK;
}
int prtc_multi_resident_step(prtc_multi * m, float dt) {
    if (!m || !m->resident) return fail(PRTC_ERR_INVALID, "prtc_multi_resident_step: call prtc_multi_resident_begin first");
    const size_t g = m->ctx.size();
    if (jacobi(m)) {
        for (size_t k = 0; k < g; ++k) {
            prtc_ctx * c = m->ctx[k];
            PRTC_CUDA(cudaSetDevice(c->cfg.device));
            const int rc = dyn_prepare(c, dt, c->resident_n, c->d_tf.ptr, c->d_ph.ptr);
            if (rc) return rc;
        }
        const int rc = exchange_records(m);
        if (rc) return rc;
    }
    for (size_t k = 0; k < g; ++k) { const int rc = resident_launch(m->ctx[k], dt); if (rc) return rc; }
    int status = PRTC_OK;
    for (size_t k = 0; k < g; ++k) { const int rc = resident_wait(m->ctx[k]); if (rc && !status) status = rc; }
    return status;
}
int prtc_multi_resident_fetch(prtc_multi * m, prtc_transform * tf, prtc_char_physics * ph) {
    if (!m || !m->resident) return fail(PRTC_ERR_INVALID, "prtc_multi_resident_fetch: no resident state");
    for (size_t k = 0; k < m->ctx.size(); ++k) {
        const int rc = prtc_resident_fetch(m->ctx[k], tf ? tf + m->first[k] : nullptr, ph ? ph + m->first[k] : nullptr);
        if (rc) return rc;
    }
    return PRTC_OK;
}
int prtc_multi_resident_end(prtc_multi * m) {
    if (!m) return fail(PRTC_ERR_INVALID, "prtc_multi_resident_end: null handle");
    for (prtc_ctx * c : m->ctx) prtc_resident_end(c);
    m->resident = false;
    m->resident_n = 0;
    return PRTC_OK;
}

} // extern "C"
