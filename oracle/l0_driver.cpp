// ORACLE / TEST INFRASTRUCTURE ONLY -- never linked into, imported by or
// executed from the product path (prototype2_b200/, include/).  Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// may load the library this file builds into.
//
// L0 = the LITERAL oracle: the reference's own collision hot path
//   /root/reference/src/game/system/physics/{physics_system,collision_system,aabb,aabb_tree}.cpp
//   /root/reference/src/util/physics_util.cpp
//   /root/reference/src/memory/{container_allocator,memory_util}.cpp
// compiled UNMODIFIED from where they lie (see oracle/Makefile) against the glm
// stand-in and shadow headers in oracle/shim/.  This file is only the C-ABI
// driver around the reference's public PhysicsSystem API
// (physics_system.h:24-77) so that Python tests can feed it scenes and read
// back states and traces.  It contains no collision arithmetic of its own.
//
// With -DL0_TRACED the same driver is linked against sed-instrumented copies
// of three reference files (generated at build time into oracle/_ref/gen/,
// never committed) that call the l0_trace_* hooks below; the tests assert the
// traced and the verbatim builds produce bit-identical states.
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <cmath>
#include <vector>
#include <limits>
#include <iostream>
#include <algorithm>
#include <functional>

// PhysicsSystem keeps its tree private; the driver needs read access to dump
// the DFS order through DynamicAABBTree::query (aabb_tree.h:26-29).  Layout is
// unaffected by access specifiers.  Standard headers are included above so the
// macro cannot touch them.
#define private public
#include "src/game/system/physics/physics_system.h"
#undef private
#include "src/game/scene/scene.h"

void (*prt_oracle_contact_hook)(EntityID) = nullptr;

namespace {

struct Trace {
    bool on = false;
    // one record per DynamicAABBTree::query issued by collideCharacterWithWorld
    std::vector<uint32_t> q_char, q_sub;
    std::vector<float>    q_aabb;              // 6 per query
    std::vector<uint32_t> q_mesh_off, q_caps_off;   // CSR offsets (size nq+1)
    std::vector<uint32_t> mesh_ids, caps_ids;
    std::vector<uint32_t> q_tri_tests, q_node_tests;
    // one record per handleCollision call
    std::vector<uint32_t> h_query;             // index into q_*
    std::vector<int32_t>  h_poly;              // aggregate polygon index, -1 = capsule-capsule
    std::vector<float>    h_impulse, h_normal, h_pos;  // 3 each
    uint32_t pending_node_tests = 0;
    void clear() {
        q_char.clear(); q_sub.clear(); q_aabb.clear();
        q_mesh_off.assign(1, 0); q_caps_off.assign(1, 0);
        mesh_ids.clear(); caps_ids.clear(); q_tri_tests.clear(); q_node_tests.clear();
        h_query.clear(); h_poly.clear(); h_impulse.clear(); h_normal.clear(); h_pos.clear();
        pending_node_tests = 0;
    }
};

struct World {
    PhysicsSystem ps;
    Scene scene;
    std::vector<Transform> transforms;
    Trace trace;
    uint64_t contacts = 0;
};

World * g_active = nullptr;   // the world whose l0_step is running (single-threaded, like the reference)

void contact_hook(EntityID id) {
    World * w = g_active;
    if (!w) return;
    ++w->contacts;
    Trace & t = w->trace;
    if (!t.on) return;
    // position right after `position += impulse` (collision_system.cpp:242)
    glm::vec3 const & p = w->transforms[size_t(id)].position;
    if (t.h_pos.size() / 3 < t.h_query.size()) {
        t.h_pos.push_back(p.x); t.h_pos.push_back(p.y); t.h_pos.push_back(p.z);
    }
}

} // namespace

// ---- hooks called from the instrumented copies (L0_TRACED build only) -------
extern "C" {
void l0_trace_node_test() { if (g_active && g_active->trace.on) ++g_active->trace.pending_node_tests; }
void l0_trace_query(uint32_t characterIndex, unsigned stepsLeft, const float * aabb6,
                    const uint16_t * mesh, size_t nMesh, const uint16_t * caps, size_t nCaps) {
    if (!g_active || !g_active->trace.on) return;
    Trace & t = g_active->trace;
    t.q_char.push_back(characterIndex);
    t.q_sub.push_back(3u - stepsLeft);
    for (int i = 0; i < 6; ++i) t.q_aabb.push_back(aabb6[i]);
    for (size_t i = 0; i < nMesh; ++i) t.mesh_ids.push_back(mesh[i]);
    for (size_t i = 0; i < nCaps; ++i) t.caps_ids.push_back(caps[i]);
    t.q_mesh_off.push_back(uint32_t(t.mesh_ids.size()));
    t.q_caps_off.push_back(uint32_t(t.caps_ids.size()));
    t.q_tri_tests.push_back(0);
    t.q_node_tests.push_back(t.pending_node_tests);
    t.pending_node_tests = 0;
}
void l0_trace_tri_test() {
    if (!g_active || !g_active->trace.on) return;
    Trace & t = g_active->trace;
    if (!t.q_tri_tests.empty()) ++t.q_tri_tests.back();
}
void l0_trace_hit(long poly, const float * impulse3, const float * normal3) {
    if (!g_active || !g_active->trace.on) return;
    Trace & t = g_active->trace;
    t.h_query.push_back(uint32_t(t.q_char.size() - 1));
    t.h_poly.push_back(int32_t(poly));
    for (int i = 0; i < 3; ++i) t.h_impulse.push_back(impulse3[i]);
    for (int i = 0; i < 3; ++i) t.h_normal.push_back(normal3[i]);
}
}

// ---- C-ABI ------------------------------------------------------------------
extern "C" {

void * l0_create() { return new World(); }
void l0_destroy(void * h) { delete static_cast<World *>(h); }

// xyz: 3*n_idx floats, index-expanded model-space vertices; mesh_first/mesh_count in
// vertices (multiples of 3); tform10 = position xyz, rotation wxyz, scale xyz.
int l0_add_model(void * h, const float * xyz, uint32_t n_idx,
                 const uint32_t * mesh_first, const uint32_t * mesh_count, uint32_t n_meshes,
                 const float * tform10) {
    World * w = static_cast<World *>(h);
    Model model;
    model.vertexBuffer.resize(n_idx);
    model.indexBuffer.resize(n_idx);
    for (uint32_t i = 0; i < n_idx; ++i) {
        model.vertexBuffer[i].pos = glm::vec3(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
        model.indexBuffer[i] = i;
    }
    model.meshes.resize(n_meshes);
    for (uint32_t m = 0; m < n_meshes; ++m) {
        model.meshes[m].startIndex = mesh_first[m];
        model.meshes[m].numIndices = mesh_count[m];
    }
    Transform t;
    t.position = glm::vec3(tform10[0], tform10[1], tform10[2]);
    t.rotation = glm::quat(tform10[3], tform10[4], tform10[5], tform10[6]);
    t.scale = glm::vec3(tform10[7], tform10[8], tform10[9]);
    ColliderTag tag = w->ps.addModelCollider(model, t);
    return int(tag.index);
}

// PhysicsSystem::removeCollider for a MODEL tag (physics_system.cpp:108-149).  Only well defined in the reference when the
// model is the most recently added one AND its first mesh index equals its model index (e.g. every model has one mesh):
// later meshes keep stale indices in the tree after the arrays are compacted (:141-146), and geometries[] is cleared at
// the MESH start index (:148, trap T19).  The tests stay inside that envelope.
void l0_remove_model(void * h, uint32_t model_index) {
    World * w = static_cast<World *>(h);
    ColliderTag tag;
    tag.index = ColliderIndex(model_index);
    tag.shape = COLLIDER_SHAPE_MODEL;
    tag.type = COLLIDER_TYPE_COLLIDE;
    w->ps.removeCollider(tag);
}

// adds the capsule collider AND the character that owns it (character i <-> capsule i)
int l0_add_capsule(void * h, float height, float radius, const float * off3) {
    World * w = static_cast<World *>(h);
    ColliderTag tag = w->ps.addCapsuleCollider(height, radius, glm::vec3(off3[0], off3[1], off3[2]));
    Character c;
    c.id = EntityID(w->scene.cs.chars.size());
    c.physics.colliderTag = tag;
    w->scene.cs.chars.push_back(c);
    w->transforms.push_back(Transform{});
    return int(tag.index);
}

// PhysicsSystem::updateModelColliders for one model collider (physics_system.cpp:161-202)
void l0_update_model(void * h, uint32_t model_index, const float * tform10) {
    World * w = static_cast<World *>(h);
    Transform t;
    t.position = glm::vec3(tform10[0], tform10[1], tform10[2]);
    t.rotation = glm::quat(tform10[3], tform10[4], tform10[5], tform10[6]);
    t.scale = glm::vec3(tform10[7], tform10[8], tform10[9]);
    ColliderTag tag = { uint16_t(model_index), ColliderShape::COLLIDER_SHAPE_MODEL, ColliderType::COLLIDER_TYPE_COLLIDE };
    w->ps.updateModelColliders(&tag, &t, 1);
}
// PhysicsSystem::updateCapsuleCollider (physics_system.cpp:151-159)
void l0_update_capsule(void * h, uint32_t capsule_index, float height, float radius, const float * off3) {
    World * w = static_cast<World *>(h);
    ColliderTag tag = { uint16_t(capsule_index), ColliderShape::COLLIDER_SHAPE_CAPSULE, ColliderType::COLLIDER_TYPE_COLLIDE };
    w->ps.updateCapsuleCollider(tag, height, radius, glm::vec3(off3[0], off3[1], off3[2]));
}

uint32_t l0_num_characters(void * h) { return uint32_t(static_cast<World *>(h)->transforms.size()); }

void l0_set_state(void * h, const float * pos3, const float * quat4_wxyz, const float * vel3,
                  const float * move3, const float * gnormal3, const uint8_t * grounded) {
    World * w = static_cast<World *>(h);
    size_t n = w->transforms.size();
    for (size_t i = 0; i < n; ++i) {
        CharacterPhysics & p = w->scene.cs.chars[i].physics;
        if (pos3) w->transforms[i].position = glm::vec3(pos3[3 * i], pos3[3 * i + 1], pos3[3 * i + 2]);
        if (quat4_wxyz) w->transforms[i].rotation = glm::quat(quat4_wxyz[4 * i], quat4_wxyz[4 * i + 1], quat4_wxyz[4 * i + 2], quat4_wxyz[4 * i + 3]);
        if (vel3) p.velocity = glm::vec3(vel3[3 * i], vel3[3 * i + 1], vel3[3 * i + 2]);
        if (move3) p.movementVector = glm::vec3(move3[3 * i], move3[3 * i + 1], move3[3 * i + 2]);
        if (gnormal3) p.groundNormal = glm::vec3(gnormal3[3 * i], gnormal3[3 * i + 1], gnormal3[3 * i + 2]);
        if (grounded) p.isGrounded = grounded[i] != 0;
    }
}

void l0_get_state(void * h, float * pos3, float * vel3, float * gnormal3, uint8_t * grounded) {
    World * w = static_cast<World *>(h);
    size_t n = w->transforms.size();
    for (size_t i = 0; i < n; ++i) {
        CharacterPhysics & p = w->scene.cs.chars[i].physics;
        if (pos3) { pos3[3 * i] = w->transforms[i].position.x; pos3[3 * i + 1] = w->transforms[i].position.y; pos3[3 * i + 2] = w->transforms[i].position.z; }
        if (vel3) { vel3[3 * i] = p.velocity.x; vel3[3 * i + 1] = p.velocity.y; vel3[3 * i + 2] = p.velocity.z; }
        if (gnormal3) { gnormal3[3 * i] = p.groundNormal.x; gnormal3[3 * i + 1] = p.groundNormal.y; gnormal3[3 * i + 2] = p.groundNormal.z; }
        if (grounded) grounded[i] = p.isGrounded ? 1 : 0;
    }
}

// one frame: PhysicsSystem::updateCharacters (physics_system.cpp:245-318).  newFrame() is never called.
void l0_step(void * h, float dt, int trace_on) {
    World * w = static_cast<World *>(h);
    w->trace.on = trace_on != 0;
    w->trace.clear();
    g_active = w;
    prt_oracle_contact_hook = contact_hook;
    w->ps.updateCharacters(dt, w->scene, w->transforms.data());
    g_active = nullptr;
}

uint64_t l0_contacts(void * h) { return static_cast<World *>(h)->contacts; }

// DFS (right-first preorder) order of every COLLIDE leaf, dumped through the tree's own
// query() with an all-enclosing box.  Returns mesh ids then capsule ids, each in DFS order.
void l0_dump_order(void * h, uint16_t * mesh_ids, uint32_t * n_mesh, uint16_t * caps_ids, uint32_t * n_caps) {
    World * w = static_cast<World *>(h);
    AABB huge;
    huge.lowerBound = glm::vec3(-std::numeric_limits<float>::max());
    huge.upperBound = glm::vec3(std::numeric_limits<float>::max());
    ColliderTag none;                      // shape NONE / type ERROR: never equals a stored tag
    none.index = 0xFFFF;
    prt::vector<uint16_t> meshes, caps;
    w->ps.m_aabbData.tree.query(none, huge, meshes, caps, COLLIDER_TYPE_COLLIDE);
    *n_mesh = uint32_t(meshes.size());
    *n_caps = uint32_t(caps.size());
    for (size_t i = 0; i < meshes.size(); ++i) if (mesh_ids) mesh_ids[i] = meshes[i];
    for (size_t i = 0; i < caps.size(); ++i) if (caps_ids) caps_ids[i] = caps[i];
}

// plain broadphase query through the reference tree (for candidate-set tests)
void l0_query(void * h, uint16_t caller_capsule, const float * aabb6,
              uint16_t * mesh_ids, uint32_t * n_mesh, uint16_t * caps_ids, uint32_t * n_caps, uint32_t cap) {
    World * w = static_cast<World *>(h);
    AABB box;
    box.lowerBound = glm::vec3(aabb6[0], aabb6[1], aabb6[2]);
    box.upperBound = glm::vec3(aabb6[3], aabb6[4], aabb6[5]);
    ColliderTag caller = { caller_capsule, COLLIDER_SHAPE_CAPSULE, COLLIDER_TYPE_COLLIDE };
    prt::vector<uint16_t> meshes, caps;
    w->ps.m_aabbData.tree.query(caller, box, meshes, caps, COLLIDER_TYPE_COLLIDE);
    *n_mesh = uint32_t(meshes.size());
    *n_caps = uint32_t(caps.size());
    for (size_t i = 0; i < meshes.size() && i < cap; ++i) mesh_ids[i] = meshes[i];
    for (size_t i = 0; i < caps.size() && i < cap; ++i) caps_ids[i] = caps[i];
}

// world-space triangle cache of a model, as the reference computed it (physics_system.cpp:86)
uint32_t l0_model_cache(void * h, uint32_t model, float * xyz, uint32_t cap_floats) {
    World * w = static_cast<World *>(h);
    auto & cache = w->ps.m_models.geometries[model].cache;
    uint32_t n = uint32_t(cache.size());
    for (uint32_t i = 0; i < n && 3 * i + 2 < cap_floats; ++i) { xyz[3 * i] = cache[i].x; xyz[3 * i + 1] = cache[i].y; xyz[3 * i + 2] = cache[i].z; }
    return n;
}

bool l0_raycast(void * h, const float * origin3, const float * dir3, float maxDistance, float * hit3) {
    World * w = static_cast<World *>(h);
    glm::vec3 hit(0.0f);
    bool r = w->ps.raycast(glm::vec3(origin3[0], origin3[1], origin3[2]), glm::vec3(dir3[0], dir3[1], dir3[2]), maxDistance, hit);
    hit3[0] = hit.x; hit3[1] = hit.y; hit3[2] = hit.z;
    return r;
}

// ---- trace read-back (valid after l0_step(..., trace_on=1) in the traced build) ----
uint32_t l0_trace_num_queries(void * h) { return uint32_t(static_cast<World *>(h)->trace.q_char.size()); }
uint32_t l0_trace_num_hits(void * h) { return uint32_t(static_cast<World *>(h)->trace.h_query.size()); }
uint32_t l0_trace_num_mesh_ids(void * h) { return uint32_t(static_cast<World *>(h)->trace.mesh_ids.size()); }
uint32_t l0_trace_num_caps_ids(void * h) { return uint32_t(static_cast<World *>(h)->trace.caps_ids.size()); }
void l0_trace_get(void * h, uint32_t * q_char, uint32_t * q_sub, float * q_aabb, uint32_t * q_mesh_off,
                  uint32_t * q_caps_off, uint32_t * mesh_ids, uint32_t * caps_ids, uint32_t * q_tri_tests,
                  uint32_t * q_node_tests, uint32_t * h_query, int32_t * h_poly, float * h_impulse,
                  float * h_normal, float * h_pos) {
    Trace & t = static_cast<World *>(h)->trace;
    auto cp = [](auto * dst, auto const & v) { if (dst && !v.empty()) std::memcpy(dst, v.data(), v.size() * sizeof(v[0])); };
    cp(q_char, t.q_char); cp(q_sub, t.q_sub); cp(q_aabb, t.q_aabb); cp(q_mesh_off, t.q_mesh_off);
    cp(q_caps_off, t.q_caps_off); cp(mesh_ids, t.mesh_ids); cp(caps_ids, t.caps_ids);
    cp(q_tri_tests, t.q_tri_tests); cp(q_node_tests, t.q_node_tests); cp(h_query, t.h_query);
    cp(h_poly, t.h_poly); cp(h_impulse, t.h_impulse); cp(h_normal, t.h_normal); cp(h_pos, t.h_pos);
}

#ifdef L0_TRACED
int l0_is_traced() { return 1; }
#else
int l0_is_traced() { return 0; }
#endif

} // extern "C"
