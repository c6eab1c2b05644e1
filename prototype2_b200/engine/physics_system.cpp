// Drop-in PhysicsSystem: the engine-facing C++ class over the C-ABI of include/prtc.h.  See physics_system.h.
#include "physics_system.h"

#include "prtc.h"

#include "src/game/scene/scene.h"

#include <cassert>
#include <cstdio>
#include <cstdlib>
#include <cstring>

static_assert(sizeof(Transform) == sizeof(prtc_transform), "prtc_transform must stay bit-compatible with the engine's Transform");

// Error behaviour of the reference: asserts (compiled out in its release build) and a bool from raycast.  A CUDA
// failure has no counterpart there: by default it is reported and the process stops instead of silently doing nothing;
// with PhysicsSystem::setErrorHandler the failing call is skipped and the engine is told (physics_system.h).
static PhysicsSystem::ErrorHandler g_errorHandler = nullptr;
static void * g_errorUser = nullptr;

void PhysicsSystem::setErrorHandler(ErrorHandler handler, void * user) {
    g_errorHandler = handler;
    g_errorUser = user;
}

bool PhysicsSystem::check(int status, char const * what) const {
    if (status == PRTC_OK) return true;
    m_lastStatus = status;
    if (g_errorHandler) {
        g_errorHandler(status, what, prtc_last_error(), g_errorUser);
        return false;
    }
    std::fprintf(stderr, "PhysicsSystem::%s: %s\n", what, prtc_last_error());
    assert(false && "prtc call failed");
    std::abort();
}

PhysicsSystem::PhysicsSystem() {
    prtc_config cfg;
    prtc_default_config(&cfg);
    cfg.gravity = m_gravity;
    // engine-sized scenes: keep the reference's exact candidate order and sequential character visibility
    cfg.order_mode = PRTC_ORDER_LITERAL;
    if (char const * e = std::getenv("PRTC_ABS_MODE")) cfg.abs_mode = unsigned(std::atoi(e));   // see prtc_abs_mode
    if (char const * e = std::getenv("PRTC_DEVICE")) cfg.device = std::atoi(e);
    if (char const * e = std::getenv("PRTC_ORDER")) cfg.order_mode = unsigned(std::atoi(e));     // see prtc_order_mode
    if (char const * e = std::getenv("PRTC_DYN")) cfg.dyn_mode = unsigned(std::atoi(e));         // see prtc_dyn_mode
    // PRTC_DEVICES=0,1,...: crowd-sized scenes over several GPUs of the node behind the same class surface (include/prtc.h
    // prtc_multi_*): geometry replicated, characters partitioned.  CANONICAL / BY_ID order only (LITERAL does not shard).
    if (char const * e = std::getenv("PRTC_DEVICES")) {
        std::vector<int32_t> devices;
        for (char const * p = e; *p;) {
            devices.push_back(int32_t(std::strtol(p, const_cast<char **>(&p), 10)));
            while (*p == ',' || *p == ' ') ++p;
        }
        if (cfg.order_mode == PRTC_ORDER_LITERAL) cfg.order_mode = PRTC_ORDER_CANONICAL;
        check(prtc_multi_create(&cfg, devices.data(), uint32_t(devices.size()), &m_multi), "PhysicsSystem");
        return;
    }
    check(prtc_create(&cfg, &m_ctx), "PhysicsSystem");
}

PhysicsSystem::~PhysicsSystem() {
    if (m_multi) prtc_multi_destroy(m_multi);
    else prtc_destroy(m_ctx);
}

// CollisionSystem::newFrame only rotates the collision-event sets, which nothing reads and which this
// implementation does not produce.
void PhysicsSystem::newFrame() {}

ColliderTag PhysicsSystem::addCapsuleCollider(float height, float radius, glm::vec3 const & offset) {
    assert(m_capsules.size() < 65535 && "Too many capsule colliders!");
    float const off[3] = {offset.x, offset.y, offset.z};
    uint32_t id = 0;
    check(m_multi ? prtc_multi_add_capsule(m_multi, height, radius, off, &id) : prtc_add_capsule(m_ctx, height, radius, off, &id), "addCapsuleCollider");
    CapsuleCollider c;
    c.height = height; c.radius = radius; c.offset = offset;
    m_capsules.push_back(c);
    m_uploaded.push_back(c);
    ColliderTag tag = { uint16_t(id), ColliderShape::COLLIDER_SHAPE_CAPSULE, ColliderType::COLLIDER_TYPE_COLLIDE };
    return tag;
}

ColliderTag PhysicsSystem::addModelCollider(Model const & model, Transform const & transform) {
    // what the reference copies out of the model (physics_system.cpp:68-91): positions through the index buffer,
    // mesh by mesh
    std::vector<float> xyz;
    std::vector<uint32_t> first, count;
    xyz.reserve(3 * model.indexBuffer.size());
    for (size_t m = 0; m < model.meshes.size(); ++m) {
        auto const & mesh = model.meshes[m];
        first.push_back(uint32_t(xyz.size() / 3));
        count.push_back(uint32_t(mesh.numIndices));
        for (size_t i = mesh.startIndex; i < mesh.startIndex + mesh.numIndices; ++i) {
            glm::vec3 const & p = model.vertexBuffer[model.indexBuffer[i]].pos;
            xyz.push_back(p.x); xyz.push_back(p.y); xyz.push_back(p.z);
        }
    }
    prtc_transform t;
    std::memcpy(&t, &transform, sizeof(t));
    uint32_t modelId = 0;
    check(m_multi ? prtc_multi_add_model(m_multi, xyz.data(), uint32_t(xyz.size() / 3), first.data(), count.data(), uint32_t(first.size()),
                         &t, &modelId, nullptr) : prtc_add_model(m_ctx, xyz.data(), uint32_t(xyz.size() / 3), first.data(), count.data(), uint32_t(first.size()),
                         &t, &modelId, nullptr), "addModelCollider");
    ColliderTag tag;
    if (!m_modelFree.empty()) {                                   // physics_system.cpp:49-56: the last freed index first
        tag.index = m_modelFree.back();
        m_modelFree.pop_back();
        m_modelSlots[tag.index] = modelId;
    } else {
        tag.index = uint16_t(m_modelSlots.size());
        m_modelSlots.push_back(modelId);
    }
    tag.shape = COLLIDER_SHAPE_MODEL;
    return tag;
}

void PhysicsSystem::removeCollider(ColliderTag const & tag) {
    switch (tag.shape) {
        case COLLIDER_SHAPE_MODEL:
            assert(tag.index < m_modelSlots.size());
            check(m_multi ? prtc_multi_remove_model(m_multi, m_modelSlots[tag.index]) : prtc_remove_model(m_ctx, m_modelSlots[tag.index]), "removeCollider");
            m_modelFree.push_back(tag.index);                     // physics_system.cpp:123
            break;
        case COLLIDER_SHAPE_CAPSULE:
            assert(false && "Not yet implemeneted!");            // as in the reference (physics_system.cpp:113-115)
            break;
        default:
            assert(false && "This collider shape can not be removed!");
    }
}

void PhysicsSystem::updateCapsuleCollider(ColliderTag const & tag, float height, float radius, glm::vec3 const & offset) {
    assert(tag.shape == COLLIDER_SHAPE_CAPSULE);
    m_capsules[tag.index].height = height;
    m_capsules[tag.index].radius = radius;
    m_capsules[tag.index].offset = offset;
}

CapsuleCollider & PhysicsSystem::getCapsuleCollider(ColliderTag tag) {
    assert(tag.shape == COLLIDER_SHAPE_CAPSULE);
    return m_capsules[tag.index];
}

void PhysicsSystem::updateModelColliders(ColliderTag const * tags, Transform const * transforms, size_t count) {
    std::vector<uint32_t> ids(count);
    std::vector<prtc_transform> tf(count);
    for (size_t i = 0; i < count; ++i) {
        assert(tags[i].shape == COLLIDER_SHAPE_MODEL);
        assert(tags[i].index < m_modelSlots.size());
        ids[i] = m_modelSlots[tags[i].index];
        std::memcpy(&tf[i], &transforms[i], sizeof(prtc_transform));
    }
    check(m_multi ? prtc_multi_update_models(m_multi, ids.data(), tf.data(), uint32_t(count)) : prtc_update_models(m_ctx, ids.data(), tf.data(), uint32_t(count)), "updateModelColliders");
}

bool PhysicsSystem::raycast(glm::vec3 const & origin, glm::vec3 const & direction, float maxDistance, glm::vec3 & hit) {
    float const o[3] = {origin.x, origin.y, origin.z}, d[3] = {direction.x, direction.y, direction.z};
    float h[3] = {0.0f, 0.0f, 0.0f};
    uint8_t mask = 0;
    check(m_multi ? prtc_multi_raycast(m_multi, 1, o, d, &maxDistance, h, &mask) : prtc_raycast(m_ctx, 1, o, d, &maxDistance, h, &mask), "raycast");
    if (mask) hit = glm::vec3(h[0], h[1], h[2]);
    return mask != 0;
}

void PhysicsSystem::updateCharacters(float deltaTime, Scene & scene, Transform * transforms) {
    CharacterSystem & characterSystem = scene.getCharacterSystem();
    size_t const n = characterSystem.getNumberOfCharacters();

    // capsules edited through getCapsuleCollider()'s reference or updateCapsuleCollider reach the device here
    for (size_t i = 0; i < m_capsules.size(); ++i) {
        CapsuleCollider const & a = m_capsules[i];
        CapsuleCollider & b = m_uploaded[i];
        if (a.height != b.height || a.radius != b.radius || a.offset.x != b.offset.x || a.offset.y != b.offset.y || a.offset.z != b.offset.z) {
            float const off[3] = {a.offset.x, a.offset.y, a.offset.z};
            check(m_multi ? prtc_multi_update_capsule(m_multi, uint32_t(i), a.height, a.radius, off) : prtc_update_capsule(m_ctx, uint32_t(i), a.height, a.radius, off), "updateCharacters");
            b = a;
        }
    }

    // gather (the engine keeps CharacterPhysics inside Character; the device wants flat arrays)
    m_tf.resize(n);
    m_ph.resize(n);
    for (size_t i = 0; i < n; ++i) {
        CharacterPhysics const & p = characterSystem.getCharacter(i).physics;
        std::memcpy(&m_tf[i], &transforms[i], sizeof(prtc_transform));
        prtc_char_physics & q = m_ph[i];
        q.velocity[0] = p.velocity.x; q.velocity[1] = p.velocity.y; q.velocity[2] = p.velocity.z;
        q.movement[0] = p.movementVector.x; q.movement[1] = p.movementVector.y; q.movement[2] = p.movementVector.z;
        q.ground_normal[0] = p.groundNormal.x; q.ground_normal[1] = p.groundNormal.y; q.ground_normal[2] = p.groundNormal.z;
        q.collider = p.colliderTag.index;
        q.grounded = p.isGrounded ? 1u : 0u;
    }

    if (!check(m_multi ? prtc_multi_step_characters(m_multi, deltaTime, uint32_t(n), m_tf.data(), m_ph.data()) : prtc_step_characters(m_ctx, deltaTime, uint32_t(n), m_tf.data(), m_ph.data()), "updateCharacters"))
        return;                                                   // (error handler installed: the frame is skipped, nothing is written)

    // scatter: position is the only Transform field the reference writes (collision_system.cpp:242, physics_system.cpp:350,437)
    for (size_t i = 0; i < n; ++i) {
        Character & character = characterSystem.getCharacter(i);
        CharacterPhysics & p = character.physics;
        transforms[i].position = glm::vec3(m_tf[i].position[0], m_tf[i].position[1], m_tf[i].position[2]);
        prtc_char_physics const & q = m_ph[i];
        p.velocity = glm::vec3(q.velocity[0], q.velocity[1], q.velocity[2]);
        p.groundNormal = glm::vec3(q.ground_normal[0], q.ground_normal[1], q.ground_normal[2]);
        p.isGrounded = q.grounded != 0;
        // the reference calls this engine-side hook at every contact (collision_system.cpp:243); it recomputes the
        // equipment transforms from the character transform and is called again by the caller
        // (character_system.cpp:74), so once per character after the step is equivalent
        character.attributeInfo.updateEquipment(character.id, scene);
    }
}
