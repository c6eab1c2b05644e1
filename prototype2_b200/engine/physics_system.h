// Drop-in replacement for the engine's src/game/system/physics/physics_system.h.
//
// Same class name, same public member functions, same argument meaning as the reference's PhysicsSystem
// (reference src/game/system/physics/physics_system.h:24-77), so Scene, CharacterSystem, SceneSerialization and the
// editor compile against it unchanged.  Instead of owning a DynamicAABBTree and a CollisionSystem it owns a prtc_ctx
// (include/prtc.h) and forwards: the collision work runs in CUDA on a B200.  Built inside the engine tree (the
// includes below are the engine's own headers); see INTEGRATION.md.
#ifndef PHYSICS_SYSTEM_H
#define PHYSICS_SYSTEM_H

#include "src/game/component/component.h"
#include "src/game/system/physics/colliders.h"
#include "src/game/system/physics/collider_tag.h"
#include "src/graphics/geometry/model.h"
#include "src/game/system/character/character.h"

#include <glm/glm.hpp>

#include <cstddef>
#include <vector>

class Scene;
struct prtc_ctx;
struct prtc_multi;
struct prtc_transform;
struct prtc_char_physics;

class PhysicsSystem {
public:
    PhysicsSystem();
    ~PhysicsSystem();
    PhysicsSystem(PhysicsSystem const &) = delete;
    PhysicsSystem & operator=(PhysicsSystem const &) = delete;

    void newFrame();                                                           // physics_system.cpp:18-20

    void updateCapsuleCollider(ColliderTag const & tag, float height, float radius, glm::vec3 const & offset);   // :151-159
    void updateModelColliders(ColliderTag const * tags, Transform const * transforms, size_t count);            // :161-202
    bool raycast(glm::vec3 const & origin, glm::vec3 const & direction, float maxDistance, glm::vec3 & hit);      // :205-242
    void updateCharacters(float deltaTime, Scene & scene, Transform * transforms);                               // :245-318
    void updateTriggers(float deltaTime, ColliderTag const * triggers, Transform * transforms, size_t n);        // declared only, as in the reference

    ColliderTag addCapsuleCollider(float height, float radius, glm::vec3 const & offset);                        // :22-42
    ColliderTag addModelCollider(Model const & model, Transform const & transform);                              // :44-106
    void removeCollider(ColliderTag const & tag);                                                                // :108-120

    CapsuleCollider & getCapsuleCollider(ColliderTag tag);                                                       // physics_system.h:74
    float getGravity() const { return m_gravity; }

private:
    prtc_ctx * m_ctx = nullptr;
    prtc_multi * m_multi = nullptr;          // PRTC_DEVICES=...: several GPUs behind the same surface (include/prtc.h prtc_multi_*)
    std::vector<CapsuleCollider> m_capsules;      // what getCapsuleCollider hands out by reference ...
    std::vector<CapsuleCollider> m_uploaded;      // ... and what the device currently has (re-synced every frame)
    // the reference hands out model indices from a LIFO free list (physics_system.cpp:49-56,122-123): the tag index is the
    // slot, the slot holds the C-ABI's (never reused) model id
    std::vector<uint32_t> m_modelSlots;
    std::vector<uint16_t> m_modelFree;
    std::vector<prtc_transform> m_tf;
    std::vector<prtc_char_physics> m_ph;
    float m_gravity = 1.0f;                       // physics_system.h:122

    bool check(int status, char const * what) const;
    mutable int m_lastStatus = 0;

public:
    // Not in the reference (it has no failure to report): what happens when the device side fails -- no GPU, out of memory, a
    // lost device.  Default: the message goes to stderr and the process stops, as before.  With a handler installed the
    // failing call becomes a no-op instead (outputs untouched, raycast answers "no hit", tags come back as 0), the handler is
    // told which call failed with which prtc_status and message, and the engine can carry on without physics, fall back to the
    // reference's CPU PhysicsSystem or shut down in an orderly way.  Process-wide, so it can be installed before the first
    // PhysicsSystem is constructed.
    typedef void (*ErrorHandler)(int status, char const * what, char const * message, void * user);
    static void setErrorHandler(ErrorHandler handler, void * user);
    int lastStatus() const { return m_lastStatus; }      // prtc_status of the last failed call of this object (0: none failed)
};

#endif
