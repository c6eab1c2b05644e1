// ORACLE / TEST INFRASTRUCTURE ONLY.  Shadow of the gameplay character header
// (the real one drags in the animation system).  CharacterPhysics keeps the
// field order and defaults of reference character.h:19-27.  updateEquipment is
// the engine-side callback the reference invokes at every COLLIDE contact
// (collision_system.cpp:243); here it is the oracle's free per-contact trace
// hook.
#ifndef PRT_ORACLE_SHADOW_CHARACTER_H
#define PRT_ORACLE_SHADOW_CHARACTER_H
#include <functional>
#include <cstdint>
#include "src/game/scene/id.h"
#include "src/game/system/physics/collider_tag.h"
#include "src/game/component/component.h"
#include "src/container/vector.h"
#include "src/container/hash_map.h"
#include "src/container/hash_set.h"
#include <glm/glm.hpp>

class Scene;

struct CharacterPhysics {
    glm::vec3   forward = {1.0f, 0.0f, 0.0f};
    glm::vec3   velocity = {0.0f, 0.0f, 0.0f};
    float       mass = 1.0f;
    glm::vec3   movementVector = {0.0f, 0.0f, 0.0f};
    glm::vec3   groundNormal;
    ColliderTag colliderTag;
    bool        isGrounded = false;
};

// set by the oracle driver; called once per COLLIDE contact with the entity id
extern void (*prt_oracle_contact_hook)(EntityID);

struct CharacterAttributeInfo {
    void updateEquipment(EntityID id, Scene &) { if (prt_oracle_contact_hook) prt_oracle_contact_hook(id); }
};

struct Character {
    EntityID id = 0;
    CharacterAttributeInfo attributeInfo;
    CharacterPhysics physics;
};
#endif
