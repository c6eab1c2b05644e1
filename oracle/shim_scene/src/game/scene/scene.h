// ORACLE / TEST INFRASTRUCTURE ONLY.  Shadow of the engine's Scene for ONE purpose: compiling the reference's
// src/game/scene/scene_serialization.cpp UNMODIFIED (oracle/Makefile, _ref/libprt_l0_scene.so) so that its parse of a
// `.prt` file can be recorded and compared with the product's parser (include/prt_scene.h).  The real header pulls in
// Vulkan, GLFW and assimp.  Every member below exists because the parser touches it
// (scene_serialization.cpp:12-18,115-224); calls into the engine's systems are recorded instead of executed.
#ifndef PRT_ORACLE_SHADOW_SCENE_SERIALIZATION_SCENE_H
#define PRT_ORACLE_SHADOW_SCENE_SERIALIZATION_SCENE_H

#include <cassert>
#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "src/game/component/component.h"     // the reference's own Transform / ColliderTag
#include "src/game/scene/id.h"
#include "src/container/vector.h"

enum LightType : uint16_t { LIGHT_TYPE_NONE, LIGHT_TYPE_POINT, TOTAL_NUM_LIGHT_TYPES };   // lighting_system.h:12-17
struct PointLight { glm::vec3 color; float constant; float linear; float quadratic; EntityID entityID; };
struct LightTag { uint16_t index; LightType type = LightType::LIGHT_TYPE_NONE; };

struct Entities {                              // entity.h:11-31, same sizes
    static constexpr size_t N = 100;
    static constexpr size_t SIZE_STR = 64;
    enum { maxSize = N };
    char        names[N][SIZE_STR];
    Transform   transforms[N];
    ModelID     modelIDs[N];
    AnimationID animationIDs[N];
    CharacterID characterIDs[N];
    ColliderTag colliderTags[N];
    LightTag    lightTags[N];
    EntityID addEntity() { ++nEntities; return nEntities - 1; }
    EntityID size() const { return nEntities; }
    EntityID nEntities = 0;
};

namespace shadow {
struct ModelRec { std::string path; bool animated; };
struct ModelColliderRec { int seq; int model; Transform t; };
struct CapsuleColliderRec { int seq; float h, r; glm::vec3 off; };
struct CharacterRec { EntityID entity; ColliderTag tag; };
struct EquipmentRec { CharacterID character; EntityID entity; std::string bone; Transform offset; };
struct AnimationRec { EntityID entity; };
struct Log {
    std::vector<ModelRec> models;
    std::vector<ModelColliderRec> modelColliders;
    std::vector<CapsuleColliderRec> capsuleColliders;
    std::vector<CharacterRec> characters;
    std::vector<EquipmentRec> equipment;
    std::vector<AnimationRec> animations;
    std::vector<PointLight> lights;
    std::string lastBone;
    int colliderSeq = 0;
};
}

class Model {                                  // what parseEntity asks of a model: its identity and a bone lookup
public:
    int id = -1;
    shadow::Log * log = nullptr;
    int getBoneIndex(char const * name) const { if (log) log->lastBone = name; return -1; }
};

class ModelManager {
public:
    shadow::Log * log = nullptr;
    std::vector<Model> table = std::vector<Model>(1);          // slot 0 answers models[-1] (a MODEL collider without a model component)
    ModelID loadModel(char const * path, bool animated) {      // model_manager.cpp:62-95: de-duplicated by path
        for (size_t i = 0; i < log->models.size(); ++i) if (log->models[i].path == path) return ModelID(i);
        log->models.push_back({path, animated});
        Model m; m.id = int(log->models.size()) - 1; m.log = log;
        table.push_back(m);
        return m.id;
    }
    void getModels(Model const *& models, size_t & nModels) { models = table.data() + 1; nModels = table.size() - 1; }
    Model const & getModel(ModelID id) { return table[size_t(id + 1)]; }
};
class AssetManager { public: ModelManager mm; ModelManager & getModelManager() { return mm; } };

class AnimationSystem {
public:
    shadow::Log * log = nullptr;
    AnimationID addAnimation(EntityID id) { log->animations.push_back({id}); return AnimationID(log->animations.size()) - 1; }
};

class PhysicsSystem {
public:
    shadow::Log * log = nullptr;
    ColliderTag addModelCollider(Model const & model, Transform const & t) {      // physics_system.cpp:44-106 (tag only)
        log->modelColliders.push_back({log->colliderSeq++, model.id, t});
        ColliderTag tag; tag.index = uint16_t(log->modelColliders.size() - 1);
        tag.shape = COLLIDER_SHAPE_MODEL; tag.type = COLLIDER_TYPE_COLLIDE; return tag;
    }
    ColliderTag addCapsuleCollider(float h, float r, glm::vec3 const & off) {     // physics_system.cpp:22-42 (tag only)
        log->capsuleColliders.push_back({log->colliderSeq++, h, r, off});
        ColliderTag tag; tag.index = uint16_t(log->capsuleColliders.size() - 1);
        tag.shape = COLLIDER_SHAPE_CAPSULE; tag.type = COLLIDER_TYPE_COLLIDE; return tag;
    }
};

class CharacterSystem {
public:
    shadow::Log * log = nullptr;
    CharacterID addCharacter(EntityID e, ColliderTag tag) {                        // character_system.cpp:37-46
        log->characters.push_back({e, tag});
        return CharacterID(log->characters.size()) - 1;
    }
    void addEquipment(CharacterID c, int /*boneIndex*/, EntityID e, Transform offset) {   // character_system.cpp:48-53
        log->equipment.push_back({c, e, log->lastBone, offset});
    }
};

struct Sun { glm::vec3 direction{0.0f, 0.0f, 0.0f}; glm::vec3 color{0.0f, 0.0f, 0.0f}; };
struct Lights { Sun sun; };

class Scene {
public:
    shadow::Log log;
    Entities m_entities;
    AssetManager m_assetManager;
    AnimationSystem m_animationSystem;
    PhysicsSystem m_physicsSystem;
    CharacterSystem m_characterSystem;
    Lights m_lights;
    Scene() {
        m_assetManager.mm.log = &log; m_assetManager.mm.table[0].log = &log; m_animationSystem.log = &log; m_physicsSystem.log = &log; m_characterSystem.log = &log;
    }
    Model const & getModel(EntityID id) { return m_assetManager.getModelManager().getModel(m_entities.modelIDs[id]); }   // scene.h:76
    void addPointLight(EntityID id, PointLight & pointLight) {                      // scene.cpp:222-225
        pointLight.entityID = id;
        log.lights.push_back(pointLight);
        m_entities.lightTags[id].index = uint16_t(log.lights.size() - 1);
        m_entities.lightTags[id].type = LIGHT_TYPE_POINT;
    }
};
#endif
