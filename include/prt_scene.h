/* prt_scene.h -- C-ABI of libprtscene.so: scene ingestion and the headless scene host around the collision path.
 *
 * Row f4 of the hot-path scope table: the data format and the callers either side of include/prtc.h.
 *   * the text `.prt` scene format        reference src/game/scene/scene_serialization.cpp:10-345 (loadScene)
 *   * OBJ geometry for MODEL colliders    stands in for Model::load (src/graphics/geometry/model.cpp:26-175, assimp)
 *   * the per-frame caller                Scene::updatePhysics / updateColliders (src/game/scene/scene.cpp:185-211) and
 *                                         CharacterSystem::updatePhysics (src/game/system/character/character_system.cpp:55-78)
 * so that a shipped level (res/scenes/cavern.prt) drives prtc_* end to end without the renderer, window or asset
 * pipeline.  Everything here is host code; the collision work itself happens behind include/prtc.h on the GPU.
 *
 * Plain C types only.  Every function returns a prts_status (0 = ok) unless stated; prts_last_error() describes the
 * last failure of the calling thread.
 */
#ifndef PRT_SCENE_H
#define PRT_SCENE_H

#include <stddef.h>
#include <stdint.h>

#include "prtc.h"

#ifdef __cplusplus
extern "C" {
#endif

enum prts_status {
    PRTS_OK = 0,
    PRTS_ERR_IO = 1,        /* file missing / unreadable, unsupported model format */
    PRTS_ERR_PARSE = 2,     /* where the reference parser asserts (scene_serialization.cpp:61-63,230-232,...) */
    PRTS_ERR_INVALID = 3,   /* bad argument */
    PRTS_ERR_LIMIT = 4,     /* more than PRTS_MAX_ENTITIES entities (Entities::N, entity.h:15) */
    PRTS_ERR_PHYSICS = 5    /* a prtc_* call failed; prts_last_error() carries prtc_last_error() */
};

#define PRTS_MAX_ENTITIES 100   /* src/game/scene/entity.h:15 */
#define PRTS_NAME_LEN 64        /* src/game/scene/entity.h:16 */
#define PRTS_PATH_LEN 128       /* scene_serialization.cpp:134,186-187 */

/* collider_tag.h:6-12 */
enum prts_collider_shape { PRTS_SHAPE_NONE = 0, PRTS_SHAPE_MODEL = 1, PRTS_SHAPE_MESH = 2, PRTS_SHAPE_CAPSULE = 3 };

/* One row of the reference's `Entities` table (entity.h:11-31) plus what the COLLIDER component carried. */
typedef struct prts_entity {
    char name[PRTS_NAME_LEN];        /* component 0; "entity_<i>" when absent (scene_serialization.cpp:13) */
    prtc_transform transform;        /* component 1; rotation normalised like parseQuat (:321-345) */
    int32_t model_id;                /* component 2: index into the scene's model table, -1 = none */
    int32_t animated;                /*              the <true|false> flag */
    int32_t animation_id;            /*              AnimationSystem::addAnimation order, -1 = none */
    int32_t character_id;            /* component 4: CharacterSystem::addCharacter order, -1 = none */
    uint32_t collider_shape;         /* component 3: enum prts_collider_shape */
    uint32_t collider_index;         /*              n-th MODEL collider / n-th CAPSULE collider (ColliderTag::index) */
    float capsule_height, capsule_radius, capsule_offset[3];
    int32_t light_index;             /* component 5: LightingSystem::addPointLight order, -1 = none */
    int32_t equipment_of;            /* entity created by a character's equipment list (:190): that character, else -1 */
} prts_entity;

typedef struct prts_equipment {      /* CharacterSystem::addEquipment (character_system.cpp:48-53) */
    int32_t character_id;
    int32_t entity;                  /* the equipment's own entity */
    int32_t model_id;
    char bone[PRTS_PATH_LEN];        /* bone NAME; resolving it needs the skinned model (out of scope) */
    prtc_transform offset;
} prts_equipment;

typedef struct prts_point_light {    /* lighting_system.h:19-25 */
    int32_t entity;
    float color[3];
    float constant, linear, quadratic;
} prts_point_light;

typedef struct prts_sun {            /* parseSun (:221-224): direction normalised */
    int32_t present;
    float direction[3];
    float color[3];
} prts_sun;

typedef struct prts_scene prts_scene;
typedef struct prts_model prts_model;
typedef struct prts_world prts_world;

const char * prts_last_error(void);

/* ---- .prt scene files: SceneSerialization::loadScene (scene_serialization.cpp:10-71) ------------------------------ */
int prts_scene_parse(const char * text, size_t len, prts_scene ** out);
int prts_scene_load(const char * path, prts_scene ** out);
void prts_scene_free(prts_scene * scene);

uint32_t prts_entity_count(const prts_scene * scene);
int prts_get_entity(const prts_scene * scene, uint32_t i, prts_entity * out);
uint32_t prts_model_count(const prts_scene * scene);           /* ModelManager::loadModel table, first-use order, de-duplicated by path */
int prts_get_model(const prts_scene * scene, uint32_t i, char path[PRTS_PATH_LEN], int32_t * animated);
uint32_t prts_character_count(const prts_scene * scene);
int prts_get_character(const prts_scene * scene, uint32_t i, int32_t * entity);
uint32_t prts_equipment_count(const prts_scene * scene);
int prts_get_equipment(const prts_scene * scene, uint32_t i, prts_equipment * out);
uint32_t prts_light_count(const prts_scene * scene);
int prts_get_light(const prts_scene * scene, uint32_t i, prts_point_light * out);
int prts_get_sun(const prts_scene * scene, prts_sun * out);
/* Canonical text rendering of everything above (floats as IEEE bit patterns); the parity tests compare it with the
 * same rendering of what the reference's own parser produced.  *needed receives the length incl. the final NUL. */
int prts_scene_dump(const prts_scene * scene, char * buf, size_t cap, size_t * needed);

/* ---- OBJ geometry ------------------------------------------------------------------------------------------------
 * What PhysicsSystem::addModelCollider reads from a Model (physics_system.cpp:60-85): positions, the index buffer and
 * one (startIndex, numIndices) per mesh.  Mesh = an `o`/`g` group (split again at `usemtl`, like assimp's
 * one-mesh-per-material); groups without faces are skipped (model.cpp:126); polygons are fan-triangulated; vertex and
 * face order are FILE order.  assimp's JoinIdenticalVertices / ImproveCacheLocality reordering is not reproduced. */
int prts_model_parse_obj(const char * text, size_t len, prts_model ** out);
int prts_model_load(const char * path, prts_model ** out);     /* by extension; only .obj is supported */
void prts_model_free(prts_model * model);
int prts_model_info(const prts_model * model, uint32_t * n_vertices, uint32_t * n_indices, uint32_t * n_meshes);
int prts_model_data(const prts_model * model, const float ** xyz, const uint32_t ** indices,
                    const uint32_t ** mesh_start_index, const uint32_t ** mesh_num_indices);

/* ---- headless scene host -------------------------------------------------------------------------------------------
 * Registers the scene's colliders in file order exactly as parseEntity does (scene_serialization.cpp:145-166: a MODEL
 * collider = every mesh of the entity's model under the entity's transform; a CAPSULE collider per character) on a
 * prtc context (cfg == NULL: prtc defaults with PRTC_ORDER_LITERAL, the engine-sized configuration), keeps the entity
 * transform table and one prtc_char_physics per character, and steps them.  model_dir is ModelManager's model
 * directory (model paths in the scene are relative to it); only models referenced by MODEL colliders are loaded. */
int prts_world_create(const prts_scene * scene, const char * model_dir, const prtc_config * cfg, prts_world ** out);
void prts_world_destroy(prts_world * world);
prtc_ctx * prts_world_ctx(prts_world * world);

uint32_t prts_world_entity_count(const prts_world * world);
uint32_t prts_world_character_count(const prts_world * world);
/* live tables owned by the world: entity transforms [entity_count]; character physics and entity ids [character_count].
 * The caller plays CharacterSystem::updateCharacters (input -> movement / jump velocity) by writing into physics. */
prtc_transform * prts_world_transforms(prts_world * world);
prtc_char_physics * prts_world_physics(prts_world * world);
const int32_t * prts_world_character_entities(const prts_world * world);
/* Scene::getTransform(id) = t for an entity with a MODEL collider: queues it like m_colliderUpdateSet (scene.cpp:190-211) */
int prts_world_move_entity(prts_world * world, uint32_t entity, const prtc_transform * t);
/* Scene::updatePhysics(dt) (scene.cpp:185-188): updateColliders -> prtc_update_models, then
 * CharacterSystem::updatePhysics (character_system.cpp:55-78): gather the characters' entity transforms, step, scatter */
int prts_world_update_physics(prts_world * world, float dt);
/* PhysicsSystem::raycast through the world's context (scene.cpp:243-252 issues four per frame for the camera) */
int prts_world_raycast(prts_world * world, uint32_t n, const float * origins, const float * directions,
                       const float * max_distances, float * hit_xyz, uint8_t * hit_mask);

#ifdef __cplusplus
}
#endif
#endif
