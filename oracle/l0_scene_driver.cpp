// ORACLE / TEST INFRASTRUCTURE ONLY -- never linked into, imported by or executed from the product path.
//
// C-ABI driver around the reference's own `.prt` parser: /root/reference/src/game/scene/scene_serialization.cpp is
// compiled UNMODIFIED (oracle/Makefile -> _ref/libprt_l0_scene.so) against the recording Scene in
// oracle/shim_scene/.  l0s_dump() runs SceneSerialization::loadScene on a file and renders what it did as canonical
// text; tests compare that text with prts_scene_dump() of the product parser (include/prt_scene.h) on the same file.
// This file contains no parsing of its own.
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>

#include "src/game/scene/scene_serialization.h"
#include "src/game/scene/scene.h"

namespace {
uint32_t bits(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
void put(std::string & s, char const * fmt, ...) __attribute__((format(printf, 2, 3)));
void put(std::string & s, char const * fmt, ...) {
    char line[512];
    va_list ap; va_start(ap, fmt); std::vsnprintf(line, sizeof line, fmt, ap); va_end(ap);
    s += line;
}
void put3(std::string & s, glm::vec3 const & v) { put(s, " %08x %08x %08x", bits(v.x), bits(v.y), bits(v.z)); }
void putT(std::string & s, Transform const & t) {
    put(s, " pos"); put3(s, t.position);
    put(s, " rot %08x %08x %08x %08x", bits(t.rotation.x), bits(t.rotation.y), bits(t.rotation.z), bits(t.rotation.w));
    put(s, " scale"); put3(s, t.scale);
}
}

extern "C" int l0s_dump(char const * path, char * buf, size_t cap, size_t * needed) {
    Scene * scene = new Scene();
    SceneSerialization::loadScene(path, *scene);
    shadow::Log const & log = scene->log;
    std::string s;
    put(s, "sun dir"); put3(s, scene->m_lights.sun.direction); put(s, " color"); put3(s, scene->m_lights.sun.color); put(s, "\n");
    for (size_t i = 0; i < log.models.size(); ++i)
        put(s, "model %zu \"%s\" animated %d\n", i, log.models[i].path.c_str(), int(log.models[i].animated));
    for (EntityID i = 0; i < scene->m_entities.size(); ++i) {
        Entities const & e = scene->m_entities;
        put(s, "entity %d name \"%s\"", i, e.names[i]); putT(s, e.transforms[i]);
        bool const has = e.colliderTags[i].shape != COLLIDER_SHAPE_NONE;
        put(s, " model %d character %d collider %u %u light %d\n", e.modelIDs[i], e.characterIDs[i], unsigned(e.colliderTags[i].shape),
            has ? unsigned(e.colliderTags[i].index) : 0u, e.lightTags[i].type == LIGHT_TYPE_POINT ? int(e.lightTags[i].index) : -1);
    }
    for (size_t i = 0; i < log.modelColliders.size(); ++i) {
        put(s, "modelcollider %zu seq %d model %d", i, log.modelColliders[i].seq, log.modelColliders[i].model);
        putT(s, log.modelColliders[i].t); put(s, "\n");
    }
    for (size_t i = 0; i < log.capsuleColliders.size(); ++i) {
        put(s, "capsulecollider %zu seq %d h %08x r %08x off", i, log.capsuleColliders[i].seq, bits(log.capsuleColliders[i].h),
            bits(log.capsuleColliders[i].r));
        put3(s, log.capsuleColliders[i].off); put(s, "\n");
    }
    for (size_t i = 0; i < log.characters.size(); ++i)
        put(s, "character %zu entity %d collider %u %u\n", i, log.characters[i].entity, unsigned(log.characters[i].tag.shape),
            log.characters[i].tag.shape != COLLIDER_SHAPE_NONE ? unsigned(log.characters[i].tag.index) : 0u);
    for (size_t i = 0; i < log.equipment.size(); ++i) {
        put(s, "equipment %zu character %d entity %d model %d bone \"%s\"", i, log.equipment[i].character, log.equipment[i].entity,
            scene->m_entities.modelIDs[log.equipment[i].entity], log.equipment[i].bone.c_str());
        putT(s, log.equipment[i].offset); put(s, "\n");
    }
    for (size_t i = 0; i < log.animations.size(); ++i) put(s, "animation %zu entity %d\n", i, log.animations[i].entity);
    for (size_t i = 0; i < log.lights.size(); ++i) {
        put(s, "light %zu entity %d color", i, log.lights[i].entityID); put3(s, log.lights[i].color);
        put(s, " clq %08x %08x %08x\n", bits(log.lights[i].constant), bits(log.lights[i].linear), bits(log.lights[i].quadratic));
    }
    delete scene;
    if (needed) *needed = s.size() + 1;
    if (buf && cap) {
        size_t n = s.size() < cap - 1 ? s.size() : cap - 1;
        std::memcpy(buf, s.data(), n);
        buf[n] = '\0';
    }
    return 0;
}
