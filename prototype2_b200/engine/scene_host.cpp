// scene_host.cpp -- libprtscene.so: `.prt` scene files, OBJ collider geometry and the headless per-frame caller of the
// collision path (C-ABI in include/prt_scene.h).
//
// Three parts, each the B200-side counterpart of an engine component that sits next to PhysicsSystem:
//   SceneText   the grammar of SceneSerialization::loadScene / parseEntity (reference
//               src/game/scene/scene_serialization.cpp:10-345), including how it skips text: every value is "scan to the
//               next '<', read, scan to the next '>'", every component ends at the next '}', every line at the next '\n'
//   ObjModel    the three arrays PhysicsSystem::addModelCollider reads from a Model (physics_system.cpp:60-85)
//   World       Scene::updatePhysics + CharacterSystem::updatePhysics (scene.cpp:185-211, character_system.cpp:55-78)
//               on top of prtc_* (include/prtc.h)
// Where the reference asserts (a release build of it then reads on through garbage), this code fails with PRTS_ERR_PARSE.
#include "prt_scene.h"

#include <cctype>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

namespace {

thread_local std::string g_error;

int fail(int status, char const * fmt, ...) __attribute__((format(printf, 2, 3)));
int fail(int status, char const * fmt, ...) {
    char msg[512];
    va_list ap;
    va_start(ap, fmt);
    std::vsnprintf(msg, sizeof msg, fmt, ap);
    va_end(ap);
    g_error = msg;
    return status;
}

bool read_file(char const * path, std::string & out) {
    std::ifstream in(path, std::ios::binary);
    if (!in) return false;
    std::ostringstream ss;
    ss << in.rdbuf();
    out = ss.str();
    return true;
}

prtc_transform identity_transform() {
    prtc_transform t;
    t.position[0] = t.position[1] = t.position[2] = 0.0f;
    t.rotation[0] = t.rotation[1] = t.rotation[2] = 0.0f;
    t.rotation[3] = 1.0f;
    t.scale[0] = t.scale[1] = t.scale[2] = 1.0f;
    return t;
}

// ---------------------------------------------------------------------------------------------------------------------
// scene text
// ---------------------------------------------------------------------------------------------------------------------
struct ColliderEvent {          // one PhysicsSystem::add*Collider call, in file order
    uint32_t shape;
    int32_t entity;
    int32_t model;              // MODEL
    prtc_transform transform;   // MODEL: the entity's transform AS PARSED SO FAR (scene_serialization.cpp:153-154)
    float height, radius, offset[3];
    uint32_t index;             // ColliderTag::index
};
struct ModelEntry { std::string path; bool animated; };
struct CharacterEntry { int32_t entity; uint32_t shape, index; };

}  // namespace

struct prts_scene {
    std::vector<prts_entity> entities;
    std::vector<ModelEntry> models;
    std::vector<ColliderEvent> colliders;
    std::vector<CharacterEntry> characters;
    std::vector<prts_equipment> equipment;
    std::vector<int32_t> animations;           // entity per AnimationSystem::addAnimation call
    std::vector<prts_point_light> lights;
    prts_sun sun;
    uint32_t n_model_colliders = 0, n_capsule_colliders = 0;
};

namespace {

class SceneText {
public:
    SceneText(char const * text, size_t len, prts_scene & scene) : m_text(text, len), m_scene(scene) {
        // the reference reads a NUL-terminated buffer (:42): an embedded NUL ends the text
        m_text.resize(std::strlen(m_text.c_str()));
        m_p = m_text.c_str();
    }

    int run() {
        std::memset(&m_scene.sun, 0, sizeof m_scene.sun);
        char const * const end = m_text.c_str() + m_text.size();
        while (m_p < end) {
            // readToken (:73-93): everything up to '[' is the token
            std::string token;
            while (*m_p != '[' && *m_p != '\0') token += *m_p++;
            if (token.empty()) return PRTS_OK;                              // END
            int st;
            if (token == "Entity") st = entity();
            else if (token == "Sun") st = sun();
            else return fail(PRTS_ERR_PARSE, "line %d: unknown token '%.40s'", line(), token.c_str());
            if (st != PRTS_OK) return st;
            while (*m_p != '\n' && *m_p != '\0') ++m_p;                     // rest of the line (:65-68)
            if (*m_p == '\0') return PRTS_OK;
            ++m_p;
        }
        return PRTS_OK;
    }

private:
    std::string m_text;
    char const * m_p;
    prts_scene & m_scene;

    int line() const {
        int n = 1;
        for (char const * q = m_text.c_str(); q < m_p; ++q) n += *q == '\n';
        return n;
    }
    int eof(char const * what) { return fail(PRTS_ERR_PARSE, "line %d: text ends inside %s", line(), what); }

    // the value readers: scan to '<', read, scan past '>'
    bool open(char const * what, int & st) {
        while (*m_p != '<' && *m_p != '\0') ++m_p;
        if (*m_p == '\0') { st = eof(what); return false; }
        ++m_p;
        return true;
    }
    bool close(char const * what, int & st) {
        while (*m_p != '>' && *m_p != '\0') ++m_p;
        if (*m_p == '\0') { st = eof(what); return false; }
        ++m_p;
        return true;
    }
    int floats(char const * what, char const * fmt, int n, float * out) {
        int st = PRTS_OK;
        if (!open(what, st)) return st;
        float v[4] = {0, 0, 0, 0};
        int got = std::sscanf(m_p, fmt, &v[0], &v[1], &v[2], &v[3]);   // same conversions as parseFloat / parseVec3 / parseQuat
        if (got != n) return fail(PRTS_ERR_PARSE, "line %d: expected %d number(s) in %s", line(), n, what);
        for (int i = 0; i < n; ++i) out[i] = v[i];
        if (!close(what, st)) return st;
        return PRTS_OK;
    }
    int number(float & f) { return floats("a float", "%f", 1, &f); }
    int vec3(float * v) { return floats("a vec3", "%f,%f,%f", 3, v); }
    int quat(float * xyzw) {                                              // file order is w,x,y,z (:333-334)
        float q[4];
        int st = floats("a quat", "%f,%f,%f,%f", 4, q);
        if (st != PRTS_OK) return st;
        float const w = q[0], x = q[1], y = q[2], z = q[3];
        // glm::normalize(quat): length = sqrt((w*w + x*x) + (y*y + z*z)); identity when it is <= 0; else q * (1/length)
        float const len = std::sqrt((w * w + x * x) + (y * y + z * z));
        if (len <= 0.0f) { xyzw[0] = xyzw[1] = xyzw[2] = 0.0f; xyzw[3] = 1.0f; return PRTS_OK; }
        float const inv = 1.0f / len;
        xyzw[0] = x * inv; xyzw[1] = y * inv; xyzw[2] = z * inv; xyzw[3] = w * inv;
        return PRTS_OK;
    }
    int string(std::string & out, size_t capacity) {                      // parseString (:226-251): <"...">
        int st = PRTS_OK;
        while (*m_p != '<' && *m_p != '\0') ++m_p;
        if (*m_p == '\0') return eof("a string");
        if (m_p[1] == '\0') return eof("a string");
        m_p += 2;                                                         // '<' and the opening quote, unchecked
        out.clear();
        while (*m_p != '"') {
            if (*m_p == '\0') return eof("a string");
            out += *m_p++;
        }
        if (out.size() >= capacity) return fail(PRTS_ERR_PARSE, "line %d: string longer than %zu characters", line(), capacity - 1);
        if (!close("a string", st)) return st;
        return PRTS_OK;
    }
    int boolean(bool & b) {                                               // parseBool (:253-263): first letter only, '>' left for the next scan
        int st = PRTS_OK;
        if (!open("a bool", st)) return st;
        b = std::tolower(static_cast<unsigned char>(*m_p)) == 't';
        return PRTS_OK;
    }
    int transform(prtc_transform & t) {
        int st;
        if ((st = vec3(t.position)) != PRTS_OK) return st;
        if ((st = quat(t.rotation)) != PRTS_OK) return st;
        return vec3(t.scale);
    }

    int new_entity(int32_t & id) {
        if (m_scene.entities.size() >= PRTS_MAX_ENTITIES)
            return fail(PRTS_ERR_LIMIT, "line %d: more than %d entities", line(), PRTS_MAX_ENTITIES);
        prts_entity e;
        std::memset(&e, 0, sizeof e);
        id = int32_t(m_scene.entities.size());
        std::snprintf(e.name, sizeof e.name, "entity_%lu", static_cast<unsigned long>(id));    // :13
        e.transform = identity_transform();
        e.model_id = e.animation_id = e.character_id = e.light_index = e.equipment_of = -1;
        e.collider_shape = PRTS_SHAPE_NONE;
        m_scene.entities.push_back(e);
        return PRTS_OK;
    }
    int32_t load_model(std::string const & path, bool animated) {        // ModelManager::loadModel: one entry per path
        for (size_t i = 0; i < m_scene.models.size(); ++i)
            if (m_scene.models[i].path == path) return int32_t(i);
        m_scene.models.push_back({path, animated});
        return int32_t(m_scene.models.size()) - 1;
    }

    // readComponentType (:95-111): the digits in front of '{'.  The first one of an entity still has the '[' of
    // "Entity[" in front of it, which atoi cannot read -> the first component is ALWAYS read as type 0.
    int component_type() {
        std::string token;
        while (*m_p != '{' && *m_p != '\n' && *m_p != '\0') token += *m_p++;
        if (*m_p == '\n' || *m_p == '\0') return -1;
        return std::atoi(token.c_str());
    }

    int entity() {
        int32_t id;
        int st = new_entity(id);
        if (st != PRTS_OK) return st;
        for (int type; (type = component_type()) != -1;) {
            // m_scene.entities may reallocate while equipment entities are added: always index, never hold a reference
            switch (type) {
            case 0: {
                std::string name;
                if ((st = string(name, PRTS_NAME_LEN)) != PRTS_OK) return st;
                std::snprintf(m_scene.entities[id].name, PRTS_NAME_LEN, "%s", name.c_str());
                break;
            }
            case 1: {
                prtc_transform t = m_scene.entities[id].transform;
                if ((st = transform(t)) != PRTS_OK) return st;
                m_scene.entities[id].transform = t;
                break;
            }
            case 2: {
                std::string path;
                bool animated = false;
                if ((st = string(path, PRTS_PATH_LEN)) != PRTS_OK) return st;
                if ((st = boolean(animated)) != PRTS_OK) return st;
                m_scene.entities[id].model_id = load_model(path, animated);
                m_scene.entities[id].animated = animated;
                if (animated) {
                    m_scene.animations.push_back(id);
                    m_scene.entities[id].animation_id = int32_t(m_scene.animations.size()) - 1;
                }
                break;
            }
            case 3: {
                std::string kind;
                if ((st = string(kind, PRTS_PATH_LEN)) != PRTS_OK) return st;
                ColliderEvent ev;
                std::memset(&ev, 0, sizeof ev);
                ev.entity = id;
                if (kind == "MODEL") {
                    if (m_scene.entities[id].model_id < 0)
                        return fail(PRTS_ERR_PARSE, "line %d: MODEL collider on an entity without a model component", line());
                    ev.shape = PRTS_SHAPE_MODEL;
                    ev.model = m_scene.entities[id].model_id;
                    ev.transform = m_scene.entities[id].transform;
                    ev.index = m_scene.n_model_colliders++;
                } else if (kind == "CAPSULE") {
                    ev.shape = PRTS_SHAPE_CAPSULE;
                    ev.model = -1;
                    if ((st = number(ev.height)) != PRTS_OK) return st;
                    if ((st = number(ev.radius)) != PRTS_OK) return st;
                    if ((st = vec3(ev.offset)) != PRTS_OK) return st;
                    ev.index = m_scene.n_capsule_colliders++;
                    prts_entity & e = m_scene.entities[id];
                    e.capsule_height = ev.height;
                    e.capsule_radius = ev.radius;
                    std::memcpy(e.capsule_offset, ev.offset, sizeof ev.offset);
                } else {
                    return fail(PRTS_ERR_PARSE, "line %d: unknown collider type '%.40s'", line(), kind.c_str());
                }
                if (ev.index > 0xFFFFu) return fail(PRTS_ERR_LIMIT, "line %d: collider index exceeds ColliderIndex (uint16)", line());
                m_scene.entities[id].collider_shape = ev.shape;
                m_scene.entities[id].collider_index = ev.index;
                m_scene.colliders.push_back(ev);
                break;
            }
            case 4: {
                CharacterEntry c;
                c.entity = id;
                c.shape = m_scene.entities[id].collider_shape;               // colliderTags[id] as it is NOW (:169)
                c.index = m_scene.entities[id].collider_index;
                m_scene.characters.push_back(c);
                int32_t const character = int32_t(m_scene.characters.size()) - 1;
                m_scene.entities[id].character_id = character;
                ++m_p;                                                       // the '{'
                while (*m_p == '<') {                                        // equipment list: <<"bone"><"model"><pos><quat><scale>>
                    ++m_p;
                    std::string bone, model;
                    if ((st = string(bone, PRTS_PATH_LEN)) != PRTS_OK) return st;
                    if ((st = string(model, PRTS_PATH_LEN)) != PRTS_OK) return st;
                    int32_t equip = -1;
                    if ((st = new_entity(equip)) != PRTS_OK) return st;
                    prts_equipment q;
                    std::memset(&q, 0, sizeof q);
                    q.character_id = character;
                    q.entity = equip;
                    q.model_id = load_model(model, false);
                    std::snprintf(q.bone, sizeof q.bone, "%s", bone.c_str());
                    q.offset = identity_transform();
                    if ((st = transform(q.offset)) != PRTS_OK) return st;
                    if (*m_p != '\0') ++m_p;                                 // the closing '>' of the item
                    m_scene.entities[equip].model_id = q.model_id;
                    m_scene.entities[equip].equipment_of = character;
                    m_scene.equipment.push_back(q);
                }
                break;
            }
            case 5: {
                prts_point_light l;
                std::memset(&l, 0, sizeof l);
                l.entity = id;
                if ((st = vec3(l.color)) != PRTS_OK) return st;
                if ((st = number(l.constant)) != PRTS_OK) return st;
                if ((st = number(l.linear)) != PRTS_OK) return st;
                if ((st = number(l.quadratic)) != PRTS_OK) return st;
                m_scene.lights.push_back(l);
                m_scene.entities[id].light_index = int32_t(m_scene.lights.size()) - 1;
                break;
            }
            default:
                break;                                                       // unknown component: skipped (:209-211)
            }
            while (*m_p != '}' && *m_p != '\0') ++m_p;                       // :213-216
            if (*m_p == '\0') return PRTS_OK;
            ++m_p;
        }
        return PRTS_OK;
    }

    int sun() {                                                              // parseSun (:221-224)
        float d[3], c[3];
        int st;
        if ((st = vec3(d)) != PRTS_OK) return st;
        if ((st = vec3(c)) != PRTS_OK) return st;
        // glm::normalize(vec3) = v * (1 / sqrt((x*x + y*y) + z*z))
        float const inv = 1.0f / std::sqrt((d[0] * d[0] + d[1] * d[1]) + d[2] * d[2]);
        m_scene.sun.present = 1;
        for (int i = 0; i < 3; ++i) { m_scene.sun.direction[i] = d[i] * inv; m_scene.sun.color[i] = c[i]; }
        return PRTS_OK;
    }
};

uint32_t bits(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
void append(std::string & s, char const * fmt, ...) __attribute__((format(printf, 2, 3)));
void append(std::string & s, char const * fmt, ...) {
    char line[640];
    va_list ap;
    va_start(ap, fmt);
    std::vsnprintf(line, sizeof line, fmt, ap);
    va_end(ap);
    s += line;
}
void append3(std::string & s, float const * v) { append(s, " %08x %08x %08x", bits(v[0]), bits(v[1]), bits(v[2])); }
void append_transform(std::string & s, prtc_transform const & t) {
    append(s, " pos"); append3(s, t.position);
    append(s, " rot %08x %08x %08x %08x", bits(t.rotation[0]), bits(t.rotation[1]), bits(t.rotation[2]), bits(t.rotation[3]));
    append(s, " scale"); append3(s, t.scale);
}

// ---------------------------------------------------------------------------------------------------------------------
// OBJ geometry
// ---------------------------------------------------------------------------------------------------------------------
}  // namespace

struct prts_model {
    std::vector<float> xyz;
    std::vector<uint32_t> indices, mesh_start, mesh_count;
};

namespace {

int parse_obj(std::string const & text, prts_model & m) {
    std::istringstream in(text);
    std::string row;
    bool open_mesh = false;
    auto begin_mesh = [&]() { open_mesh = false; };                          // the next face opens a new mesh
    int lineno = 0;
    while (std::getline(in, row)) {
        ++lineno;
        if (!row.empty() && row.back() == '\r') row.pop_back();
        char const * p = row.c_str();
        while (*p == ' ' || *p == '\t') ++p;
        if (p[0] == 'v' && (p[1] == ' ' || p[1] == '\t')) {
            float v[3];
            if (std::sscanf(p + 1, "%f %f %f", &v[0], &v[1], &v[2]) != 3) return fail(PRTS_ERR_PARSE, "obj line %d: bad vertex", lineno);
            m.xyz.insert(m.xyz.end(), v, v + 3);
        } else if ((p[0] == 'o' || p[0] == 'g') && (p[1] == ' ' || p[1] == '\t' || p[1] == '\0')) {
            begin_mesh();
        } else if (std::strncmp(p, "usemtl", 6) == 0) {
            begin_mesh();
        } else if (p[0] == 'f' && (p[1] == ' ' || p[1] == '\t')) {
            std::vector<uint32_t> corner;
            char const * q = p + 1;
            long const nv = long(m.xyz.size() / 3);
            for (;;) {
                while (*q == ' ' || *q == '\t') ++q;
                if (*q == '\0') break;
                char * after;
                long idx = std::strtol(q, &after, 10);
                if (after == q) return fail(PRTS_ERR_PARSE, "obj line %d: bad face", lineno);
                if (idx < 0) idx = nv + idx; else idx -= 1;                    // negative = relative to the vertices read so far
                if (idx < 0 || idx >= nv) return fail(PRTS_ERR_PARSE, "obj line %d: vertex index out of range", lineno);
                corner.push_back(uint32_t(idx));
                q = after;
                while (*q != ' ' && *q != '\t' && *q != '\0') ++q;             // the /vt/vn part
            }
            if (corner.size() < 3) return fail(PRTS_ERR_PARSE, "obj line %d: face with fewer than 3 corners", lineno);
            if (!open_mesh) {
                m.mesh_start.push_back(uint32_t(m.indices.size()));
                m.mesh_count.push_back(0);
                open_mesh = true;
            }
            for (size_t k = 1; k + 1 < corner.size(); ++k) {                   // fan
                m.indices.push_back(corner[0]);
                m.indices.push_back(corner[k]);
                m.indices.push_back(corner[k + 1]);
                m.mesh_count.back() += 3;
            }
        }
    }
    return PRTS_OK;
}

bool has_suffix(std::string const & s, char const * suffix) {
    size_t n = std::strlen(suffix);
    if (s.size() < n) return false;
    for (size_t i = 0; i < n; ++i)
        if (std::tolower(static_cast<unsigned char>(s[s.size() - n + i])) != suffix[i]) return false;
    return true;
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------------------
// headless scene host
// ---------------------------------------------------------------------------------------------------------------------
struct prts_world {
    prtc_ctx * ctx = nullptr;
    std::vector<prtc_transform> transforms;          // Entities::transforms
    std::vector<int32_t> model_collider;             // per entity: prtc model id of its MODEL collider, -1
    std::vector<int32_t> character_entity;
    std::vector<prtc_char_physics> physics;          // Character::physics, one per character
    std::vector<prtc_transform> gathered;            // scratch of CharacterSystem::updatePhysics
    std::vector<uint32_t> moved;                     // m_colliderUpdateSet, insertion order, unique
};

namespace {
int physics_fail(char const * what) { return fail(PRTS_ERR_PHYSICS, "%s: %s", what, prtc_last_error()); }
}

extern "C" {

const char * prts_last_error(void) { return g_error.c_str(); }

int prts_scene_parse(const char * text, size_t len, prts_scene ** out) {
    if (!out || (!text && len)) return fail(PRTS_ERR_INVALID, "prts_scene_parse: null argument");
    prts_scene * scene = new prts_scene();
    SceneText reader(text ? text : "", len, *scene);
    int st = reader.run();
    if (st != PRTS_OK) { delete scene; return st; }
    *out = scene;
    return PRTS_OK;
}

int prts_scene_load(const char * path, prts_scene ** out) {
    if (!path || !out) return fail(PRTS_ERR_INVALID, "prts_scene_load: null argument");
    std::string text;
    if (!read_file(path, text)) return fail(PRTS_ERR_IO, "cannot read scene file '%s'", path);
    return prts_scene_parse(text.data(), text.size(), out);
}

void prts_scene_free(prts_scene * scene) { delete scene; }

uint32_t prts_entity_count(const prts_scene * s) { return s ? uint32_t(s->entities.size()) : 0; }
int prts_get_entity(const prts_scene * s, uint32_t i, prts_entity * out) {
    if (!s || !out || i >= s->entities.size()) return fail(PRTS_ERR_INVALID, "prts_get_entity: bad argument");
    *out = s->entities[i];
    return PRTS_OK;
}
uint32_t prts_model_count(const prts_scene * s) { return s ? uint32_t(s->models.size()) : 0; }
int prts_get_model(const prts_scene * s, uint32_t i, char path[PRTS_PATH_LEN], int32_t * animated) {
    if (!s || !path || i >= s->models.size()) return fail(PRTS_ERR_INVALID, "prts_get_model: bad argument");
    std::snprintf(path, PRTS_PATH_LEN, "%s", s->models[i].path.c_str());
    if (animated) *animated = s->models[i].animated;
    return PRTS_OK;
}
uint32_t prts_character_count(const prts_scene * s) { return s ? uint32_t(s->characters.size()) : 0; }
int prts_get_character(const prts_scene * s, uint32_t i, int32_t * entity) {
    if (!s || !entity || i >= s->characters.size()) return fail(PRTS_ERR_INVALID, "prts_get_character: bad argument");
    *entity = s->characters[i].entity;
    return PRTS_OK;
}
uint32_t prts_equipment_count(const prts_scene * s) { return s ? uint32_t(s->equipment.size()) : 0; }
int prts_get_equipment(const prts_scene * s, uint32_t i, prts_equipment * out) {
    if (!s || !out || i >= s->equipment.size()) return fail(PRTS_ERR_INVALID, "prts_get_equipment: bad argument");
    *out = s->equipment[i];
    return PRTS_OK;
}
uint32_t prts_light_count(const prts_scene * s) { return s ? uint32_t(s->lights.size()) : 0; }
int prts_get_light(const prts_scene * s, uint32_t i, prts_point_light * out) {
    if (!s || !out || i >= s->lights.size()) return fail(PRTS_ERR_INVALID, "prts_get_light: bad argument");
    *out = s->lights[i];
    return PRTS_OK;
}
int prts_get_sun(const prts_scene * s, prts_sun * out) {
    if (!s || !out) return fail(PRTS_ERR_INVALID, "prts_get_sun: null argument");
    *out = s->sun;
    return PRTS_OK;
}

int prts_scene_dump(const prts_scene * s, char * buf, size_t cap, size_t * needed) {
    if (!s) return fail(PRTS_ERR_INVALID, "prts_scene_dump: null scene");
    std::string t;
    append(t, "sun dir"); append3(t, s->sun.direction); append(t, " color"); append3(t, s->sun.color); append(t, "\n");
    for (size_t i = 0; i < s->models.size(); ++i)
        append(t, "model %zu \"%s\" animated %d\n", i, s->models[i].path.c_str(), int(s->models[i].animated));
    for (size_t i = 0; i < s->entities.size(); ++i) {
        prts_entity const & e = s->entities[i];
        append(t, "entity %zu name \"%s\"", i, e.name); append_transform(t, e.transform);
        append(t, " model %d character %d collider %u %u light %d\n", e.model_id, e.character_id, e.collider_shape,
               e.collider_shape != PRTS_SHAPE_NONE ? e.collider_index : 0u, e.light_index);
    }
    for (int pass = 0; pass < 2; ++pass)
        for (size_t i = 0; i < s->colliders.size(); ++i) {
            ColliderEvent const & c = s->colliders[i];
            if (pass == 0 && c.shape == PRTS_SHAPE_MODEL) {
                append(t, "modelcollider %u seq %zu model %d", c.index, i, c.model); append_transform(t, c.transform); append(t, "\n");
            } else if (pass == 1 && c.shape == PRTS_SHAPE_CAPSULE) {
                append(t, "capsulecollider %u seq %zu h %08x r %08x off", c.index, i, bits(c.height), bits(c.radius));
                append3(t, c.offset); append(t, "\n");
            }
        }
    for (size_t i = 0; i < s->characters.size(); ++i)
        append(t, "character %zu entity %d collider %u %u\n", i, s->characters[i].entity, s->characters[i].shape,
               s->characters[i].shape != PRTS_SHAPE_NONE ? s->characters[i].index : 0u);
    for (size_t i = 0; i < s->equipment.size(); ++i) {
        prts_equipment const & q = s->equipment[i];
        append(t, "equipment %zu character %d entity %d model %d bone \"%s\"", i, q.character_id, q.entity, q.model_id, q.bone);
        append_transform(t, q.offset); append(t, "\n");
    }
    for (size_t i = 0; i < s->animations.size(); ++i) append(t, "animation %zu entity %d\n", i, s->animations[i]);
    for (size_t i = 0; i < s->lights.size(); ++i) {
        prts_point_light const & l = s->lights[i];
        append(t, "light %zu entity %d color", i, l.entity); append3(t, l.color);
        append(t, " clq %08x %08x %08x\n", bits(l.constant), bits(l.linear), bits(l.quadratic));
    }
    if (needed) *needed = t.size() + 1;
    if (buf && cap) {
        size_t n = t.size() < cap - 1 ? t.size() : cap - 1;
        std::memcpy(buf, t.data(), n);
        buf[n] = '\0';
    }
    return PRTS_OK;
}

int prts_model_parse_obj(const char * text, size_t len, prts_model ** out) {
    if (!out || (!text && len)) return fail(PRTS_ERR_INVALID, "prts_model_parse_obj: null argument");
    prts_model * m = new prts_model();
    int st = parse_obj(std::string(text ? text : "", len), *m);
    if (st != PRTS_OK) { delete m; return st; }
    *out = m;
    return PRTS_OK;
}

int prts_model_load(const char * path, prts_model ** out) {
    if (!path || !out) return fail(PRTS_ERR_INVALID, "prts_model_load: null argument");
    if (!has_suffix(path, ".obj")) return fail(PRTS_ERR_IO, "unsupported model format '%s' (only .obj geometry is read)", path);
    std::string text;
    if (!read_file(path, text)) return fail(PRTS_ERR_IO, "cannot read model file '%s'", path);
    return prts_model_parse_obj(text.data(), text.size(), out);
}

void prts_model_free(prts_model * model) { delete model; }

int prts_model_info(const prts_model * m, uint32_t * n_vertices, uint32_t * n_indices, uint32_t * n_meshes) {
    if (!m) return fail(PRTS_ERR_INVALID, "prts_model_info: null model");
    if (n_vertices) *n_vertices = uint32_t(m->xyz.size() / 3);
    if (n_indices) *n_indices = uint32_t(m->indices.size());
    if (n_meshes) *n_meshes = uint32_t(m->mesh_start.size());
    return PRTS_OK;
}

int prts_model_data(const prts_model * m, const float ** xyz, const uint32_t ** indices, const uint32_t ** mesh_start_index,
                    const uint32_t ** mesh_num_indices) {
    if (!m) return fail(PRTS_ERR_INVALID, "prts_model_data: null model");
    if (xyz) *xyz = m->xyz.data();
    if (indices) *indices = m->indices.data();
    if (mesh_start_index) *mesh_start_index = m->mesh_start.data();
    if (mesh_num_indices) *mesh_num_indices = m->mesh_count.data();
    return PRTS_OK;
}

int prts_world_create(const prts_scene * scene, const char * model_dir, const prtc_config * cfg, prts_world ** out) {
    if (!scene || !out) return fail(PRTS_ERR_INVALID, "prts_world_create: null argument");
    prtc_config c;
    if (cfg) c = *cfg;
    else { prtc_default_config(&c); c.order_mode = PRTC_ORDER_LITERAL; }

    // geometry first, so a missing file fails before a device context exists
    std::vector<prts_model *> loaded(scene->models.size(), nullptr);
    auto release_models = [&]() { for (prts_model * m : loaded) delete m; };
    for (ColliderEvent const & ev : scene->colliders) {
        if (ev.shape != PRTS_SHAPE_MODEL || loaded[size_t(ev.model)]) continue;
        std::string path = std::string(model_dir ? model_dir : "") + scene->models[size_t(ev.model)].path;
        int st = prts_model_load(path.c_str(), &loaded[size_t(ev.model)]);
        if (st != PRTS_OK) { release_models(); return st; }
    }
    for (CharacterEntry const & ch : scene->characters)
        if (ch.shape != PRTS_SHAPE_CAPSULE) {
            release_models();
            return fail(PRTS_ERR_INVALID, "character on entity %d has no CAPSULE collider (component 3 must precede component 4)", ch.entity);
        }

    prts_world * w = new prts_world();
    if (prtc_create(&c, &w->ctx) != PRTC_OK) { delete w; release_models(); return physics_fail("prtc_create"); }
    auto abandon = [&](int st) { prtc_destroy(w->ctx); delete w; release_models(); return st; };

    w->transforms.resize(scene->entities.size());
    w->model_collider.assign(scene->entities.size(), -1);
    for (size_t i = 0; i < scene->entities.size(); ++i) w->transforms[i] = scene->entities[i].transform;

    // colliders in file order: the unified tree of the LITERAL order depends on it (aabb_tree.cpp:134-180)
    std::vector<float> expanded;
    std::vector<uint32_t> first, count;
    for (ColliderEvent const & ev : scene->colliders) {
        if (ev.shape == PRTS_SHAPE_MODEL) {
            prts_model const & m = *loaded[size_t(ev.model)];
            // positions through the index buffer, mesh by mesh (physics_system.cpp:68-91)
            expanded.clear(); first.clear(); count.clear();
            for (size_t k = 0; k < m.mesh_start.size(); ++k) {
                first.push_back(uint32_t(expanded.size() / 3));
                count.push_back(m.mesh_count[k]);
                for (uint32_t j = 0; j < m.mesh_count[k]; ++j) {
                    uint32_t const v = m.indices[m.mesh_start[k] + j];
                    expanded.insert(expanded.end(), &m.xyz[3 * size_t(v)], &m.xyz[3 * size_t(v)] + 3);
                }
            }
            uint32_t model_id = 0, first_mesh = 0;
            if (prtc_add_model(w->ctx, expanded.data(), uint32_t(expanded.size() / 3), first.data(), count.data(), uint32_t(first.size()),
                               &ev.transform, &model_id, &first_mesh) != PRTC_OK)
                return abandon(physics_fail("prtc_add_model"));
            w->model_collider[size_t(ev.entity)] = int32_t(model_id);
        } else {
            uint32_t id = 0;
            if (prtc_add_capsule(w->ctx, ev.height, ev.radius, ev.offset, &id) != PRTC_OK) return abandon(physics_fail("prtc_add_capsule"));
            if (id != ev.index) return abandon(fail(PRTS_ERR_PHYSICS, "capsule id %u does not match ColliderTag index %u", id, ev.index));
        }
    }
    release_models();
    if (prtc_commit(w->ctx) != PRTC_OK) { int st = physics_fail("prtc_commit"); prtc_destroy(w->ctx); delete w; return st; }

    for (CharacterEntry const & ch : scene->characters) {
        prtc_char_physics p;                                 // CharacterPhysics defaults (character.h:19-27)
        std::memset(&p, 0, sizeof p);
        p.collider = ch.index;
        w->physics.push_back(p);
        w->character_entity.push_back(ch.entity);
    }
    w->gathered.resize(w->physics.size());
    *out = w;
    return PRTS_OK;
}

void prts_world_destroy(prts_world * w) {
    if (!w) return;
    prtc_destroy(w->ctx);
    delete w;
}

prtc_ctx * prts_world_ctx(prts_world * w) { return w ? w->ctx : nullptr; }
uint32_t prts_world_entity_count(const prts_world * w) { return w ? uint32_t(w->transforms.size()) : 0; }
uint32_t prts_world_character_count(const prts_world * w) { return w ? uint32_t(w->physics.size()) : 0; }
prtc_transform * prts_world_transforms(prts_world * w) { return w ? w->transforms.data() : nullptr; }
prtc_char_physics * prts_world_physics(prts_world * w) { return w ? w->physics.data() : nullptr; }
const int32_t * prts_world_character_entities(const prts_world * w) { return w ? w->character_entity.data() : nullptr; }

int prts_world_move_entity(prts_world * w, uint32_t entity, const prtc_transform * t) {
    if (!w || !t || entity >= w->transforms.size()) return fail(PRTS_ERR_INVALID, "prts_world_move_entity: bad argument");
    w->transforms[entity] = *t;
    if (w->model_collider[entity] >= 0) {
        bool queued = false;
        for (uint32_t e : w->moved) queued |= e == entity;
        if (!queued) w->moved.push_back(entity);
    }
    return PRTS_OK;
}

int prts_world_update_physics(prts_world * w, float dt) {
    if (!w) return fail(PRTS_ERR_INVALID, "prts_world_update_physics: null world");
    // Scene::updateColliders (scene.cpp:190-211)
    if (!w->moved.empty()) {
        std::vector<uint32_t> ids;
        std::vector<prtc_transform> tf;
        for (uint32_t e : w->moved) { ids.push_back(uint32_t(w->model_collider[e])); tf.push_back(w->transforms[e]); }
        w->moved.clear();
        if (prtc_update_models(w->ctx, ids.data(), tf.data(), uint32_t(ids.size())) != PRTC_OK) return physics_fail("prtc_update_models");
    }
    // CharacterSystem::updatePhysics (character_system.cpp:55-78): gather, step, scatter
    size_t const n = w->physics.size();
    if (n == 0) return PRTS_OK;
    for (size_t i = 0; i < n; ++i) w->gathered[i] = w->transforms[size_t(w->character_entity[i])];
    if (prtc_step_characters(w->ctx, dt, uint32_t(n), w->gathered.data(), w->physics.data()) != PRTC_OK)
        return physics_fail("prtc_step_characters");
    for (size_t i = 0; i < n; ++i) w->transforms[size_t(w->character_entity[i])] = w->gathered[i];
    return PRTS_OK;
}

int prts_world_raycast(prts_world * w, uint32_t n, const float * origins, const float * directions, const float * max_distances,
                       float * hit_xyz, uint8_t * hit_mask) {
    if (!w) return fail(PRTS_ERR_INVALID, "prts_world_raycast: null world");
    if (prtc_raycast(w->ctx, n, origins, directions, max_distances, hit_xyz, hit_mask) != PRTC_OK) return physics_fail("prtc_raycast");
    return PRTS_OK;
}

}  // extern "C"
