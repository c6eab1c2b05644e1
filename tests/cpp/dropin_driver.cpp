// Engine-side caller used to demonstrate the drop-in (TEST code).  It only uses the public PhysicsSystem API of the
// reference (physics_system.h:24-77) plus the engine types around it, exactly as Scene / CharacterSystem /
// SceneSerialization do (scene.cpp:185-220, character_system.cpp:55-78, scene_serialization.cpp:145-166).
// tests/cpp/Makefile compiles THIS SAME FILE twice:
//   driver_reference  against the reference's own physics sources (unmodified, from /root/reference)
//   driver_dropin     against prototype2_b200/engine/physics_system.{h,cpp} + libprtc.so
// and tests/test_dropin.py requires the two programs to print byte-identical state dumps.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <chrono>
#include <cstring>
#include <vector>

#include "src/game/system/physics/physics_system.h"
#include "src/game/scene/scene.h"

void (*prt_oracle_contact_hook)(EntityID) = nullptr;   // the shadow Character header's per-contact hook (unused here)

static uint32_t rng_state = 12345u;
static float frand() {                                  // small LCG: identical streams in both builds
    rng_state = rng_state * 1664525u + 1013904223u;
    return float(rng_state >> 8) / 16777216.0f;
}
static float height(float x, float z) { return 0.75f * float(std::sin(0.07 * double(x))) * float(std::cos(0.05 * double(z))); }

static void put(std::vector<uint8_t> & out, void const * p, size_t n) {
    uint8_t const * b = static_cast<uint8_t const *>(p);
    out.insert(out.end(), b, b + n);
}

#ifdef PRTC_DROPIN
// "errors": the drop-in with an error handler installed on a machine where the device side cannot come up (no GPU): every
// call must come back, the handler must be told, nothing may abort
static int g_handled = 0;
static void onError(int status, char const * what, char const * message, void *) {
    ++g_handled;
    std::printf("handler: %s status %d: %s\n", what, status, message);
}
static int errorsMode() {
    PhysicsSystem::setErrorHandler(onError, nullptr);
    PhysicsSystem physics;
    int const afterCreate = g_handled;
    ColliderTag tag = physics.addCapsuleCollider(0.5f, 0.5f, glm::vec3(0.0f, 0.5f, 0.0f));
    glm::vec3 hit(0.0f);
    bool const any = physics.raycast(glm::vec3(0.0f, 5.0f, 0.0f), glm::vec3(0.0f, -1.0f, 0.0f), 10.0f, hit);
    Scene scene;
    Transform tf[1];
    physics.updateCharacters(1.0f / 30.0f, scene, tf);
    std::printf("errors mode: create failed %d, handled %d, last status %d, tag %u, hit %d\n", afterCreate, g_handled, physics.lastStatus(),
                unsigned(tag.index), int(any));
    return 0;
}
#endif

int main(int argc, char ** argv) {
#ifdef PRTC_DROPIN
    if (argc > 1 && std::strcmp(argv[1], "errors") == 0) return errorsMode();
#endif
    int const N = argc > 1 ? std::atoi(argv[1]) : 24;          // terrain cells per side
    int const nChars = argc > 2 ? std::atoi(argv[2]) : 12;
    int const frames = argc > 3 ? std::atoi(argv[3]) : 40;
    char const * outPath = argc > 4 ? argv[4] : "dump.bin";
    if (argc > 5) rng_state = uint32_t(std::strtoul(argv[5], nullptr, 10));   // another crowd on the same level
    // "props": the map is ONE mesh and two one-mesh platforms stand on it, which are removed and re-added during the run
    // (Scene::loadModel, scene.cpp:354-359: removeCollider + addModelCollider).  One mesh per model and last-added-first
    // removal is the envelope in which the reference's removeModelCollider is well defined (physics_system.cpp:141-148:
    // stale tree tags after the arrays are compacted, geometries[] cleared at the MESH start index).
    bool const props = argc > 6 && std::strcmp(argv[6], "props") == 0;

    PhysicsSystem physics;
    Scene scene;

    // a terrain model with one mesh per 2x2 cells, registered like SceneSerialization does for a MODEL collider
    Model model;
    int const block = props ? N : 2;
    for (int bx = 0; bx < N; bx += block) for (int bz = 0; bz < N; bz += block) {
        Model::Mesh mesh;
        mesh.startIndex = model.indexBuffer.size();
        for (int x = bx; x < bx + block && x < N; ++x) for (int z = bz; z < bz + block && z < N; ++z) {
            float const xs[4] = {float(x), float(x), float(x + 1), float(x + 1)};
            float const zs[4] = {float(z), float(z + 1), float(z + 1), float(z)};
            int const order[6] = {0, 1, 2, 0, 2, 3};
            for (int k = 0; k < 6; ++k) {
                Model::Vertex v;
                v.pos = glm::vec3(xs[order[k]], height(xs[order[k]], zs[order[k]]), zs[order[k]]);
                model.indexBuffer.push_back(uint32_t(model.vertexBuffer.size()));
                model.vertexBuffer.push_back(v);
            }
        }
        mesh.numIndices = model.indexBuffer.size() - mesh.startIndex;
        model.meshes.push_back(mesh);
    }
    Transform mapTransform;
    mapTransform.scale = glm::vec3(1.5f, 1.0f, 1.5f);
    ColliderTag mapTag = physics.addModelCollider(model, mapTransform);

    // the platforms: a 4x4-cell slab, one mesh each
    auto slab = [](float y) {
        Model m;
        Model::Mesh mesh;
        mesh.startIndex = 0;
        for (int x = 0; x < 4; ++x) for (int z = 0; z < 4; ++z) {
            float const xs[4] = {float(x), float(x), float(x + 1), float(x + 1)};
            float const zs[4] = {float(z), float(z + 1), float(z + 1), float(z)};
            int const order[6] = {0, 1, 2, 0, 2, 3};
            for (int k = 0; k < 6; ++k) {
                Model::Vertex v;
                v.pos = glm::vec3(xs[order[k]], y, zs[order[k]]);
                m.indexBuffer.push_back(uint32_t(m.vertexBuffer.size()));
                m.vertexBuffer.push_back(v);
            }
        }
        mesh.numIndices = m.indexBuffer.size();
        m.meshes.push_back(mesh);
        return m;
    };
    Model const slabModel = slab(1.2f);
    ColliderTag propTag[2];
    Transform propTransform[2];
    std::vector<uint8_t> dump;
    if (props) {
        propTransform[0].position = glm::vec3(4.0f, 0.0f, 4.0f);
        propTransform[1].position = glm::vec3(0.4f * 1.5f * float(N), 0.3f, 0.5f * 1.5f * float(N));
        for (int k = 0; k < 2; ++k) { propTag[k] = physics.addModelCollider(slabModel, propTransform[k]); put(dump, &propTag[k].index, sizeof(propTag[k].index)); }
    }

    // characters with capsule colliders (cavern.prt:9 values), some of them overlapping each other
    std::vector<Transform> transforms{static_cast<size_t>(nChars)};
    for (int i = 0; i < nChars; ++i) {
        ColliderTag tag = physics.addCapsuleCollider(0.5f, 0.5f, glm::vec3(0.0f, 0.5f, 0.0f));
        Character c;
        c.id = EntityID(i);
        c.physics.colliderTag = tag;
        float const x = 3.0f + frand() * (1.5f * float(N) - 6.0f), z = 3.0f + frand() * (1.5f * float(N) - 6.0f);
        transforms[size_t(i)].position = glm::vec3(x, 1.5f + frand(), z);
        float const yaw = 6.2831853f * frand();
        transforms[size_t(i)].rotation = glm::quat(std::cos(0.5f * yaw), 0.0f, std::sin(0.5f * yaw), 0.0f);
        c.physics.velocity = glm::vec3(-0.2f + 0.4f * frand(), -0.2f * frand(), -0.2f + 0.4f * frand());
        scene.getCharacterSystem().chars.push_back(c);
    }
    if (nChars >= 2) transforms[1].position = transforms[0].position + glm::vec3(0.3f, 0.0f, 0.1f);   // a touching pair

    if (char const * initPath = std::getenv("PRT_DRIVER_DUMP_INIT")) {      // lets a Python harness rebuild exactly this scene
        std::vector<uint8_t> init;
        for (int i = 0; i < nChars; ++i) {
            put(init, &transforms[size_t(i)].position, sizeof(glm::vec3));
            put(init, &transforms[size_t(i)].rotation, sizeof(glm::quat));            // x, y, z, w
            put(init, &scene.getCharacterSystem().getCharacter(size_t(i)).physics.velocity, sizeof(glm::vec3));
        }
        if (FILE * fi = std::fopen(initPath, "wb")) { std::fwrite(init.data(), 1, init.size(), fi); std::fclose(fi); }
    }
    std::vector<double> stepTimes, rayTimes;             // native cost of the two per-frame entry points, microseconds
    typedef std::chrono::steady_clock Clock;
    for (int f = 0; f < frames; ++f) {
        for (int i = 0; i < nChars; ++i) {               // gameplay input, like CharacterSystem::updateCharacters
            CharacterPhysics & p = scene.getCharacterSystem().getCharacter(size_t(i)).physics;
            p.movementVector = (f / 10) % 2 == 0 ? glm::vec3(0.1f, 0.0f, 0.05f) : glm::vec3(0.0f, 0.0f, 0.0f);
            if (f == 15 && p.isGrounded) p.velocity.y = 0.35f;            // jump (character_system.cpp:92)
        }
        if (f == 20) {                                   // the editor moves the map (editor.cpp:40 -> scene.cpp:190-211)
            mapTransform.position = glm::vec3(0.25f, -0.1f, 0.0f);
            physics.updateModelColliders(&mapTag, &mapTransform, 1);
        }
        if (props) {                                     // Scene::loadModel: removeCollider + addModelCollider (scene.cpp:354-359)
            if (f == 30) physics.removeCollider(propTag[1]);
            if (f == 45) {
                propTransform[1].position = propTransform[1].position + glm::vec3(1.0f, 0.4f, -1.0f);
                propTag[1] = physics.addModelCollider(slabModel, propTransform[1]);
                put(dump, &propTag[1].index, sizeof(propTag[1].index));          // the freed index comes back (physics_system.cpp:49-51)
            }
            if (f == 60) { physics.removeCollider(propTag[1]); physics.removeCollider(propTag[0]); }
            if (f == 70) {
                for (int k = 0; k < 2; ++k) {
                    propTransform[k].position = propTransform[k].position + glm::vec3(-0.5f, -0.2f, 0.5f);
                    propTag[k] = physics.addModelCollider(slabModel, propTransform[k]);
                    put(dump, &propTag[k].index, sizeof(propTag[k].index));
                }
            }
        }
        if (f == 25 && nChars > 2)                       // the editor edits a capsule (editor_gui.cpp:293-301)
            physics.updateCapsuleCollider(scene.getCharacterSystem().getCharacter(2).physics.colliderTag, 0.8f, 0.4f, glm::vec3(0.0f, 0.4f, 0.0f));
        Clock::time_point const t0 = Clock::now();
        physics.updateCharacters(1.0f / 30.0f, scene, transforms.data());
        stepTimes.push_back(std::chrono::duration<double, std::micro>(Clock::now() - t0).count());
        for (int i = 0; i < nChars; ++i) {
            CharacterPhysics const & p = scene.getCharacterSystem().getCharacter(size_t(i)).physics;
            put(dump, &transforms[size_t(i)].position, sizeof(glm::vec3));
            put(dump, &p.velocity, sizeof(glm::vec3));
            put(dump, &p.groundNormal, sizeof(glm::vec3));
            uint8_t g = p.isGrounded ? 1 : 0;
            put(dump, &g, 1);
        }
        // camera occlusion rays (scene.cpp:248-255)
        Clock::time_point const t1 = Clock::now();
        for (int k = 0; k < 4; ++k) {
            glm::vec3 origin = transforms[0].position + glm::vec3(0.0f, 1.0f, 0.0f);
            glm::vec3 dir = glm::normalize(glm::vec3(float(k) - 1.5f, -0.8f, 1.0f));
            glm::vec3 hit(0.0f);
            uint8_t r = physics.raycast(origin, dir, 5.0f, hit) ? 1 : 0;
            put(dump, &r, 1);
            if (r) put(dump, &hit, sizeof(glm::vec3));
        }
        rayTimes.push_back(std::chrono::duration<double, std::micro>(Clock::now() - t1).count());
    }
    if (std::getenv("PRT_DRIVER_TIMING") && frames > 10) {
        std::sort(stepTimes.begin() + 5, stepTimes.end());   // the first frames carry one-off setup (tree build / upload)
        std::sort(rayTimes.begin() + 5, rayTimes.end());
        size_t const mid = 5 + (stepTimes.size() - 5) / 2;
        std::fprintf(stderr, "timing: %d characters, updateCharacters median %.1f us, 4 raycasts median %.1f us\n", nChars,
                     stepTimes[mid], rayTimes[mid]);
    }
    FILE * fp = std::fopen(outPath, "wb");
    if (!fp) return 2;
    std::fwrite(dump.data(), 1, dump.size(), fp);
    std::fclose(fp);
    std::printf("%s: %d characters, %d frames, %zu bytes, gravity %.1f\n", outPath, nChars, frames, dump.size(), double(physics.getGravity()));
    return 0;
}
