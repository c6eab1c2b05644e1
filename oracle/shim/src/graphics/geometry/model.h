// ORACLE / TEST INFRASTRUCTURE ONLY.  Shadow of the engine's asset class: the
// real header pulls in assimp and Vulkan, which are absent here.  Only the
// three members PhysicsSystem::addModelCollider reads
// (reference physics_system.cpp:60-85) are provided.
#ifndef PRT_ORACLE_SHADOW_MODEL_H
#define PRT_ORACLE_SHADOW_MODEL_H
#include "src/container/vector.h"
#include <glm/glm.hpp>
#include <cstdint>
#include <cstddef>

class Model {
public:
    struct Mesh { size_t startIndex = 0; size_t numIndices = 0; };
    struct Vertex { glm::vec3 pos; };
    prt::vector<Mesh> meshes;
    prt::vector<Vertex> vertexBuffer;
    prt::vector<uint32_t> indexBuffer;
};
#endif
