// ORACLE / TEST INFRASTRUCTURE ONLY.  Stands in for the header CMake would
// generate from src/config/prototype2Config.h.in (the checked-in copy in the
// reference is stale: 64 MB arena).  Only the container-arena constants matter
// to the collision path; the arena is enlarged so 64 000-capsule scenes fit.
#ifndef PROTOTYPE_2_CONFIG_H
#define PROTOTYPE_2_CONFIG_H
#define PRT_BIG_ENDIAN 0
#define RESOURCE_PATH ""
#ifndef DEFAULT_CONTAINER_ALLOCATOR_SIZE_BYTES
#define DEFAULT_CONTAINER_ALLOCATOR_SIZE_BYTES (1536ull*1024ull*1024ull)
#endif
#define DEFAULT_CONTAINER_ALLOCATOR_BLOCK_SIZE_BYTES 256
#define DEFAULT_CONTAINER_ALLOCATOR_ALIGNMENT_BYTES 4
#define FRAME_RATE 30
#endif
