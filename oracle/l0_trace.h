// ORACLE / TEST INFRASTRUCTURE ONLY.  Declarations of the trace hooks that the
// sed-instrumented copies of the reference files call (see oracle/Makefile);
// defined in oracle/l0_driver.cpp.
#ifndef PRT_ORACLE_L0_TRACE_H
#define PRT_ORACLE_L0_TRACE_H
#include <cstddef>
#include <cstdint>
extern "C" {
void l0_trace_node_test();
void l0_trace_query(uint32_t characterIndex, unsigned stepsLeft, const float * aabb6,
                    const uint16_t * mesh, size_t nMesh, const uint16_t * caps, size_t nCaps);
void l0_trace_tri_test();
void l0_trace_hit(long poly, const float * impulse3, const float * normal3);
}
#endif
