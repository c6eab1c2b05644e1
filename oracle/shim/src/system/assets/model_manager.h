// ORACLE ONLY: empty shadow (the real header needs assimp)
#ifndef PRT_ORACLE_SHADOW_MODEL_MANAGER_H
#define PRT_ORACLE_SHADOW_MODEL_MANAGER_H
#endif
