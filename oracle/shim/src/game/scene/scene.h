// ORACLE / TEST INFRASTRUCTURE ONLY.  Shadow of the engine's Scene: just the
// accessors PhysicsSystem::updateCharacters uses (reference
// physics_system.cpp:249-263,335-337).
#ifndef PRT_ORACLE_SHADOW_SCENE_H
#define PRT_ORACLE_SHADOW_SCENE_H
#include "src/game/system/character/character.h"

class CharacterSystem {
public:
    prt::vector<Character> chars;
    Character & getCharacter(size_t i) { return chars[i]; }
    size_t getNumberOfCharacters() const { return chars.size(); }
};

class Scene {
public:
    CharacterSystem cs;
    CharacterSystem & getCharacterSystem() { return cs; }
};
#endif
