"""Scene ingestion and the headless scene host (include/prt_scene.h).

CPU part: the `.prt` parser against what the REFERENCE's own parser (scene_serialization.cpp compiled unmodified into
oracle/_ref/libprt_l0_scene.so) did with the same files -- committed golden dumps, the shipped level, and random files --
plus the OBJ reader and the error behaviour.  GPU part: a scene written to disk, loaded by the host and stepped through
prtc_* frame by frame, against the reference's PhysicsSystem (L0) fed with the generator's own numbers."""
import ctypes
import glob
import os
import re

import numpy as np
import pytest

import helpers as H
from oraclelib import L0World
from prototype2_b200 import capi, scene_host as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCENES = os.path.join(ROOT, "tests", "golden", "scenes")
L0_SCENE_LIB = os.path.join(ROOT, "oracle", "_ref", "libprt_l0_scene.so")
DT = np.float32(1.0 / 30.0)


def reference_parse(path):
    lib = ctypes.CDLL(L0_SCENE_LIB)
    lib.l0s_dump.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_size_t)]
    need = ctypes.c_size_t()
    lib.l0s_dump(os.fsencode(path), None, 0, ctypes.byref(need))
    buf = ctypes.create_string_buffer(need.value)
    lib.l0s_dump(os.fsencode(path), buf, need.value, None)
    return buf.value.decode()


needs_reference_parser = pytest.mark.skipif(not os.path.exists(L0_SCENE_LIB), reason="oracle/_ref/libprt_l0_scene.so not built")


# ------------------------------------------------------------------------------------------------------------
# library surface
# ------------------------------------------------------------------------------------------------------------
def test_every_declared_symbol_is_exported():
    text = open(os.path.join(ROOT, "include", "prt_scene.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = sorted(set(re.findall(r"\b(prts_[a-z_0-9]+)\s*\(", text)))
    lib = S.lib()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), "libprtscene.so does not export %s" % n
    assert sorted(S.SYMBOLS) == names, "scene_host.SYMBOLS out of date with include/prt_scene.h"


def test_struct_layouts():
    assert ctypes.sizeof(S.Transform) == 40                       # bit-compatible with the engine's Transform
    assert ctypes.sizeof(S.Entity) == 64 + 40 + 4 * 4 + 2 * 4 + 5 * 4 + 2 * 4
    assert ctypes.sizeof(S.Equipment) == 12 + 128 + 40


# ------------------------------------------------------------------------------------------------------------
# .prt parser
# ------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["grammar", "quirks", "crlf"])
def test_parser_matches_the_reference_parsers_golden_dump(name):
    got = S.SceneFile(path=os.path.join(SCENES, name + ".prt")).dump()
    assert got == open(os.path.join(SCENES, name + ".dump")).read()


def test_quirks_are_the_reference_parsers():
    f = S.SceneFile(path=os.path.join(SCENES, "quirks.prt"))
    e = f.entities()
    # the first component of an entity is read as type 0 whatever its number says (atoi("[3") == 0)
    assert e[0].name == b"first component is always read as a name" and e[0].collider_shape == S.SHAPE_NONE
    assert e[0].character_id == 0
    assert e[1].name == b"entity_1"                                # Entity[]: defaults
    assert (e[2].collider_shape, e[2].collider_index) == (S.SHAPE_CAPSULE, 1)     # second collider wins the tag, both exist
    assert f.models() == [("a/first.obj", False), ("a/second.obj", True)]         # de-duplicated by path, first flag kept
    assert e[3].model_id == 0 and e[3].collider_index == 0
    sun = f.sun()                                                   # the last Sun line wins, direction normalised
    assert sun.present == 1 and abs(sun.direction[0] - 2 ** -0.5) < 1e-6 and sun.direction[2] == 0.0


def test_parsed_values_of_the_grammar_fixture():
    f = S.SceneFile(path=os.path.join(SCENES, "grammar.prt"))
    e = f.entities()
    by_name = {x.name.decode(): x for x in e}
    assert len(e) == 11 and f.characters() == [4, 7, 8]         # equipment entities are created in between
    hero = by_name["hero"]
    assert list(hero.transform.rotation) == [0.0, 0.0, 0.0, 1.0]    # <2,0,0,0> normalised, stored x,y,z,w
    assert (hero.capsule_height, hero.capsule_radius, list(hero.capsule_offset)) == (0.5, 0.5, [0.0, 0.5, 0.0])
    assert hero.animated == 1 and hero.animation_id == 0
    assert list(by_name["second floor"].transform.rotation) == [0.0, 0.0, 0.0, 1.0]   # zero quaternion -> identity
    q = f.equipment()
    assert [(x.character_id, x.entity, x.bone) for x in q] == [(0, 5, b"hand.r"), (0, 6, b"hand.l"), (2, 9, b"head")]
    assert e[6].equipment_of == 0 and e[6].name == b"entity_6" and e[9].equipment_of == 2
    lights = f.lights()
    assert len(lights) == 1 and abs(lights[0].quadratic - 0.032) < 1e-7
    # a MODEL collider is registered with the transform parsed SO FAR: "ramp" has its transform after the collider
    dump = f.dump()
    ramp = [l for l in dump.splitlines() if l.startswith("modelcollider 1 ")][0]
    assert "pos 00000000 00000000 00000000 rot 00000000 00000000 00000000 3f800000" in ramp
    assert by_name["ramp"].transform.position[0] == 4.0


@needs_reference_parser
def test_parser_matches_the_reference_parser_on_the_shipped_level():
    shipped = "/root/reference/res/scenes/cavern.prt"
    golden = open(os.path.join(SCENES, "cavern.dump")).read()
    if os.path.exists(shipped):                                     # dev container: the file itself
        assert S.SceneFile(path=shipped).dump() == golden == reference_parse(shipped)
    f = golden.splitlines()
    assert sum(l.startswith("character ") for l in f) == 9 and sum(l.startswith("light ") for l in f) == 6


def _random_scene_text(rs):
    def num():
        v = float(rs.choice([rs.uniform(-40, 40), rs.integers(-9, 9), rs.uniform(-1, 1) * 1e-3]))
        return rs.choice(["%.7g", "%.3f", "%g", " %.5g", "%e"]) % v

    def vec3():
        return "<%s,%s,%s>" % (num(), num(), num())

    def quat():
        return "<%s,%s,%s,%s>" % (num(), num(), num(), num())

    lines = []
    n_entities = 0
    for _ in range(int(rs.integers(1, 14))):
        if rs.random() < 0.12:
            lines.append("Sun[direction%scolor%s]" % (vec3(), vec3()))
            continue
        if n_entities >= 60:
            break
        comps = ['0{<"e %d %s">}' % (n_entities, rs.choice(["a", "b b", "x_y", ""]))]
        kinds = list(rs.permutation([1, 2, 3, 5, 7, 1]))[: int(rs.integers(0, 6))]
        has_model = False
        is_capsule = False
        for k in kinds:
            if k == 1:
                comps.append("1{%s%s%s}" % (vec3(), quat(), vec3()))
            elif k == 2:
                comps.append('2{<"m/%s.obj"><%s>}' % (rs.choice(["a", "b", "c"]), rs.choice(["true", "false", "True", "F", "t"])))
                has_model = True
            elif k == 3:
                if has_model and rs.random() < 0.5:
                    comps.append('3{<"MODEL">}')
                else:
                    comps.append('3{<"CAPSULE"><%s><%s>%s}' % (num(), num(), vec3()))
                    is_capsule = True
            elif k == 5:
                comps.append("5{%s<%s><%s><%s>}" % (vec3(), num(), num(), num()))
            else:
                comps.append('7{<"skipped"><%s>}' % num())
        if is_capsule and rs.random() < 0.7:
            items = "".join('<<"bone%d"><"g/%s.dae">%s%s%s>' % (j, rs.choice(["s", "t"]), vec3(), quat(), vec3())
                            for j in range(int(rs.integers(0, 3))))
            comps.append("4{%s}" % items)
            n_entities += items.count("<<")
        n_entities += 1
        lines.append("Entity[%s]%s" % ("".join(comps), rs.choice(["", "", " tail", "\r"])))
    return "\n".join(lines) + ("\n" if rs.random() < 0.7 else "")


@needs_reference_parser
def test_parser_matches_the_reference_parser_on_random_files(tmp_path):
    rs = np.random.default_rng(20240611)
    for i in range(150):
        text = _random_scene_text(rs)
        p = tmp_path / ("s%03d.prt" % i)
        p.write_bytes(text.encode())
        assert S.SceneFile(path=str(p)).dump() == reference_parse(str(p)), text
        assert S.SceneFile(text=text).dump() == S.SceneFile(path=str(p)).dump()


def test_parse_errors_where_the_reference_asserts(tmp_path):
    cases = {
        'Entity[0{<"cut': S.PRTS_ERR_PARSE,                                    # text ends inside a string
        'Entity[0{<"a">}1{<1,2,3><1,0,0,0><1,1': S.PRTS_ERR_PARSE,             # ... inside a vec3
        'Entity[0{<"a">}1{<1,2><1,0,0,0><1,1,1>}]\n': S.PRTS_ERR_PARSE,        # vec3 with two numbers
        'Entity[0{<"a">}1{<1 ,2,3><1,0,0,0><1,1,1>}]\n': S.PRTS_ERR_PARSE,     # sscanf("%f,%f,%f") stops at " ,"
        'Entity[0{<"a">}3{<"SPHERE">}]\n': S.PRTS_ERR_PARSE,                   # unknown collider type
        'Entity[0{<"a">}3{<"MODEL">}]\n': S.PRTS_ERR_PARSE,                    # MODEL collider without a model: models[-1]
        'Mesh[0{<"a">}]\n': S.PRTS_ERR_PARSE,                                  # unknown token
        '\nEntity[0{<"a">}]\n': S.PRTS_ERR_PARSE,                              # blank line: the token is "\nEntity"
    }
    for text, code in cases.items():
        with pytest.raises(S.SceneError) as e:
            S.SceneFile(text=text)
        assert e.value.code == code, text
    with pytest.raises(S.SceneError) as e:
        S.SceneFile(text="".join('Entity[0{<"e">}]\n' for _ in range(S.MAX_ENTITIES + 1)))
    assert e.value.code == S.PRTS_ERR_LIMIT
    assert len(S.SceneFile(text="".join('Entity[0{<"e">}]\n' for _ in range(S.MAX_ENTITIES))).entities()) == S.MAX_ENTITIES
    with pytest.raises(S.SceneError) as e:
        S.SceneFile(path=str(tmp_path / "missing.prt"))
    assert e.value.code == S.PRTS_ERR_IO
    assert S.SceneFile(text="").entities() == []


# ------------------------------------------------------------------------------------------------------------
# OBJ geometry
# ------------------------------------------------------------------------------------------------------------
OBJ = """# comment
mtllib x.mtl
o empty_group_is_skipped
o quad
v 0 0 0
v 1 0 0
v 1 0 1
v 0 0 1
vn 0 1 0
f 1//1 2//1 3//1 4//1
o two_materials
v 0 1 0
v 1 1 0
v 1 1 1
usemtl red
f 5/1/1 6/2/1 7/3/1
usemtl blue
f -3 -1 -2
g five_gon
v 2 0 0
v 3 0 0
v 3 0 1
v 2.5 0 2
v 2 0 1
f 8 9 10 11 12
"""


def test_obj_reader_groups_fans_and_index_forms():
    m = S.ObjModel(text=OBJ)
    assert m.xyz.shape == (12, 3) and np.array_equal(m.xyz[10], [2.5, 0, 2])
    assert m.mesh_start.tolist() == [0, 6, 9, 12] and m.mesh_count.tolist() == [6, 3, 3, 9]
    assert m.indices.tolist() == [0, 1, 2, 0, 2, 3,   4, 5, 6,   4, 6, 5,   7, 8, 9, 7, 9, 10, 7, 10, 11]
    for bad in ("v 1 2\n", "v 0 0 0\nf 1 2 3\n", "v 0 0 0\nv 1 0 0\nf 1 2\n"):
        with pytest.raises(S.SceneError) as e:
            S.ObjModel(text=bad)
        assert e.value.code == S.PRTS_ERR_PARSE
    with pytest.raises(S.SceneError) as e:
        S.ObjModel(path="/nonexistent/model.fbx")
    assert e.value.code == S.PRTS_ERR_IO


def test_obj_reader_reads_the_shipped_level_like_the_c1_fixture():
    path = "/root/reference/res/models/level1/level1.obj"
    if not os.path.exists(path):
        pytest.skip("reference resources not present")
    g = np.load(os.path.join(ROOT, "tests", "golden", "c1_level.npz"))
    m = S.ObjModel(path=path)
    assert len(m.mesh_start) == 172 and m.mesh_count.sum() == 3 * 3007
    assert H.same(m.xyz[m.indices], g["xyz"]) and np.array_equal(m.mesh_start, g["mesh_first"]) and np.array_equal(m.mesh_count, g["mesh_count"])


# ------------------------------------------------------------------------------------------------------------
# headless scene host on the GPU
# ------------------------------------------------------------------------------------------------------------
def _f(v):
    return "%.9g" % float(np.float32(v))


def _write_level(tmp_path, n_chars, seed):
    """A two-model level (rolling terrain, one mesh per 2x2 cells, under a scaled + yawed transform; a tilted slab) and
    n_chars capsule characters, written as OBJ + .prt.  Returns what was written, as float32 numbers."""
    rs = np.random.default_rng(seed)
    n = 16
    xyz, mf, mc = H.scenes.terrain(n, 2, 2)                       # index-expanded triangles, mesh by mesh
    os.makedirs(tmp_path / "models" / "lvl", exist_ok=True)
    with open(tmp_path / "models" / "lvl" / "terrain.obj", "w") as f:
        for v in xyz:
            f.write("v %s %s %s\n" % (_f(v[0]), _f(v[1]), _f(v[2])))
        for k, (a, c) in enumerate(zip(mf, mc)):
            f.write("o block%d\n" % k)
            for t in range(int(a), int(a + c), 3):
                f.write("f %d %d %d\n" % (t + 1, t + 2, t + 3))
    slab = np.array([[0, 0, 0], [0, 0, 4], [4, 0, 4], [0, 0, 0], [4, 0, 4], [4, 0, 0]], np.float32)
    with open(tmp_path / "models" / "lvl" / "slab.obj", "w") as f:
        for v in slab:
            f.write("v %s %s %s\n" % (_f(v[0]), _f(v[1]), _f(v[2])))
        f.write("o slab\nf 1 2 3\nf 4 5 6\n")
    terrain_t = np.float32([0.0, 0.0, 0.0, 0.9914449, 0.0, 0.1305262, 0.0, 1.5, 1.0, 1.5])      # pos, quat wxyz, scale
    slab_t = np.float32([6.0, 1.2, 6.0, 0.9961947, 0.0871557, 0.0, 0.0, 1.0, 1.0, 1.0])
    lines = ["Sun[direction<0,-1,-2>color<1,1,1>]",
             'Entity[0{<"terrain">}1{<%s,%s,%s><%s,%s,%s,%s><%s,%s,%s>}2{<"lvl/terrain.obj"><false>}3{<"MODEL">}]' % tuple(map(_f, terrain_t)),
             'Entity[0{<"lamp">}1{<0,9,0><1,0,0,0><1,1,1>}5{<1,1,1><1><0><0.1>}]',
             'Entity[0{<"slab">}1{<%s,%s,%s><%s,%s,%s,%s><%s,%s,%s>}2{<"lvl/slab.obj"><false>}3{<"MODEL">}]' % tuple(map(_f, slab_t))]
    chars = []
    for i in range(n_chars):
        p = np.float32([rs.uniform(5, 15), rs.uniform(1.0, 2.5), rs.uniform(5, 15)])
        cap = np.float32([rs.choice([0.5, 0.75]), rs.choice([0.5, 0.35])])
        off = np.float32([0.0, cap[1], 0.0])
        chars.append((p, cap, off))
        equipment = '<<"hand.r"><"gear/sword.dae"><0,0,0><1,0,0,0><1,1,1>>' if i % 2 == 0 else ""
        lines.append('Entity[0{<"c%d">}1{<%s,%s,%s><1,0,0,0><0.25,0.25,0.25>}2{<"people/hero.fbx"><true>}3{<"CAPSULE"><%s><%s><%s,%s,%s>}4{%s}]'
                     % (i, _f(p[0]), _f(p[1]), _f(p[2]), _f(cap[0]), _f(cap[1]), _f(off[0]), _f(off[1]), _f(off[2]), equipment))
    (tmp_path / "level.prt").write_text("\n".join(lines) + "\n")
    return {"terrain": (xyz, mf, mc, terrain_t), "slab": (slab, np.uint32([0]), np.uint32([6]), slab_t), "chars": chars}


def _reference_world(level, variant="plain"):
    """the reference's PhysicsSystem driven the way SceneSerialization + CharacterSystem drive it"""
    orc = L0World(variant)
    for key in ("terrain", "slab"):
        xyz, mf, mc, t = level[key]
        orc.add_model(xyz, mf, mc, t)
    # file order: terrain, slab, then the capsules (colliders interleave only with non-collider entities here)
    for p, cap, off in level["chars"]:
        orc.add_capsule(float(cap[0]), float(cap[1]), off)
    orc.set_state(pos=np.array([c[0] for c in level["chars"]], np.float32))
    return orc


@pytest.mark.gpu
def test_world_steps_a_loaded_level_like_the_reference(tmp_path):
    if not L0World.available("plain"):
        pytest.skip("oracle/_ref not built")
    n_chars = 12
    level = _write_level(tmp_path, n_chars, seed=5)
    scene = S.SceneFile(path=str(tmp_path / "level.prt"))
    world = S.World(scene, str(tmp_path / "models") + "/")
    assert world.n_characters == n_chars and world.n_entities == 3 + n_chars + n_chars // 2
    ents = world.character_entities
    assert np.array_equal(world.transforms["position"][ents], np.array([c[0] for c in level["chars"]]))
    orc = _reference_world(level)
    rs = np.random.default_rng(1)
    move = np.zeros((n_chars, 3), np.float32)
    contacts = 0
    for frame in range(90):
        if frame % 15 == 0:                                       # the caller plays CharacterSystem::updateCharacters
            move = np.zeros((n_chars, 3), np.float32)
            move[:, [0, 2]] = rs.uniform(-0.2, 0.2, (n_chars, 2)).astype(np.float32)
            move[rs.random(n_chars) < 0.3] = 0.0
        world.physics["movement"][:] = move
        if frame % 30 == 10:                                      # jump (character_system.cpp:92)
            world.physics["velocity"][:, 1] = np.float32(0.35)
        s = {"pos": world.transforms["position"][ents].copy(), "vel": world.physics["velocity"].copy(),
             "gnormal": world.physics["ground_normal"].copy(), "grounded": world.physics["grounded"].astype(np.uint8)}
        orc.set_state(pos=s["pos"], vel=s["vel"], move=move, gnormal=s["gnormal"], grounded=s["grounded"])
        orc.step(DT)
        world.update_physics(DT)
        so = orc.get_state()
        assert H.same(world.transforms["position"][ents], so["pos"]), "frame %d positions" % frame
        assert H.same(world.physics["velocity"], so["vel"]), "frame %d velocities" % frame
        assert np.array_equal(world.physics["grounded"].astype(np.uint8), so["grounded"]), "frame %d grounded" % frame
        g = so["grounded"].astype(bool)
        assert H.same(world.physics["ground_normal"][g], so["gnormal"][g]), "frame %d ground normals" % frame
        contacts += int(g.sum())
    assert contacts > 100                                          # the characters really stood on the level
    # non-character entities are untouched; scale / rotation of the characters survive the gather / scatter
    assert np.array_equal(world.transforms["position"][1], np.float32([0, 9, 0]))
    assert np.array_equal(world.transforms["scale"][ents], np.full((n_chars, 3), 0.25, np.float32))
    world.close()


@pytest.mark.gpu
def test_world_runs_the_shipped_level_from_files(tmp_path):
    """C1 through the scene host: the shipped level's collider geometry (tests/golden/c1_level.npz) written back out as
    an OBJ, the map and player lines of res/scenes/cavern.prt:8-9, 300 scripted frames -- against the reference's own
    output stored in the fixture (tests/golden/make_c1.py)."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "c1_level.npz"))
    os.makedirs(tmp_path / "models" / "level1")
    with open(tmp_path / "models" / "level1" / "level1.obj", "w") as f:
        for v in g["xyz"]:
            f.write("v %s %s %s\n" % (_f(v[0]), _f(v[1]), _f(v[2])))
        for k, (a, c) in enumerate(zip(g["mesh_first"], g["mesh_count"])):
            f.write("o part%d\n" % k)
            for t in range(int(a), int(a + c), 3):
                f.write("f %d %d %d\n" % (t + 1, t + 2, t + 3))
    (tmp_path / "cavern.prt").write_text(
        'Sun[direction<0.0,-1.0,-2.0>color<1.0,1.0,1.0>]\n'
        'Entity[0{<"map">}1{<0.0,0.0,0.0><0.7071068,0,0.7071068,0><1.5,1.5,1.5>}2{<"level1/level1.obj"><false>}3{<"MODEL">}]\n'
        'Entity[0{<"player">}1{<5.6,8.0,5.0><1.0,0.0,0.0,0.0><0.25,0.25,0.25>}2{<"flygy/flygy.fbx"><true>}'
        '3{<"CAPSULE"><0.5><0.5><0.0,0.5,0.0>}4{<<"hand1.r"><"swd/swd.dae"><-0.1,0.1,0.0><0.7071068,0,0,0.7071068><0.3,0.3,0.3>>}]\n')
    world = S.World(S.SceneFile(path=str(tmp_path / "cavern.prt")), str(tmp_path / "models") + "/")
    player = int(world.character_entities[0])
    assert player == 1 and world.n_entities == 3
    for frame in range(int(g["frames"])):
        if g["jump_script"][frame]:
            world.physics["velocity"][0, 1] = np.float32(0.35)
        world.physics["movement"][0] = g["move_script"][frame]
        world.update_physics(float(g["dt"]))
        assert H.same(world.transforms["position"][player], g["int_pos"][frame][0]), "frame %d position" % frame
        assert H.same(world.physics["velocity"][0], g["int_vel"][frame][0]), "frame %d velocity" % frame
        assert int(world.physics["grounded"][0]) == int(g["int_grounded"][frame][0]), "frame %d grounded" % frame
    world.close()


@pytest.mark.gpu
def test_world_moves_a_model_collider_and_raycasts(tmp_path):
    if not L0World.available("plain"):
        pytest.skip("oracle/_ref not built")
    level = _write_level(tmp_path, 6, seed=9)
    scene = S.SceneFile(path=str(tmp_path / "level.prt"))
    world = S.World(scene, str(tmp_path / "models") + "/")
    ents = world.character_entities
    for _ in range(5):
        world.update_physics(DT)
    # lift the slab (entity 2): Scene::updateColliders re-registers it before the characters move (scene.cpp:190-211)
    new_t = np.float32([6.0, 2.0, 5.0, 0.9961947, 0.0871557, 0.0, 0.0, 1.0, 1.0, 1.5])
    world.move_entity(2, new_t[0:3], [new_t[4], new_t[5], new_t[6], new_t[3]], new_t[7:10])
    pos0 = world.transforms["position"][ents].copy()
    state = {k: world.physics[k].copy() for k in ("velocity", "ground_normal", "grounded")}
    world.update_physics(DT)
    # reference world built directly in the moved configuration (the tree differs in shape, so compare the rays and
    # the characters that have one candidate sequence: all of them must at least stay finite and near)
    moved = dict(level)
    moved["slab"] = (level["slab"][0], level["slab"][1], level["slab"][2], new_t)
    orc = _reference_world(moved)
    orc.set_state(pos=pos0, vel=state["velocity"], gnormal=state["ground_normal"], grounded=state["grounded"].astype(np.uint8))
    orc.step(DT)
    same = np.all(H.bits(world.transforms["position"][ents]) == H.bits(orc.get_state()["pos"]), axis=1)
    assert same.mean() >= 0.8
    rs = np.random.default_rng(3)
    n = 64
    origins = np.stack([rs.uniform(2, 20, n), rs.uniform(4, 8, n), rs.uniform(2, 20, n)], 1).astype(np.float32)
    origins[: n // 4] = np.float32([8.0, 6.0, 8.0])               # some straight over the lifted slab
    dirs = np.tile(np.float32([0.05, -1.0, 0.02]), (n, 1))
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True).astype(np.float32)
    maxd = np.full(n, 12.0, np.float32)
    hit, mask = world.raycast(origins, dirs, maxd)
    assert mask.sum() > n // 2
    for i in range(n):
        ok, h = orc.raycast(origins[i], dirs[i], 12.0)
        assert ok == bool(mask[i]), i
        if ok:
            assert H.same(h, hit[i]), i
    world.close()


@pytest.mark.gpu
def test_world_creation_errors(tmp_path):
    _write_level(tmp_path, 2, seed=1)
    scene = S.SceneFile(path=str(tmp_path / "level.prt"))
    with pytest.raises(S.SceneError) as e:
        S.World(scene, str(tmp_path / "nowhere") + "/")
    assert e.value.code == S.PRTS_ERR_IO
    bad = S.SceneFile(text='Entity[0{<"c">}4{}3{<"CAPSULE"><0.5><0.5><0,0.5,0>}]\n')
    with pytest.raises(S.SceneError) as e:
        S.World(bad, str(tmp_path / "models") + "/")
    assert e.value.code == S.PRTS_ERR_INVALID
    fbx = S.SceneFile(text='Entity[0{<"m">}2{<"people/hero.fbx"><false>}3{<"MODEL">}]\n')
    with pytest.raises(S.SceneError) as e:
        S.World(fbx, str(tmp_path / "models") + "/")
    assert e.value.code == S.PRTS_ERR_IO and "unsupported" in str(e.value)
