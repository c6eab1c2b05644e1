"""ctypes binding of include/prt_scene.h (libprtscene.so): `.prt` scene files, OBJ collider geometry and the headless
scene host that drives the collision path frame by frame (reference src/game/scene/scene_serialization.cpp,
src/game/scene/scene.cpp:185-211, src/game/system/character/character_system.cpp:55-78)."""
import ctypes
import os

import numpy as np

from . import capi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libprtscene.so")

PRTS_OK, PRTS_ERR_IO, PRTS_ERR_PARSE, PRTS_ERR_INVALID, PRTS_ERR_LIMIT, PRTS_ERR_PHYSICS = range(6)
SHAPE_NONE, SHAPE_MODEL, SHAPE_MESH, SHAPE_CAPSULE = range(4)
MAX_ENTITIES, NAME_LEN, PATH_LEN = 100, 64, 128

# every symbol include/prt_scene.h declares (tests check the library exports exactly these)
SYMBOLS = (
    "prts_last_error", "prts_scene_parse", "prts_scene_load", "prts_scene_free", "prts_entity_count", "prts_get_entity",
    "prts_model_count", "prts_get_model", "prts_character_count", "prts_get_character", "prts_equipment_count",
    "prts_get_equipment", "prts_light_count", "prts_get_light", "prts_get_sun", "prts_scene_dump",
    "prts_model_parse_obj", "prts_model_load", "prts_model_free", "prts_model_info", "prts_model_data",
    "prts_world_create", "prts_world_destroy", "prts_world_ctx", "prts_world_entity_count", "prts_world_character_count",
    "prts_world_transforms", "prts_world_physics", "prts_world_character_entities", "prts_world_move_entity",
    "prts_world_update_physics", "prts_world_raycast",
)


class Transform(ctypes.Structure):
    _fields_ = [("position", ctypes.c_float * 3), ("rotation", ctypes.c_float * 4), ("scale", ctypes.c_float * 3)]


class Entity(ctypes.Structure):
    _fields_ = [("name", ctypes.c_char * NAME_LEN), ("transform", Transform), ("model_id", ctypes.c_int32),
                ("animated", ctypes.c_int32), ("animation_id", ctypes.c_int32), ("character_id", ctypes.c_int32),
                ("collider_shape", ctypes.c_uint32), ("collider_index", ctypes.c_uint32),
                ("capsule_height", ctypes.c_float), ("capsule_radius", ctypes.c_float), ("capsule_offset", ctypes.c_float * 3),
                ("light_index", ctypes.c_int32), ("equipment_of", ctypes.c_int32)]


class Equipment(ctypes.Structure):
    _fields_ = [("character_id", ctypes.c_int32), ("entity", ctypes.c_int32), ("model_id", ctypes.c_int32),
                ("bone", ctypes.c_char * PATH_LEN), ("offset", Transform)]


class PointLight(ctypes.Structure):
    _fields_ = [("entity", ctypes.c_int32), ("color", ctypes.c_float * 3), ("constant", ctypes.c_float),
                ("linear", ctypes.c_float), ("quadratic", ctypes.c_float)]


class Sun(ctypes.Structure):
    _fields_ = [("present", ctypes.c_int32), ("direction", ctypes.c_float * 3), ("color", ctypes.c_float * 3)]


class SceneError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("prts error %d: %s" % (code, msg))
        self.code = code


_lib = None


def lib():
    """Loads libprtscene.so (which pulls in libprtc.so); raises when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("libprtscene.so is missing: build it with `make -C prototype2_b200/csrc all`")
        L = ctypes.CDLL(LIB_PATH)
        L.prts_last_error.restype = ctypes.c_char_p
        for name in ("prts_entity_count", "prts_model_count", "prts_character_count", "prts_equipment_count", "prts_light_count",
                     "prts_world_entity_count", "prts_world_character_count"):
            getattr(L, name).restype = ctypes.c_uint32
            getattr(L, name).argtypes = [ctypes.c_void_p]
        L.prts_scene_parse.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_void_p)]
        L.prts_scene_load.argtypes = [ctypes.c_char_p, ctypes.POINTER(ctypes.c_void_p)]
        L.prts_scene_free.argtypes = [ctypes.c_void_p]
        L.prts_scene_free.restype = None
        L.prts_get_entity.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.POINTER(Entity)]
        L.prts_get_model.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_char_p, ctypes.POINTER(ctypes.c_int32)]
        L.prts_get_character.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.POINTER(ctypes.c_int32)]
        L.prts_get_equipment.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.POINTER(Equipment)]
        L.prts_get_light.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.POINTER(PointLight)]
        L.prts_get_sun.argtypes = [ctypes.c_void_p, ctypes.POINTER(Sun)]
        L.prts_scene_dump.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_size_t)]
        L.prts_model_parse_obj.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_void_p)]
        L.prts_model_load.argtypes = [ctypes.c_char_p, ctypes.POINTER(ctypes.c_void_p)]
        L.prts_model_free.argtypes = [ctypes.c_void_p]
        L.prts_model_free.restype = None
        u32p = ctypes.POINTER(ctypes.c_uint32)
        L.prts_model_info.argtypes = [ctypes.c_void_p, u32p, u32p, u32p]
        L.prts_model_data.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.POINTER(ctypes.c_float)), ctypes.POINTER(u32p),
                                      ctypes.POINTER(u32p), ctypes.POINTER(u32p)]
        L.prts_world_create.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.POINTER(capi.Config), ctypes.POINTER(ctypes.c_void_p)]
        L.prts_world_destroy.argtypes = [ctypes.c_void_p]
        L.prts_world_destroy.restype = None
        L.prts_world_ctx.argtypes = [ctypes.c_void_p]
        L.prts_world_ctx.restype = ctypes.c_void_p
        L.prts_world_transforms.argtypes = [ctypes.c_void_p]
        L.prts_world_transforms.restype = ctypes.c_void_p
        L.prts_world_physics.argtypes = [ctypes.c_void_p]
        L.prts_world_physics.restype = ctypes.c_void_p
        L.prts_world_character_entities.argtypes = [ctypes.c_void_p]
        L.prts_world_character_entities.restype = ctypes.c_void_p
        L.prts_world_move_entity.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.POINTER(Transform)]
        L.prts_world_update_physics.argtypes = [ctypes.c_void_p, ctypes.c_float]
        L.prts_world_raycast.argtypes = [ctypes.c_void_p, ctypes.c_uint32] + [ctypes.c_void_p] * 5
        _lib = L
    return _lib


def _check(code):
    if code != PRTS_OK:
        raise SceneError(code, lib().prts_last_error().decode())


def _view(ptr, count, dtype):
    if count == 0:
        return np.zeros(0, dtype)
    buf = (ctypes.c_char * (count * np.dtype(dtype).itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype, count=count)


class SceneFile:
    """A parsed `.prt` file (SceneSerialization::loadScene)."""

    def __init__(self, path=None, text=None):
        self._h = ctypes.c_void_p()
        if path is not None:
            _check(lib().prts_scene_load(os.fsencode(path), ctypes.byref(self._h)))
        else:
            data = text.encode() if isinstance(text, str) else bytes(text)
            _check(lib().prts_scene_parse(data, len(data), ctypes.byref(self._h)))

    def close(self):
        if self._h:
            lib().prts_scene_free(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    def entities(self):
        out = []
        for i in range(lib().prts_entity_count(self._h)):
            e = Entity()
            _check(lib().prts_get_entity(self._h, i, ctypes.byref(e)))
            out.append(e)
        return out

    def models(self):
        out = []
        for i in range(lib().prts_model_count(self._h)):
            path = ctypes.create_string_buffer(PATH_LEN)
            animated = ctypes.c_int32()
            _check(lib().prts_get_model(self._h, i, path, ctypes.byref(animated)))
            out.append((path.value.decode(), bool(animated.value)))
        return out

    def characters(self):
        out = []
        for i in range(lib().prts_character_count(self._h)):
            e = ctypes.c_int32()
            _check(lib().prts_get_character(self._h, i, ctypes.byref(e)))
            out.append(e.value)
        return out

    def equipment(self):
        out = []
        for i in range(lib().prts_equipment_count(self._h)):
            q = Equipment()
            _check(lib().prts_get_equipment(self._h, i, ctypes.byref(q)))
            out.append(q)
        return out

    def lights(self):
        out = []
        for i in range(lib().prts_light_count(self._h)):
            l = PointLight()
            _check(lib().prts_get_light(self._h, i, ctypes.byref(l)))
            out.append(l)
        return out

    def sun(self):
        s = Sun()
        _check(lib().prts_get_sun(self._h, ctypes.byref(s)))
        return s

    def dump(self):
        need = ctypes.c_size_t()
        _check(lib().prts_scene_dump(self._h, None, 0, ctypes.byref(need)))
        buf = ctypes.create_string_buffer(need.value)
        _check(lib().prts_scene_dump(self._h, buf, need.value, None))
        return buf.value.decode()


class ObjModel:
    """Collider geometry of one OBJ file: positions, index buffer, (startIndex, numIndices) per mesh."""

    def __init__(self, path=None, text=None):
        h = ctypes.c_void_p()
        if path is not None:
            _check(lib().prts_model_load(os.fsencode(path), ctypes.byref(h)))
        else:
            data = text.encode() if isinstance(text, str) else bytes(text)
            _check(lib().prts_model_parse_obj(data, len(data), ctypes.byref(h)))
        nv, ni, nm = ctypes.c_uint32(), ctypes.c_uint32(), ctypes.c_uint32()
        _check(lib().prts_model_info(h, ctypes.byref(nv), ctypes.byref(ni), ctypes.byref(nm)))
        xyz = ctypes.POINTER(ctypes.c_float)()
        u32p = ctypes.POINTER(ctypes.c_uint32)
        idx, ms, mc = u32p(), u32p(), u32p()
        _check(lib().prts_model_data(h, ctypes.byref(xyz), ctypes.byref(idx), ctypes.byref(ms), ctypes.byref(mc)))

        def copy(p, n, dtype):
            return np.ctypeslib.as_array(p, shape=(n,)).astype(dtype, copy=True) if n else np.zeros(0, dtype)

        self.xyz = copy(xyz, 3 * nv.value, np.float32).reshape(-1, 3)
        self.indices = copy(idx, ni.value, np.uint32)
        self.mesh_start = copy(ms, nm.value, np.uint32)
        self.mesh_count = copy(mc, nm.value, np.uint32)
        lib().prts_model_free(h)


class World:
    """Headless Scene: colliders registered in file order, `update_physics(dt)` = Scene::updatePhysics."""

    def __init__(self, scene, model_dir, config=None):
        self._h = ctypes.c_void_p()
        cfg = ctypes.byref(config) if config is not None else None
        _check(lib().prts_world_create(scene.handle, os.fsencode(model_dir) if model_dir is not None else None, cfg,
                                       ctypes.byref(self._h)))
        L = lib()
        self.n_entities = L.prts_world_entity_count(self._h)
        self.n_characters = L.prts_world_character_count(self._h)
        # live views of the world's own tables
        self.transforms = _view(L.prts_world_transforms(self._h), self.n_entities, capi.TRANSFORM_DTYPE)
        self.physics = _view(L.prts_world_physics(self._h), self.n_characters, capi.PHYSICS_DTYPE)
        self.character_entities = _view(L.prts_world_character_entities(self._h), self.n_characters, np.int32)

    def close(self):
        if self._h:
            self.transforms = self.physics = self.character_entities = None
            lib().prts_world_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def move_entity(self, entity, position, rotation_xyzw, scale):
        t = Transform((ctypes.c_float * 3)(*position), (ctypes.c_float * 4)(*rotation_xyzw), (ctypes.c_float * 3)(*scale))
        _check(lib().prts_world_move_entity(self._h, entity, ctypes.byref(t)))

    def update_physics(self, dt):
        _check(lib().prts_world_update_physics(self._h, ctypes.c_float(dt)))

    def raycast(self, origins, directions, max_distances):
        o = np.ascontiguousarray(origins, np.float32).reshape(-1, 3)
        d = np.ascontiguousarray(directions, np.float32).reshape(-1, 3)
        m = np.ascontiguousarray(max_distances, np.float32).reshape(-1)
        hit = np.zeros((len(o), 3), np.float32)
        mask = np.zeros(len(o), np.uint8)
        _check(lib().prts_world_raycast(self._h, len(o), o.ctypes.data, d.ctypes.data, m.ctypes.data, hit.ctypes.data, mask.ctypes.data))
        return hit, mask.astype(bool)
