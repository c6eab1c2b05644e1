"""ctypes binding of include/prtc.h (the C-ABI of libprtc.so).  Mirrors the header one to one; the small
`PhysicsContext` class on top keeps the argument meaning of the reference's PhysicsSystem methods
(reference src/game/system/physics/physics_system.h:24-77)."""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, os.environ.get("PRTC_LIB", "libprtc.so"))   # PRTC_LIB: A/B builds of the same library

PRTC_OK, PRTC_ERR_INVALID, PRTC_ERR_CUDA, PRTC_ERR_CAPACITY, PRTC_ERR_ALLOC = range(5)
ABS_INT, ABS_FLOAT = 0, 1
ORDER_LITERAL, ORDER_CANONICAL, ORDER_BY_ID = 0, 1, 2
DYN_OFF, DYN_JACOBI = 0, 2
OPT_FRAME_CACHE, OPT_QUICK_REJECT, OPT_PERSISTENT, OPT_CROSS_FRAME_CACHE, OPT_WARP_KERNEL, OPT_STREAMED_COPY, OPT_TOP_STAGING = 1, 2, 3, 4, 5, 6, 7
OPT_POOL_KERNEL, OPT_TIGHT_REJECT, OPT_CHAIN_FRAMES = 8, 9, 10

# every symbol include/prtc.h declares (tests check the library exports exactly these)
SYMBOLS = (
    "prtc_default_config", "prtc_last_error", "prtc_device_count", "prtc_create", "prtc_destroy",
    "prtc_add_model", "prtc_add_capsule", "prtc_update_capsule", "prtc_commit", "prtc_update_models", "prtc_remove_model",
    "prtc_step_characters", "prtc_step_characters_device", "prtc_step_characters_traced", "prtc_raycast", "prtc_sweep_ellipsoids",
    "prtc_set_stream", "prtc_synchronize", "prtc_get_counters", "prtc_set_option", "prtc_dyn_configure", "prtc_dyn_prepare", "prtc_dyn_buffer", "prtc_get_tree_info", "prtc_get_leaf_order",
    "prtc_get_world_vertices", "prtc_query_box", "prtc_host_build_tree", "prtc_host_top_table", "prtc_host_balanced_tree", "prtc_selftest_math",
    "prtc_resident_begin", "prtc_resident_io", "prtc_resident_step", "prtc_resident_edit", "prtc_resident_fetch", "prtc_resident_end",
    "prtc_multi_create", "prtc_multi_destroy", "prtc_multi_device_count", "prtc_multi_context", "prtc_multi_add_model", "prtc_multi_add_capsule",
    "prtc_multi_update_capsule", "prtc_multi_remove_model", "prtc_multi_update_models", "prtc_multi_commit", "prtc_multi_set_option",
    "prtc_multi_raycast", "prtc_multi_get_counters", "prtc_multi_step_characters", "prtc_multi_resident_begin", "prtc_multi_resident_io",
    "prtc_multi_resident_step", "prtc_multi_resident_fetch", "prtc_multi_resident_end", "prtc_get_commit_stats", "prtc_debug_timeline",
    "prtc_dyn_ipc_export", "prtc_dyn_ipc_import",
)


class Config(ctypes.Structure):
    _fields_ = [("gravity", ctypes.c_float), ("leaf_margin", ctypes.c_float), ("ground_dot", ctypes.c_float),
                ("friction", ctypes.c_float), ("stick", ctypes.c_float), ("substeps", ctypes.c_uint32),
                ("abs_mode", ctypes.c_uint32), ("order_mode", ctypes.c_uint32), ("dyn_mode", ctypes.c_uint32),
                ("device", ctypes.c_int32)]


class Counters(ctypes.Structure):
    _fields_ = [(k, ctypes.c_uint64) for k in ("substeps", "tri_tests", "node_tests", "contacts", "cc_tests", "exact_tests", "launches")]


_u32p = ctypes.POINTER(ctypes.c_uint32)
_i32p = ctypes.POINTER(ctypes.c_int32)
_f32p = ctypes.POINTER(ctypes.c_float)


class TraceStruct(ctypes.Structure):
    _fields_ = [("cap_queries", ctypes.c_uint32), ("cap_mesh_ids", ctypes.c_uint32), ("cap_caps_ids", ctypes.c_uint32),
                ("cap_hits", ctypes.c_uint32), ("n_queries", ctypes.c_uint32), ("n_mesh_ids", ctypes.c_uint32),
                ("n_caps_ids", ctypes.c_uint32), ("n_hits", ctypes.c_uint32),
                ("q_char", _u32p), ("q_sub", _u32p), ("q_aabb", _f32p), ("q_mesh_off", _u32p), ("q_caps_off", _u32p),
                ("mesh_ids", _u32p), ("caps_ids", _u32p), ("q_tri_tests", _u32p), ("h_query", _u32p),
                ("h_poly", _i32p), ("h_impulse", _f32p), ("h_normal", _f32p), ("h_pos", _f32p)]


# numpy views of the two AoS records of the boundary (bit-compatible with prtc_transform / prtc_char_physics)
TRANSFORM_DTYPE = np.dtype([("position", np.float32, 3), ("rotation", np.float32, 4), ("scale", np.float32, 3)])
PHYSICS_DTYPE = np.dtype([("velocity", np.float32, 3), ("movement", np.float32, 3), ("ground_normal", np.float32, 3),
                          ("collider", np.uint32), ("grounded", np.uint32)])
assert TRANSFORM_DTYPE.itemsize == 40 and PHYSICS_DTYPE.itemsize == 44


class PrtcError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("prtc error %d: %s" % (code, msg))
        self.code = code


def load_library():
    """Loads libprtc.so; raises (never falls back) when it has not been built."""
    if not os.path.exists(LIB_PATH):
        raise ImportError("libprtc.so is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "or `make -C prototype2_b200/csrc` (there is no CPU fallback)")
    lib = ctypes.CDLL(LIB_PATH)
    lib.prtc_last_error.restype = ctypes.c_char_p
    return lib


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = load_library()
    return _lib


def device_count():
    return int(lib().prtc_device_count())


def _vp(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def make_transforms(pos, rot_xyzw=None, scale=None):
    n = len(pos)
    t = np.zeros(n, TRANSFORM_DTYPE)
    t["position"] = pos
    if rot_xyzw is None:
        t["rotation"][:, 3] = 1.0
    else:
        t["rotation"] = rot_xyzw
    t["scale"] = 1.0 if scale is None else scale
    return t


def make_physics(vel, movement=None, ground_normal=None, grounded=None, collider=None):
    n = len(vel)
    p = np.zeros(n, PHYSICS_DTYPE)
    p["velocity"] = vel
    if movement is not None:
        p["movement"] = movement
    if ground_normal is not None:
        p["ground_normal"] = ground_normal
    if grounded is not None:
        p["grounded"] = grounded
    p["collider"] = np.arange(n, dtype=np.uint32) if collider is None else collider
    return p


class PhysicsContext:
    """One prtc_ctx.  Method names follow the reference's PhysicsSystem."""

    def __init__(self, device=0, abs_mode=ABS_INT, order_mode=ORDER_CANONICAL, dyn_mode=DYN_OFF, substeps=4, **kw):
        self.lib = lib()
        cfg = Config()
        self.lib.prtc_default_config(ctypes.byref(cfg))
        cfg.device, cfg.abs_mode, cfg.order_mode, cfg.dyn_mode, cfg.substeps = device, abs_mode, order_mode, dyn_mode, substeps
        for k, v in kw.items():
            setattr(cfg, k, v)
        self.cfg = cfg
        self.h = ctypes.c_void_p()
        self._check(self.lib.prtc_create(ctypes.byref(cfg), ctypes.byref(self.h)))
        self.n_capsules = 0

    def _check(self, rc):
        if rc != PRTC_OK:
            raise PrtcError(rc, (self.lib.prtc_last_error() or b"").decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.prtc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- registration ------------------------------------------------------------------------------------
    def add_model_collider(self, xyz, mesh_first, mesh_count, transform=None):
        """PhysicsSystem::addModelCollider.  transform = (pos3, quat wxyz, scale3) like the .prt scene files."""
        xyz = np.ascontiguousarray(xyz, np.float32).reshape(-1, 3)
        mf = np.ascontiguousarray(mesh_first, np.uint32)
        mc = np.ascontiguousarray(mesh_count, np.uint32)
        t = np.zeros(1, TRANSFORM_DTYPE)
        if transform is None:
            t["rotation"][0, 3] = 1.0
            t["scale"] = 1.0
        else:
            tr = np.asarray(transform, np.float32)
            t["position"][0] = tr[0:3]
            t["rotation"][0] = (tr[4], tr[5], tr[6], tr[3])      # wxyz -> xyzw storage
            t["scale"][0] = tr[7:10]
        mid, first = ctypes.c_uint32(), ctypes.c_uint32()
        self._check(self.lib.prtc_add_model(self.h, _vp(xyz), ctypes.c_uint32(xyz.shape[0]), _vp(mf), _vp(mc),
                                            ctypes.c_uint32(len(mf)), _vp(t), ctypes.byref(mid), ctypes.byref(first)))
        return mid.value, first.value

    def add_capsule_collider(self, height=0.5, radius=0.5, offset=(0.0, 0.5, 0.0)):
        off = np.asarray(offset, np.float32)
        cid = ctypes.c_uint32()
        self._check(self.lib.prtc_add_capsule(self.h, ctypes.c_float(height), ctypes.c_float(radius), _vp(off), ctypes.byref(cid)))
        self.n_capsules += 1
        return cid.value

    def update_capsule_collider(self, cid, height, radius, offset):
        off = np.asarray(offset, np.float32)
        self._check(self.lib.prtc_update_capsule(self.h, ctypes.c_uint32(cid), ctypes.c_float(height), ctypes.c_float(radius), _vp(off)))

    def update_model_colliders(self, model_ids, transforms):
        ids = np.ascontiguousarray(model_ids, np.uint32)
        assert transforms.dtype == TRANSFORM_DTYPE
        self._check(self.lib.prtc_update_models(self.h, _vp(ids), _vp(transforms), ctypes.c_uint32(len(ids))))

    def remove_model_collider(self, model_id):
        self._check(self.lib.prtc_remove_model(self.h, ctypes.c_uint32(model_id)))

    def raycast(self, origins, directions, max_distances):
        """PhysicsSystem::raycast, batched: returns (hit_mask[n] bool, hit_xyz[n,3])"""
        o = np.ascontiguousarray(origins, np.float32).reshape(-1, 3)
        d = np.ascontiguousarray(directions, np.float32).reshape(-1, 3)
        m = np.ascontiguousarray(np.broadcast_to(np.asarray(max_distances, np.float32), (len(o),)), np.float32)
        hit = np.zeros((len(o), 3), np.float32)
        mask = np.zeros(len(o), np.uint8)
        self._check(self.lib.prtc_raycast(self.h, ctypes.c_uint32(len(o)), _vp(o), _vp(d), _vp(m), _vp(hit), _vp(mask)))
        return mask.astype(bool), hit

    def sweep_ellipsoids(self, centers, radii, velocities, iterations=3, very_close=0.005):
        """TIER B (parity unpinned): swept ellipsoids with collide-and-slide.  Returns a dict: pos[n,3], vel[n,3] (what is left
        of the motion), hit_tri[n,iterations] (-1 = none), hit_toi[n,iterations], hit_normal[n,iterations,3]."""
        c = np.ascontiguousarray(centers, np.float32).reshape(-1, 3)
        r = np.ascontiguousarray(np.broadcast_to(np.asarray(radii, np.float32), c.shape), np.float32)
        v = np.ascontiguousarray(velocities, np.float32).reshape(-1, 3)
        n = len(c)
        out = {"pos": np.zeros((n, 3), np.float32), "vel": np.zeros((n, 3), np.float32),
               "hit_tri": np.zeros((n, iterations), np.int32), "hit_toi": np.zeros((n, iterations), np.float32),
               "hit_normal": np.zeros((n, iterations, 3), np.float32)}
        self._check(self.lib.prtc_sweep_ellipsoids(self.h, ctypes.c_uint32(n), _vp(c), _vp(r), _vp(v), ctypes.c_uint32(iterations),
                                                   ctypes.c_float(very_close), _vp(out["pos"]), _vp(out["vel"]), _vp(out["hit_tri"]),
                                                   _vp(out["hit_toi"]), _vp(out["hit_normal"])))
        return out

    def commit(self):
        self._check(self.lib.prtc_commit(self.h))

    # -- per-frame -----------------------------------------------------------------------------------------
    def update_characters(self, dt, transforms, physics):
        """PhysicsSystem::updateCharacters on HOST arrays (TRANSFORM_DTYPE / PHYSICS_DTYPE), in place."""
        assert transforms.dtype == TRANSFORM_DTYPE and physics.dtype == PHYSICS_DTYPE and len(transforms) == len(physics)
        self._check(self.lib.prtc_step_characters(self.h, ctypes.c_float(dt), ctypes.c_uint32(len(transforms)),
                                                  _vp(transforms), _vp(physics)))

    def update_characters_device(self, dt, n, d_transforms_ptr, d_physics_ptr):
        self._check(self.lib.prtc_step_characters_device(self.h, ctypes.c_float(dt), ctypes.c_uint32(n),
                                                         ctypes.c_void_p(d_transforms_ptr), ctypes.c_void_p(d_physics_ptr)))

    def update_characters_traced(self, dt, transforms, physics, cap_queries=None, cap_ids=None, cap_hits=None):
        n = len(transforms)
        nsub = self.cfg.substeps
        cq = cap_queries or n * nsub
        ci = cap_ids or cq * 64
        ch = cap_hits or cq * 96
        arr = {
            "q_char": np.zeros(cq, np.uint32), "q_sub": np.zeros(cq, np.uint32), "q_aabb": np.zeros((cq, 6), np.float32),
            "q_mesh_off": np.zeros(cq + 1, np.uint32), "q_caps_off": np.zeros(cq + 1, np.uint32),
            "mesh_ids": np.zeros(ci, np.uint32), "caps_ids": np.zeros(ci, np.uint32), "q_tri_tests": np.zeros(cq, np.uint32),
            "h_query": np.zeros(ch, np.uint32), "h_poly": np.zeros(ch, np.int32), "h_impulse": np.zeros((ch, 3), np.float32),
            "h_normal": np.zeros((ch, 3), np.float32), "h_pos": np.zeros((ch, 3), np.float32),
        }
        ts = TraceStruct()
        ts.cap_queries, ts.cap_mesh_ids, ts.cap_caps_ids, ts.cap_hits = cq, ci, ci, ch
        for k, a in arr.items():
            setattr(ts, k, a.ctypes.data_as(dict(TraceStruct._fields_)[k]))
        self._check(self.lib.prtc_step_characters_traced(self.h, ctypes.c_float(dt), ctypes.c_uint32(n), _vp(transforms),
                                                         _vp(physics), ctypes.byref(ts)))
        nq, nm, nc, nh = ts.n_queries, ts.n_mesh_ids, ts.n_caps_ids, ts.n_hits
        out = {"q_char": arr["q_char"][:nq], "q_sub": arr["q_sub"][:nq], "q_aabb": arr["q_aabb"][:nq],
               "q_mesh_off": arr["q_mesh_off"][:nq + 1], "q_caps_off": arr["q_caps_off"][:nq + 1],
               "mesh_ids": arr["mesh_ids"][:nm], "caps_ids": arr["caps_ids"][:nc], "q_tri_tests": arr["q_tri_tests"][:nq],
               "h_query": arr["h_query"][:nh], "h_poly": arr["h_poly"][:nh], "h_impulse": arr["h_impulse"][:nh],
               "h_normal": arr["h_normal"][:nh], "h_pos": arr["h_pos"][:nh]}
        return out

    # -- resident state (prtc_resident_*) ------------------------------------------------------------------
    def resident_begin(self, transforms, physics):
        assert transforms.dtype == TRANSFORM_DTYPE and physics.dtype == PHYSICS_DTYPE and len(transforms) == len(physics)
        self._check(self.lib.prtc_resident_begin(self.h, ctypes.c_uint32(len(transforms)), _vp(transforms), _vp(physics)))
        self._resident_n = len(transforms)

    def resident_io(self):
        """(movement [n,3] float32 IN, result [n,4] float32 OUT): numpy views of the context's mapped pinned arrays"""
        mv, res = ctypes.POINTER(ctypes.c_float)(), ctypes.POINTER(ctypes.c_float)()
        self._check(self.lib.prtc_resident_io(self.h, ctypes.byref(mv), ctypes.byref(res)))
        n = self._resident_n
        if n == 0:
            return np.zeros((0, 3), np.float32), np.zeros((0, 4), np.float32)
        return (np.ctypeslib.as_array(mv, shape=(n, 3)), np.ctypeslib.as_array(res, shape=(n, 4)))

    def resident_step(self, dt):
        self._check(self.lib.prtc_resident_step(self.h, ctypes.c_float(dt)))

    def resident_edit(self, indices, transforms=None, physics=None):
        idx = np.ascontiguousarray(indices, dtype=np.uint32)
        self._check(self.lib.prtc_resident_edit(self.h, ctypes.c_uint32(len(idx)), _vp(idx), _vp(transforms), _vp(physics)))

    def resident_fetch(self):
        tf, ph = np.zeros(self._resident_n, TRANSFORM_DTYPE), np.zeros(self._resident_n, PHYSICS_DTYPE)
        self._check(self.lib.prtc_resident_fetch(self.h, _vp(tf), _vp(ph)))
        return tf, ph

    def resident_end(self):
        self._check(self.lib.prtc_resident_end(self.h))

    def dyn_configure(self, first_global, n_global):
        self._check(self.lib.prtc_dyn_configure(self.h, ctypes.c_uint32(first_global), ctypes.c_uint32(n_global)))

    def dyn_prepare(self, dt, n, d_transforms_ptr, d_physics_ptr):
        self._check(self.lib.prtc_dyn_prepare(self.h, ctypes.c_float(dt), ctypes.c_uint32(n), ctypes.c_void_p(d_transforms_ptr),
                                              ctypes.c_void_p(d_physics_ptr)))

    def dyn_buffer(self):
        """(device pointer, bytes per record, first global index, global count) of the record array to all-gather"""
        ptr, rb, first, ng = ctypes.c_void_p(), ctypes.c_size_t(), ctypes.c_uint32(), ctypes.c_uint32()
        self._check(self.lib.prtc_dyn_buffer(self.h, ctypes.byref(ptr), ctypes.byref(rb), ctypes.byref(first), ctypes.byref(ng)))
        return ptr.value, rb.value, first.value, ng.value

    def dyn_ipc_export(self, rank, world):
        """128 bytes of CUDA IPC handles (record array + arrival table) for the other ranks"""
        buf = ctypes.create_string_buffer(128)
        self._check(self.lib.prtc_dyn_ipc_export(self.h, ctypes.c_uint32(rank), ctypes.c_uint32(world), buf))
        return buf.raw

    def dyn_ipc_import(self, handles):
        """handles: list of the ranks' 128-byte blobs in rank order (own slot is ignored)"""
        if handles is None:                                        # give the exchange up again
            self._check(self.lib.prtc_dyn_ipc_import(self.h, None))
            return
        blob = b"".join(handles)
        self._check(self.lib.prtc_dyn_ipc_import(self.h, ctypes.c_char_p(blob)))

    def set_option(self, option, value):
        self._check(self.lib.prtc_set_option(self.h, ctypes.c_int(option), ctypes.c_int(value)))

    def set_stream(self, cuda_stream_handle):
        self._check(self.lib.prtc_set_stream(self.h, ctypes.c_void_p(cuda_stream_handle)))

    def synchronize(self):
        self._check(self.lib.prtc_synchronize(self.h))

    # -- introspection -----------------------------------------------------------------------------------
    def counters(self, reset=False):
        c = Counters()
        self._check(self.lib.prtc_get_counters(self.h, ctypes.byref(c), ctypes.c_int(1 if reset else 0)))
        return {k: int(getattr(c, k)) for k, _ in Counters._fields_}

    def debug_timeline(self):
        """(claims, 3) uint64 {start ns, end ns, SM} of the last pooled launch (PRTC_TIMELINE=1), rows of unused claims are zero"""
        nw = ctypes.c_size_t(0)
        self._check(self.lib.prtc_debug_timeline(self.h, None, 0, ctypes.byref(nw)))
        out = np.zeros(nw.value, np.uint64)
        if nw.value:
            self._check(self.lib.prtc_debug_timeline(self.h, out.ctypes.data_as(ctypes.c_void_p), nw.value, ctypes.byref(nw)))
        return out[: (len(out) // 3) * 3].reshape(-1, 3)

    def commit_stats(self):
        a, b = ctypes.c_uint32(), ctypes.c_uint32()
        self._check(self.lib.prtc_get_commit_stats(self.h, ctypes.byref(a), ctypes.byref(b)))
        return {"incremental": a.value, "full": b.value}

    def tree_info(self):
        a, b, c = ctypes.c_uint32(), ctypes.c_uint32(), ctypes.c_uint32()
        self._check(self.lib.prtc_get_tree_info(self.h, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c)))
        return {"nodes": a.value, "leaves": b.value, "triangles": c.value}

    def leaf_order(self):
        n = self.tree_info()["leaves"]
        out = np.zeros(max(n, 1), np.uint32)
        self._check(self.lib.prtc_get_leaf_order(self.h, _vp(out), ctypes.c_uint32(len(out))))
        return out[:n]

    def world_vertices(self):
        n = ctypes.c_uint32()
        self._check(self.lib.prtc_get_world_vertices(self.h, None, ctypes.c_uint32(0), ctypes.byref(n)))
        out = np.zeros((max(n.value, 1), 3), np.float32)
        self._check(self.lib.prtc_get_world_vertices(self.h, _vp(out), ctypes.c_uint32(len(out)), ctypes.byref(n)))
        return out[:n.value]

    def selftest_math(self, n_threads, seed, mode):
        bad = ctypes.c_uint64()
        self._check(self.lib.prtc_selftest_math(self.h, ctypes.c_uint32(n_threads), ctypes.c_uint64(seed), ctypes.c_int(mode), ctypes.byref(bad)))
        return bad.value

    def query_box(self, aabb6, capacity=4096):
        box = np.ascontiguousarray(aabb6, np.float32)
        out = np.zeros(capacity, np.uint32)
        n = ctypes.c_uint32()
        self._check(self.lib.prtc_query_box(self.h, _vp(box), _vp(out), ctypes.c_uint32(capacity), ctypes.byref(n)))
        return out[:n.value]


def host_build_tree(aabbs6, margin=0.05):
    """prtc_host_build_tree: reference-algorithm tree over boxes, flattened.  No device needed."""
    boxes = np.ascontiguousarray(aabbs6, np.float32).reshape(-1, 6)
    n = len(boxes)
    order = np.zeros(max(n, 1), np.uint32)
    nodes = np.zeros((max(2 * n, 1), 8), np.float32)
    nn = ctypes.c_uint32()
    rc = lib().prtc_host_build_tree(_vp(boxes), ctypes.c_uint32(n), ctypes.c_float(margin), _vp(order), _vp(nodes),
                                    ctypes.c_uint32(len(nodes)), ctypes.byref(nn))
    if rc != PRTC_OK:
        raise PrtcError(rc, (lib().prtc_last_error() or b"").decode())
    return order[:n], nodes[:nn.value]


def host_top_table(nodes8, cap=1024):
    """prtc_host_top_table: the top-of-tree table the device stages in shared memory.  No device needed."""
    nodes = np.ascontiguousarray(nodes8, np.float32).reshape(-1, 8)
    top = np.zeros((max(cap, 1), 8), np.float32)
    nt = ctypes.c_uint32()
    rc = lib().prtc_host_top_table(_vp(nodes), ctypes.c_uint32(len(nodes)), ctypes.c_uint32(cap), _vp(top), ctypes.c_uint32(len(top)),
                                   ctypes.byref(nt))
    if rc != PRTC_OK:
        raise PrtcError(rc, (lib().prtc_last_error() or b"").decode())
    return top[:nt.value]


def host_balanced_tree(aabbs6):
    """prtc_host_balanced_tree: the tree PRTC_ORDER_BY_ID builds on the device, by the same index arithmetic.  No device needed."""
    boxes = np.ascontiguousarray(aabbs6, np.float32).reshape(-1, 6)
    n = len(boxes)
    nodes = np.zeros((max(2 * n - 1, 1), 8), np.float32)
    nn = ctypes.c_uint32()
    rc = lib().prtc_host_balanced_tree(_vp(boxes), ctypes.c_uint32(n), _vp(nodes), ctypes.c_uint32(len(nodes)), ctypes.byref(nn))
    if rc != PRTC_OK:
        raise PrtcError(rc, (lib().prtc_last_error() or b"").decode())
    return nodes[:nn.value]


class MultiContext:
    """prtc_multi: several devices of one node behind one handle (one process).  Same method names as PhysicsContext."""

    def __init__(self, devices, abs_mode=ABS_INT, order_mode=ORDER_CANONICAL, dyn_mode=DYN_OFF, substeps=4, **kw):
        self.lib = lib()
        cfg = Config()
        self.lib.prtc_default_config(ctypes.byref(cfg))
        cfg.abs_mode, cfg.order_mode, cfg.dyn_mode, cfg.substeps = abs_mode, order_mode, dyn_mode, substeps
        for k, v in kw.items():
            setattr(cfg, k, v)
        self.cfg = cfg
        self.devices = list(devices)
        dev = (ctypes.c_int32 * len(self.devices))(*self.devices)
        self.h = ctypes.c_void_p()
        self._check(self.lib.prtc_multi_create(ctypes.byref(cfg), dev, ctypes.c_uint32(len(self.devices)), ctypes.byref(self.h)))
        self._resident_n = 0

    def _check(self, rc):
        if rc != PRTC_OK:
            raise PrtcError(rc, self.lib.prtc_last_error().decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.prtc_multi_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def add_model_collider(self, xyz, mesh_first, mesh_count, transform=None):
        xyz = np.ascontiguousarray(xyz, dtype=np.float32).reshape(-1, 3)
        mf = np.ascontiguousarray(mesh_first, dtype=np.uint32)
        mc = np.ascontiguousarray(mesh_count, dtype=np.uint32)
        t = make_transforms(np.zeros((1, 3), np.float32))
        if transform is not None:
            tr = np.asarray(transform, dtype=np.float32)
            t["position"][0] = tr[0:3]
            t["rotation"][0] = (tr[4], tr[5], tr[6], tr[3])
            t["scale"][0] = tr[7:10]
        mid, first = ctypes.c_uint32(), ctypes.c_uint32()
        self._check(self.lib.prtc_multi_add_model(self.h, _vp(xyz), ctypes.c_uint32(xyz.shape[0]), _vp(mf), _vp(mc), ctypes.c_uint32(len(mf)),
                                                  _vp(t), ctypes.byref(mid), ctypes.byref(first)))
        return mid.value, first.value

    def add_capsule_collider(self, height=0.5, radius=0.5, offset=(0.0, 0.5, 0.0)):
        off = (ctypes.c_float * 3)(*offset)
        cid = ctypes.c_uint32()
        self._check(self.lib.prtc_multi_add_capsule(self.h, ctypes.c_float(height), ctypes.c_float(radius), off, ctypes.byref(cid)))
        return cid.value

    def commit(self):
        self._check(self.lib.prtc_multi_commit(self.h))

    def set_option(self, option, value):
        self._check(self.lib.prtc_multi_set_option(self.h, ctypes.c_int(option), ctypes.c_int(value)))

    def update_characters(self, dt, transforms, physics):
        assert transforms.dtype == TRANSFORM_DTYPE and physics.dtype == PHYSICS_DTYPE and len(transforms) == len(physics)
        self._check(self.lib.prtc_multi_step_characters(self.h, ctypes.c_float(dt), ctypes.c_uint32(len(transforms)), _vp(transforms), _vp(physics)))

    def resident_begin(self, transforms, physics):
        self._check(self.lib.prtc_multi_resident_begin(self.h, ctypes.c_uint32(len(transforms)), _vp(transforms), _vp(physics)))
        self._resident_n = len(transforms)

    def resident_io(self):
        mv, res = ctypes.POINTER(ctypes.c_float)(), ctypes.POINTER(ctypes.c_float)()
        self._check(self.lib.prtc_multi_resident_io(self.h, ctypes.byref(mv), ctypes.byref(res)))
        n = self._resident_n
        return (np.ctypeslib.as_array(mv, shape=(n, 3)), np.ctypeslib.as_array(res, shape=(n, 4)))

    def resident_step(self, dt):
        self._check(self.lib.prtc_multi_resident_step(self.h, ctypes.c_float(dt)))

    def resident_fetch(self):
        tf, ph = np.zeros(self._resident_n, TRANSFORM_DTYPE), np.zeros(self._resident_n, PHYSICS_DTYPE)
        self._check(self.lib.prtc_multi_resident_fetch(self.h, _vp(tf), _vp(ph)))
        return tf, ph

    def counters(self, reset=False):
        c = Counters()
        self._check(self.lib.prtc_multi_get_counters(self.h, ctypes.byref(c), ctypes.c_int(1 if reset else 0)))
        return {k: int(getattr(c, k)) for k, _ in Counters._fields_}
