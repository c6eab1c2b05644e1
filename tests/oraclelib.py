"""ctypes bindings to the CPU oracles (TEST INFRASTRUCTURE ONLY).

L0 = the reference's own hot-path sources compiled verbatim (oracle/_ref/libprt_l0*.so),
L1 = our restatement (oracle/_build/libprt_l1.so).  Both expose the same tiny C-ABI
(oracle/l0_driver.cpp, oracle/l1_oracle.cpp); this module gives them one Python face so the
parity tests can drive either.  Nothing under prototype2_b200/ may import this module.
"""
import ctypes
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")

_vp = ctypes.c_void_p
_f = ctypes.c_float
_u32 = ctypes.c_uint32


def _ptr(a):
    return None if a is None else a.ctypes.data_as(_vp)


def _f32(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float32)
    if shape is not None:
        a = a.reshape(shape)
    return a


class Trace:
    """Per-frame trace: one record per broadphase query and one per contact."""

    def __init__(self, d):
        self.__dict__.update(d)

    def mesh_candidates(self, q):
        return self.mesh_ids[self.q_mesh_off[q]:self.q_mesh_off[q + 1]]

    def capsule_candidates(self, q):
        return self.caps_ids[self.q_caps_off[q]:self.q_caps_off[q + 1]]

    FIELDS = ("q_char", "q_sub", "q_aabb", "q_mesh_off", "q_caps_off", "mesh_ids", "caps_ids",
              "q_tri_tests", "q_node_tests", "h_query", "h_poly", "h_impulse", "h_normal", "h_pos")


class _World:
    prefix = None

    def __init__(self, lib, handle):
        self.lib = lib
        self.h = _vp(handle)
        self.n = 0

    def _fn(self, name, restype=None):
        f = getattr(self.lib, self.prefix + name)
        f.restype = restype
        return f

    def close(self):
        if self.h:
            self._fn("destroy")(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def add_model(self, xyz, mesh_first, mesh_count, transform=None):
        xyz = _f32(xyz, (-1, 3))
        mesh_first = np.ascontiguousarray(mesh_first, dtype=np.uint32)
        mesh_count = np.ascontiguousarray(mesh_count, dtype=np.uint32)
        tf = _f32([0, 0, 0, 1, 0, 0, 0, 1, 1, 1] if transform is None else transform)
        return self._fn("add_model", ctypes.c_int)(self.h, _ptr(xyz), _u32(xyz.shape[0]), _ptr(mesh_first),
                                                   _ptr(mesh_count), _u32(len(mesh_first)), _ptr(tf))

    def add_capsule(self, height=0.5, radius=0.5, offset=(0.0, 0.5, 0.0)):
        off = _f32(offset)
        self.n += 1
        return self._fn("add_capsule", ctypes.c_int)(self.h, _f(height), _f(radius), _ptr(off))

    def set_state(self, pos=None, rot=None, vel=None, move=None, gnormal=None, grounded=None):
        args = [None if a is None else _f32(a, s) for a, s in
                ((pos, (self.n, 3)), (rot, (self.n, 4)), (vel, (self.n, 3)), (move, (self.n, 3)), (gnormal, (self.n, 3)))]
        g = None if grounded is None else np.ascontiguousarray(grounded, dtype=np.uint8)
        self._fn("set_state")(self.h, *[_ptr(a) for a in args], _ptr(g))

    def get_state(self):
        pos = np.zeros((self.n, 3), np.float32)
        vel = np.zeros((self.n, 3), np.float32)
        gn = np.zeros((self.n, 3), np.float32)
        gr = np.zeros(self.n, np.uint8)
        self._fn("get_state")(self.h, _ptr(pos), _ptr(vel), _ptr(gn), _ptr(gr))
        return {"pos": pos, "vel": vel, "gnormal": gn, "grounded": gr}

    def step(self, dt=1.0 / 30.0, trace=False):
        self._fn("step")(self.h, _f(dt), ctypes.c_int(1 if trace else 0))
        return self.get_trace() if trace else None

    def get_trace(self):
        nq = self._fn("trace_num_queries", _u32)(self.h)
        nh = self._fn("trace_num_hits", _u32)(self.h)
        nm = self._fn("trace_num_mesh_ids", _u32)(self.h)
        nc = self._fn("trace_num_caps_ids", _u32)(self.h)
        d = {
            "q_char": np.zeros(nq, np.uint32), "q_sub": np.zeros(nq, np.uint32), "q_aabb": np.zeros((nq, 6), np.float32),
            "q_mesh_off": np.zeros(nq + 1, np.uint32), "q_caps_off": np.zeros(nq + 1, np.uint32),
            "mesh_ids": np.zeros(nm, np.uint32), "caps_ids": np.zeros(nc, np.uint32),
            "q_tri_tests": np.zeros(nq, np.uint32), "q_node_tests": np.zeros(nq, np.uint32),
            "h_query": np.zeros(nh, np.uint32), "h_poly": np.zeros(nh, np.int32),
            "h_impulse": np.zeros((nh, 3), np.float32), "h_normal": np.zeros((nh, 3), np.float32),
            "h_pos": np.zeros((nh, 3), np.float32),
        }
        self._fn("trace_get")(self.h, *[_ptr(d[k]) for k in Trace.FIELDS])
        return Trace(d)

    def raycast(self, origin, direction, max_distance):
        hit = np.zeros(3, np.float32)
        ok = self._fn("raycast", ctypes.c_bool)(self.h, _ptr(_f32(origin)), _ptr(_f32(direction)), _f(max_distance), _ptr(hit))
        return bool(ok), hit


class L0World(_World):
    """The reference's own code.  variant: 'plain' (as this toolchain compiles it: `abs(dist)` truncates to int),
    'fabs' (float abs, libc++ semantics) or 'traced' (plain + candidate/hit traces)."""
    prefix = "l0_"
    LIBS = {"plain": "libprt_l0.so", "fabs": "libprt_l0_fabs.so", "traced": "libprt_l0_traced.so"}

    def __init__(self, variant="plain"):
        path = os.path.join(ORACLE_DIR, "_ref", self.LIBS[variant])
        lib = ctypes.CDLL(path)
        lib.l0_create.restype = _vp
        super().__init__(lib, lib.l0_create())
        self.variant = variant

    @staticmethod
    def available(variant="plain"):
        return os.path.exists(os.path.join(ORACLE_DIR, "_ref", L0World.LIBS[variant]))

    def update_model(self, model_index, transform):
        self._fn("update_model")(self.h, _u32(model_index), _ptr(_f32(transform)))

    def update_capsule(self, capsule_index, height, radius, offset):
        self._fn("update_capsule")(self.h, _u32(capsule_index), _f(height), _f(radius), _ptr(_f32(offset)))

    def remove_model(self, model_index):
        """PhysicsSystem::removeCollider(MODEL tag); see the envelope note at l0_remove_model"""
        self._fn("remove_model")(self.h, _u32(model_index))

    def dump_order(self):
        m = np.zeros(70000, np.uint16)
        c = np.zeros(70000, np.uint16)
        nm, nc = _u32(0), _u32(0)
        self.lib.l0_dump_order(self.h, _ptr(m), ctypes.byref(nm), _ptr(c), ctypes.byref(nc))
        return m[:nm.value].astype(np.uint32), c[:nc.value].astype(np.uint32)

    def query(self, caller, aabb6, cap=70000):
        m = np.zeros(cap, np.uint16)
        c = np.zeros(cap, np.uint16)
        nm, nc = _u32(0), _u32(0)
        self.lib.l0_query(self.h, ctypes.c_uint16(caller), _ptr(_f32(aabb6)), _ptr(m), ctypes.byref(nm), _ptr(c),
                          ctypes.byref(nc), _u32(cap))
        return m[:nm.value].astype(np.uint32), c[:nc.value].astype(np.uint32)

    def contacts(self):
        self.lib.l0_contacts.restype = ctypes.c_uint64
        return self.lib.l0_contacts(self.h)


class L1World(_World):
    """Our CPU restatement.  order: 'literal' | 'canonical' | 'by_id' (canonical with the mesh candidates of a query in ascending collider id); dyn: 'off' | 'jacobi' (canonical only)."""
    prefix = "l1_"

    def __init__(self, order="literal", dyn="off", abs_mode="int", substeps=4, threads=1, gravity=1.0,
                 leaf_margin=0.05, ground_dot=0.3, friction=10.0, stick=0.05):
        path = os.path.join(ORACLE_DIR, "_build", "libprt_l1.so")
        lib = ctypes.CDLL(path)
        lib.l1_create.restype = _vp
        cf = np.array([gravity, leaf_margin, ground_dot, friction, stick], np.float32)
        cu = np.array([substeps, {"int": 0, "float": 1}[abs_mode], {"literal": 0, "canonical": 1, "by_id": 2}[order],
                       {"off": 0, "jacobi": 2}[dyn], threads], np.uint32)
        super().__init__(lib, lib.l1_create(_ptr(cf), _ptr(cu)))

    def update_model(self, model_index, transform):
        self._fn("update_model")(self.h, _u32(model_index), _ptr(_f32(transform)))

    def remove_model(self, model_index):
        self._fn("remove_model")(self.h, _u32(model_index))

    def counters(self, reset=False):
        out = np.zeros(5, np.uint64)
        self.lib.l1_counters(self.h, _ptr(out), ctypes.c_int(1 if reset else 0))
        return dict(zip(("substeps", "tri_tests", "node_tests", "contacts", "cc_tests"), (int(v) for v in out)))

    def dump_order(self, cap=1 << 22):
        nm, nc = _u32(0), _u32(0)
        self.lib.l1_dump_order(self.h, None, ctypes.byref(nm), None, ctypes.byref(nc))
        m = np.zeros(max(1, nm.value), np.uint32)
        c = np.zeros(max(1, nc.value), np.uint32)
        self.lib.l1_dump_order(self.h, _ptr(m), ctypes.byref(nm), _ptr(c), ctypes.byref(nc))
        return m[:nm.value], c[:nc.value]

    def query(self, caller, aabb6, cap=1 << 16):
        m = np.zeros(cap, np.uint32)
        c = np.zeros(cap, np.uint32)
        nm, nc = _u32(0), _u32(0)
        self.lib.l1_query(self.h, _u32(caller), _ptr(_f32(aabb6)), _ptr(m), ctypes.byref(nm), _ptr(c), ctypes.byref(nc), _u32(cap))
        return m[:nm.value], c[:nc.value]


def sweep_reference_available():
    return os.path.exists(os.path.join(ORACLE_DIR, "_ref", "libprt_l0_sweep.so"))


def sweep_reference(tri_xyz, centers, radii, velocities, iterations=3, very_close=0.005):
    """Checker of the Tier-B kernel (oracle/l0_sweep_driver.cpp): the reference's surviving Nettle helpers under our
    restatement of the loop, brute force over all triangles.  tri_xyz: world-space vertices, 3 per triangle, registration order."""
    lib = ctypes.CDLL(os.path.join(ORACLE_DIR, "_ref", "libprt_l0_sweep.so"))
    t = _f32(tri_xyz, (-1, 3))
    c = _f32(centers, (-1, 3))
    r = np.ascontiguousarray(np.broadcast_to(np.asarray(radii, np.float32), c.shape), np.float32)
    v = _f32(velocities, (-1, 3))
    n = len(c)
    out = {"pos": np.zeros((n, 3), np.float32), "vel": np.zeros((n, 3), np.float32),
           "hit_tri": np.zeros((n, iterations), np.int32), "hit_toi": np.zeros((n, iterations), np.float32),
           "hit_normal": np.zeros((n, iterations, 3), np.float32)}
    tests = ctypes.c_uint64()
    lib.l0_sweep.restype = None
    lib.l0_sweep(_ptr(t), _u32(len(t) // 3), _u32(n), _ptr(c), _ptr(r), _ptr(v), _u32(iterations), _f(very_close), _ptr(out["pos"]),
                 _ptr(out["vel"]), _ptr(out["hit_tri"]), _ptr(out["hit_toi"]), _ptr(out["hit_normal"]), ctypes.byref(tests))
    out["tri_tests"] = tests.value
    return out
