"""Shared helpers of the parity tests: scene dictionaries, world builders, exact comparisons."""
import numpy as np

from oraclelib import L0World, L1World, Trace
from prototype2_b200 import scenes


def bits(a):
    """float32 -> uint32 view with every NaN mapped to one pattern (CPU and GPU default NaNs differ)."""
    a = np.ascontiguousarray(a)
    if a.dtype != np.float32:
        return a
    u = a.view(np.uint32).copy()
    u[np.isnan(a)] = 0x7FC00000
    return u


def same(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return a.shape == b.shape and np.array_equal(bits(a), bits(b))


def make_scene(n_terrain, bx, bz, n_caps, seed=12345, separated=False, rot=None, model_transform=None,
               capsule=(0.5, 0.5, (0.0, 0.5, 0.0)), yaw=False):
    xyz, mf, mc = scenes.terrain(n_terrain, bx, bz)
    if separated:
        pos, vel = scenes.capsules_grid(n_caps, n_terrain, seed=seed)
    else:
        pos, vel = scenes.capsules(n_caps, n_terrain, seed=seed)
    rs = np.random.RandomState(seed + 1)
    if rot is None:
        rot = np.zeros((n_caps, 4), np.float32)
        rot[:, 0] = 1.0                                   # w,x,y,z identity
        if yaw:
            ang = rs.uniform(-np.pi, np.pi, n_caps)
            rot[:, 0] = np.cos(ang / 2)
            rot[:, 2] = np.sin(ang / 2)
    return {"xyz": xyz, "mesh_first": mf, "mesh_count": mc, "model_transform": model_transform, "capsule": capsule,
            "pos": pos.astype(np.float32), "vel": vel.astype(np.float32), "rot": np.asarray(rot, np.float32),
            "move": np.zeros((n_caps, 3), np.float32), "n": n_caps, "terrain_n": n_terrain}


def random_rotations(n, seed=7, tilt=0.35):
    """unit-ish quaternions (w,x,y,z) within `tilt` rad of upright plus arbitrary yaw, deliberately NOT normalised
    exactly (the reference normalises them itself)."""
    rs = np.random.RandomState(seed)
    yaw = rs.uniform(-np.pi, np.pi, n)
    ax = rs.normal(size=(n, 3))
    ax /= np.linalg.norm(ax, axis=1, keepdims=True)
    ang = rs.uniform(0, tilt, n)
    qt = np.concatenate([np.cos(ang / 2)[:, None], ax * np.sin(ang / 2)[:, None]], axis=1)
    qy = np.stack([np.cos(yaw / 2), np.zeros(n), np.sin(yaw / 2), np.zeros(n)], axis=1)

    def mul(a, b):
        w1, x1, y1, z1 = a.T
        w2, x2, y2, z2 = b.T
        return np.stack([w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2, w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2,
                         w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2, w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2], axis=1)
    q = mul(qy, qt) * rs.uniform(0.9, 1.1, (n, 1))
    return q.astype(np.float32)


def build_oracle(scene, kind="l1", **kw):
    w = L0World(kw.pop("variant", "plain")) if kind == "l0" else L1World(**kw)
    w.add_model(scene["xyz"], scene["mesh_first"], scene["mesh_count"], scene["model_transform"])
    h, r, off = scene["capsule"]
    for _ in range(scene["n"]):
        w.add_capsule(h, r, off)
    w.set_state(pos=scene["pos"], rot=scene["rot"], vel=scene["vel"], move=scene["move"])
    return w


def build_gpu(scene, **kw):
    from prototype2_b200 import capi
    ctx = capi.PhysicsContext(**kw)
    ctx.add_model_collider(scene["xyz"], scene["mesh_first"], scene["mesh_count"], scene["model_transform"])
    h, r, off = scene["capsule"]
    for _ in range(scene["n"]):
        ctx.add_capsule_collider(h, r, off)
    rot = scene["rot"]
    tf = capi.make_transforms(scene["pos"], rot_xyzw=rot[:, [1, 2, 3, 0]])
    ph = capi.make_physics(scene["vel"], movement=scene["move"])
    return ctx, tf, ph


def gpu_state(tf, ph):
    return {"pos": tf["position"].copy(), "vel": ph["velocity"].copy(), "gnormal": ph["ground_normal"].copy(),
            "grounded": ph["grounded"].astype(np.uint8)}


def state_diff(a, b):
    return [k for k in ("pos", "vel", "gnormal", "grounded") if not same(a[k], b[k])]


TRACE_KEYS = ("q_char", "q_sub", "q_aabb", "q_mesh_off", "mesh_ids", "q_tri_tests", "h_query", "h_poly", "h_impulse",
              "h_normal", "h_pos")


def trace_diff(oracle_trace, other, keys=TRACE_KEYS):
    bad = []
    for k in keys:
        x = getattr(oracle_trace, k) if isinstance(oracle_trace, Trace) else oracle_trace[k]
        y = getattr(other, k) if isinstance(other, Trace) else other[k]
        if not same(x, y):
            bad.append(k)
    return bad


def _tget(t, k):
    return getattr(t, k) if isinstance(t, Trace) else t[k]


def compare_with_literal(t_lit, t_can, s_lit, s_can, n_chars):
    """P1/P3 comparison of a CANONICAL-order run (static-only tree) against the LITERAL reference order (unified
    tree whose capsule-leaf churn permutes the DFS order of static leaves, SURVEY.md finding 3 / H4).

    A character is 'in sync' until one of its queries sees a different candidate SEQUENCE.  While in sync every
    query box, candidate list, hit index, push vector and post-push position must be bit-identical, and a
    character that stays in sync for the whole frame must end bit-identical.  At the first out-of-sync query the
    candidate SET must still be equal (the set is tree independent).  Returns statistics."""
    per_char_lit = {}
    for q in range(len(_tget(t_lit, "q_char"))):
        per_char_lit.setdefault(int(_tget(t_lit, "q_char")[q]), []).append(q)
    per_char_can = {}
    for q in range(len(_tget(t_can, "q_char"))):
        per_char_can.setdefault(int(_tget(t_can, "q_char")[q]), []).append(q)

    def hits_of(t):
        out = {}
        hq = _tget(t, "h_query")
        for h in range(len(hq)):
            out.setdefault(int(hq[h]), []).append(h)
        return out
    hl, hc = hits_of(t_lit), hits_of(t_can)
    in_sync = np.ones(n_chars, bool)
    stats = {"queries_compared": 0, "queries_seq_equal": 0, "set_checked": 0}
    for ch in range(n_chars):
        ql, qc = per_char_lit.get(ch, []), per_char_can.get(ch, [])
        for a, b in zip(ql, qc):
            assert same(_tget(t_lit, "q_aabb")[a], _tget(t_can, "q_aabb")[b]), "query box differs while in sync (char %d)" % ch
            ma = _tget(t_lit, "mesh_ids")[_tget(t_lit, "q_mesh_off")[a]:_tget(t_lit, "q_mesh_off")[a + 1]]
            mb = _tget(t_can, "mesh_ids")[_tget(t_can, "q_mesh_off")[b]:_tget(t_can, "q_mesh_off")[b + 1]]
            stats["queries_compared"] += 1
            assert np.array_equal(np.sort(ma), np.sort(mb)), "candidate SET differs (char %d)" % ch
            stats["set_checked"] += 1
            if not np.array_equal(ma, mb):
                in_sync[ch] = False
                break
            stats["queries_seq_equal"] += 1
            ha, hb = hl.get(a, []), hc.get(b, [])
            # capsule-capsule contacts (poly -1) only exist in the literal run; scenes here keep capsules apart
            assert len(ha) == len(hb), "hit count differs while in sync (char %d)" % ch
            for x, y in zip(ha, hb):
                assert _tget(t_lit, "h_poly")[x] == _tget(t_can, "h_poly")[y]
                for k in ("h_impulse", "h_normal", "h_pos"):
                    assert same(_tget(t_lit, k)[x], _tget(t_can, k)[y]), "%s differs while in sync (char %d)" % (k, ch)
        else:
            assert len(ql) == len(qc), "substep count differs while in sync (char %d)" % ch
    for k in ("pos", "vel", "gnormal", "grounded"):
        assert same(s_lit[k][in_sync], s_can[k][in_sync]), "final %s differs for an in-sync character" % k
    out = ~in_sync
    stats["chars_in_sync"] = int(in_sync.sum())
    stats["chars_total"] = n_chars
    if out.any():
        d = np.abs(s_lit["pos"][out] - s_can["pos"][out]).max(axis=1)
        stats["out_of_sync_max_pos_dev"] = float(d.max())
        stats["out_of_sync_median_pos_dev"] = float(np.median(d))
    return stats


def c1_scene(g):
    """scene dict of the shipped-level fixture tests/golden/c1_level.npz (see tests/golden/make_c1.py)"""
    cap = g["capsule"]
    return {"xyz": g["xyz"], "mesh_first": g["mesh_first"], "mesh_count": g["mesh_count"],
            "model_transform": list(g["model_transform"]), "capsule": (float(cap[0]), float(cap[1]), tuple(float(x) for x in cap[2:5])),
            "pos": g["pos"], "vel": np.zeros((1, 3), np.float32), "rot": np.array([[1, 0, 0, 0]], np.float32),
            "move": np.zeros((1, 3), np.float32), "n": 1}
