"""Synthetic scenes of the shapes BASELINE.json names: pure numpy, no device code.

Terrain: N x N unit cells on the xz-plane, heights y = 0.75*sin(0.07 x)*cos(0.05 z) at integer x,z; each cell
-> triangles (a,b,c),(a,c,d) with a=(x,.,z) b=(x,.,z+1) c=(x+1,.,z+1) d=(x+1,.,z) so that (b-a)x(c-a) faces +y
(back faces are culled, reference collision_system.cpp:71).  Mesh colliders ("leaves" of the reference's BVH) are
blocks of BX x BZ cells in block row-major order.  Capsules use the shipped collider h=0.5 r=0.5 offset (0,0.5,0)
(reference res/scenes/cavern.prt:9), placed with a Mersenne twister seeded 12345.
"""
import numpy as np


def height(x, z):
    x = np.asarray(x, dtype=np.float32)
    z = np.asarray(z, dtype=np.float32)
    return (np.float32(0.75) * np.sin(np.float32(0.07) * x) * np.cos(np.float32(0.05) * z)).astype(np.float32)


def terrain(n, bx=1, bz=1):
    """Returns (xyz [3*ntri,3] float32 index-expanded, mesh_first, mesh_count) with meshes = bx x bz cell blocks."""
    gx, gz = np.meshgrid(np.arange(n + 1, dtype=np.float32), np.arange(n + 1, dtype=np.float32), indexing="ij")
    gy = height(gx, gz)
    nbx = (n + bx - 1) // bx
    nbz = (n + bz - 1) // bz
    # cell order: block row-major (block x outer, block z inner), cells inside a block x outer, z inner
    cx, cz = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    bxi, bzi = cx // bx, cz // bz
    key = ((bxi * nbz + bzi).astype(np.int64) * (bx * bz) + (cx % bx) * bz + (cz % bz)).ravel()
    order = np.argsort(key, kind="stable")
    cx = cx.ravel()[order]
    cz = cz.ravel()[order]
    block = (bxi * nbz + bzi).ravel()[order]

    def vert(ix, iz):
        return np.stack([gx[ix, iz], gy[ix, iz], gz[ix, iz]], axis=-1)

    a, b, c, d = vert(cx, cz), vert(cx, cz + 1), vert(cx + 1, cz + 1), vert(cx + 1, cz)
    tris = np.stack([a, b, c, a, c, d], axis=1).reshape(-1, 3).astype(np.float32)   # 6 vertices per cell
    cells_per_block = np.bincount(block, minlength=nbx * nbz)
    cells_per_block = cells_per_block[cells_per_block > 0]
    mesh_count = (cells_per_block * 6).astype(np.uint32)
    mesh_first = np.concatenate([[0], np.cumsum(mesh_count)[:-1]]).astype(np.uint32)
    return tris, mesh_first, mesh_count


def _uniform01(rs, count):
    """std::uniform_real_distribution<float>(0,1) over std::mt19937 as libstdc++ computes it:
    float(raw32) / 2^32, clamped below 1."""
    raw = rs.randint(0, 2 ** 32, size=count, dtype=np.uint64).astype(np.uint32)
    u = raw.astype(np.float32) / np.float32(4294967296.0)
    return np.minimum(u, np.nextafter(np.float32(1.0), np.float32(0.0)))


def capsules(n_caps, n, seed=12345, min_separation=None):
    """Placement per SURVEY 8d: x,z in [2, N-2], y = H(x,z) - 0.1 + 0.4u, v = (-0.2+0.4u, -0.2u, -0.2+0.4u)."""
    rs = np.random.RandomState(seed)
    u = _uniform01(rs, 6 * n_caps).reshape(n_caps, 6)
    x = np.float32(2.0) + u[:, 0] * np.float32(n - 4)
    z = np.float32(2.0) + u[:, 1] * np.float32(n - 4)
    y = height(x, z) - np.float32(0.1) + np.float32(0.4) * u[:, 2]
    pos = np.stack([x, y, z], axis=1).astype(np.float32)
    vel = np.stack([np.float32(-0.2) + np.float32(0.4) * u[:, 3], np.float32(-0.2) * u[:, 4],
                    np.float32(-0.2) + np.float32(0.4) * u[:, 5]], axis=1).astype(np.float32)
    return pos, vel


def capsules_grid(n_caps, n, spacing=2.5, seed=12345):
    """Capsules on a jittered grid at least `spacing - 0.5` apart in x or z, so no two fat capsule AABBs can
    overlap within a frame: the reference's capsule-capsule pass never fires and the static pass is isolated."""
    rs = np.random.RandomState(seed)
    per_row = int((n - 4) // spacing)
    assert per_row * per_row >= n_caps, "terrain too small for that many separated capsules"
    idx = rs.permutation(per_row * per_row)[:n_caps]
    gx, gz = idx // per_row, idx % per_row
    u = _uniform01(rs, 6 * n_caps).reshape(n_caps, 6)
    x = np.float32(2.0) + gx.astype(np.float32) * np.float32(spacing) + np.float32(0.5) * u[:, 0]
    z = np.float32(2.0) + gz.astype(np.float32) * np.float32(spacing) + np.float32(0.5) * u[:, 1]
    y = height(x, z) - np.float32(0.1) + np.float32(0.4) * u[:, 2]
    pos = np.stack([x, y, z], axis=1).astype(np.float32)
    vel = np.stack([np.float32(-0.2) + np.float32(0.4) * u[:, 3], np.float32(-0.2) * u[:, 4],
                    np.float32(-0.2) + np.float32(0.4) * u[:, 5]], axis=1).astype(np.float32)
    return pos, vel


def morton_order(pos, n):
    """Permutation sorting capsules along a 2-D Morton curve in xz (warps then traverse the BVH coherently)."""
    def part1by1(v):
        v = v.astype(np.uint64) & np.uint64(0xFFFF)
        v = (v | (v << np.uint64(8))) & np.uint64(0x00FF00FF)
        v = (v | (v << np.uint64(4))) & np.uint64(0x0F0F0F0F)
        v = (v | (v << np.uint64(2))) & np.uint64(0x33333333)
        v = (v | (v << np.uint64(1))) & np.uint64(0x55555555)
        return v
    scale = 65535.0 / max(1.0, float(n))
    qx = np.clip(pos[:, 0] * scale, 0, 65535).astype(np.uint32)
    qz = np.clip(pos[:, 2] * scale, 0, 65535).astype(np.uint32)
    code = part1by1(qx) | (part1by1(qz) << np.uint64(1))
    return np.argsort(code, kind="stable")


CONFIGS = {
    # name: (terrain N, block x, block z, capsules)      
    "C2": (224, 1, 1, 1000),
    "C3": (708, 4, 2, 64000),
    "C4": (2237, 2, 2, 1000000),
    "C5": (708, 2, 2, 256000),
}
