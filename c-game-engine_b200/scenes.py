"""Synthetic workloads named in BASELINE.json / SURVEY §8d, built from the map fixtures under tests/golden/.

Host-side harness code (numpy only): deterministic inputs for the tests and bench.py.  Random numbers come from
a counter-based Philox4x32-10 so that any world / ray can be regenerated independently on any rank.
"""
from __future__ import annotations

import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

RAY_DTYPE = np.dtype([("origin", "<f4", 3), ("tmax", "<f4"), ("dir", "<f4", 3), ("mask", "<u4")])

KEY_RAYS = 0x5EED0003
KEY_LATTICE = 0x5EED0004
KEY_ENSEMBLE = 0x5EED0005


def _mulhilo(a: np.ndarray, b: int):
    p = a.astype(np.uint64) * np.uint64(b)
    return (p >> np.uint64(32)).astype(np.uint32), (p & np.uint64(0xFFFFFFFF)).astype(np.uint32)


def philox4x32(counter: np.ndarray, key: int, key_hi: int = 0) -> np.ndarray:
    """Philox4x32-10 (Salmon et al. 2011).  counter: (n, 4) uint32 -> (n, 4) uint32."""
    c = np.ascontiguousarray(counter, dtype=np.uint32).copy()
    k0, k1 = np.uint32(key & 0xFFFFFFFF), np.uint32(key_hi & 0xFFFFFFFF)
    with np.errstate(over="ignore"):
        for _ in range(10):
            hi0, lo0 = _mulhilo(c[:, 0], 0xD2511F53)
            hi1, lo1 = _mulhilo(c[:, 2], 0xCD9E8D57)
            c = np.stack([hi1 ^ c[:, 1] ^ k0, lo1, hi0 ^ c[:, 3] ^ k1, lo0], axis=1)
            k0 = np.uint32((int(k0) + 0x9E3779B9) & 0xFFFFFFFF)
            k1 = np.uint32((int(k1) + 0xBB67AE85) & 0xFFFFFFFF)
    return c


def uniform01(bits: np.ndarray) -> np.ndarray:
    """uint32 -> float32 in [0, 1) with 24 bits."""
    return ((bits >> np.uint32(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)).astype(np.float32)


def load_static(name: str):
    """Static collision meshes of a shipped map: list of (pos(3,), tris(T,3,3) relative to pos)."""
    z = np.load(os.path.join(GOLDEN, f"static_{name}.npz"))
    start = z["mesh_start"]
    return [(z["mesh_pos"][i], z["tris"][start[i]:start[i + 1]]) for i in range(len(z["mesh_pos"]))]


def box_map(x0=-512.0, x1=512.0, y0=-512.0, y1=512.0, z0=-512.0, z1=512.0):
    """12-triangle inward-facing box for mapSources/max_box.json, which ships no .gmap.
    Winding follows orb.gmap's sector: normals point into the room."""
    def quad(a, b, c, d):
        return [[a, b, c], [c, d, a]]
    p = lambda x, y, z: [x, y, z]
    tris = []
    tris += quad(p(x1, y0, z1), p(x1, y0, z0), p(x0, y0, z0), p(x0, y0, z1))  # floor, +y
    tris += quad(p(x0, y1, z0), p(x1, y1, z0), p(x1, y1, z1), p(x0, y1, z1))  # ceiling, -y
    tris += quad(p(x0, y0, z0), p(x0, y1, z0), p(x0, y1, z1), p(x0, y0, z1))  # x0 wall, +x
    tris += quad(p(x1, y0, z1), p(x1, y1, z1), p(x1, y1, z0), p(x1, y0, z0))  # x1 wall, -x
    tris += quad(p(x1, y0, z0), p(x1, y1, z0), p(x0, y1, z0), p(x0, y0, z0))  # z0 wall, +z
    tris += quad(p(x0, y0, z1), p(x0, y1, z1), p(x1, y1, z1), p(x1, y0, z1))  # z1 wall, -z
    t = np.array(tris, dtype=np.float32)
    n = np.cross(t[:, 1] - t[:, 0], t[:, 2] - t[:, 0])
    c = t.mean(axis=1)
    centre = np.array([(x0 + x1) / 2, (y0 + y1) / 2, (z0 + z1) / 2], np.float32)
    flip = np.einsum("ij,ij->i", n, centre - c) < 0
    t[flip] = t[flip][:, [0, 2, 1]]
    return [(np.zeros(3, np.float32), t)]


# ---- C2 / C5: column of physboxes over sector 0 of stacked.gmap
STACK_X, STACK_Z, STACK_FLOOR_Y = 0.0, -1.5, -1.5
BOX_HALF = 0.2
STACK_PITCH = 0.401


def stack_positions(n=8) -> np.ndarray:
    y0 = STACK_FLOOR_Y + BOX_HALF + 0.05
    return np.array([[STACK_X, y0 + STACK_PITCH * i, STACK_Z] for i in range(n)], dtype=np.float32)


def ensemble_velocities(worlds: int, boxes: int, first_world: int = 0, key: int = KEY_ENSEMBLE) -> np.ndarray:
    """Per-box initial linear velocity U(-0.5, 0.5)^3, counter = (world, box)."""
    w = np.repeat(np.arange(first_world, first_world + worlds, dtype=np.uint32), boxes)
    b = np.tile(np.arange(boxes, dtype=np.uint32), worlds)
    ctr = np.stack([w, b, np.zeros_like(w), np.zeros_like(w)], axis=1)
    r = philox4x32(ctr, key)
    v = uniform01(r[:, :3]) - np.float32(0.5)
    return v.reshape(worlds, boxes, 3).astype(np.float32)


def block_positions(nx=4, ny=4, nz=4, pitch=0.45) -> np.ndarray:
    """4x4x4 block variant of C5 (64 boxes) centred over the same floor spot."""
    out = []
    for j in range(ny):
        for i in range(nx):
            for k in range(nz):
                out.append([STACK_X + (i - (nx - 1) / 2) * pitch, STACK_FLOOR_Y + BOX_HALF + 0.05 + j * pitch,
                            STACK_Z + (k - (nz - 1) / 2) * pitch])
    return np.array(out, dtype=np.float32)


# ---- C3: batched hitscan rays against shapes.gmap
def shapes_rays(n: int, mesh_pos: np.ndarray, first: int = 0, tmax: float = 50.0, mask: int = 1,
                key: int = KEY_RAYS) -> np.ndarray:
    """origin = pos of mesh (i mod M) + U(-0.25, 0.25)^3; direction uniform on the sphere (Marsaglia 1972)."""
    idx = np.arange(first, first + n, dtype=np.uint32)
    zeros = np.zeros_like(idx)
    r0 = philox4x32(np.stack([idx, zeros, zeros, zeros], axis=1), key)
    jitter = (uniform01(r0[:, :3]) - np.float32(0.5)) * np.float32(0.5)
    origin = mesh_pos[idx % len(mesh_pos)].astype(np.float32) + jitter
    d = np.zeros((n, 3), np.float32)
    todo = np.arange(n)
    attempt = 1
    while len(todo):
        r = philox4x32(np.stack([idx[todo], np.full(len(todo), attempt, np.uint32), zeros[todo], zeros[todo]], axis=1), key)
        x1 = uniform01(r[:, 0]) * np.float32(2) - np.float32(1)
        x2 = uniform01(r[:, 1]) * np.float32(2) - np.float32(1)
        s = x1 * x1 + x2 * x2
        ok = (s < 1.0) & (s > 0.0)
        sq = np.sqrt(np.float32(1) - s[ok]).astype(np.float32)
        d[todo[ok], 0] = 2 * x1[ok] * sq
        d[todo[ok], 1] = 2 * x2[ok] * sq
        d[todo[ok], 2] = 1 - 2 * s[ok]
        todo = todo[~ok]
        attempt += 1
    d /= np.linalg.norm(d, axis=1, keepdims=True).astype(np.float32)
    rays = np.zeros(n, RAY_DTYPE)
    rays["origin"] = origin
    rays["tmax"] = tmax
    rays["dir"] = d.astype(np.float32)
    rays["mask"] = mask
    return rays


# ---- C4: lattice of boxes in the big room
def lattice_positions(nx=100, ny=10, nz=100, pitch=0.5, floor_y=-512.0, key: int = KEY_LATTICE) -> np.ndarray:
    i, j, k = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    idx = np.arange(nx * ny * nz, dtype=np.uint32)
    zeros = np.zeros_like(idx)
    r = philox4x32(np.stack([idx, zeros, zeros, zeros], axis=1), key)
    jit = (uniform01(r[:, :2]) - np.float32(0.5)) * np.float32(0.04)
    x = (i.ravel() - (nx - 1) / 2) * pitch + jit[:, 0]
    z = (k.ravel() - (nz - 1) / 2) * pitch + jit[:, 1]
    y = floor_y + 0.25 + j.ravel() * pitch
    return np.stack([x, y, z], axis=1).astype(np.float32)


# ---- C1: mapSources/test.json with its actors, from the committed fixtures
def euler_to_quat(e) -> np.ndarray:
    """JPH_Quat_FromEulerAngles: rotate about X, then Y, then Z (q = qz * qy * qx); returns x y z w."""
    hx, hy, hz = (0.5 * float(v) for v in e)
    cx, sx, cy, sy, cz, sz = np.cos(hx), np.sin(hx), np.cos(hy), np.sin(hy), np.cos(hz), np.sin(hz)
    return np.array([cz * sx * cy - sz * cx * sy, cz * cx * sy + sz * sx * cy, sz * cx * cy - cz * sx * sy,
                     cz * cx * cy + sz * sx * sy], dtype=np.float32)


def _qrot(q, v):
    u = q[:3].astype(np.float64)
    t = 2.0 * np.cross(u, v)
    return v + q[3] * t + np.cross(u, t)


LASER_HEIGHT_OFFSET = {0: -0.3, 1: 0.0, 2: 0.3, 3: 0.0}   # floor / middle / ceiling / triple (Laser.c:192-204)


def test_map_scene():
    """The physics content of test.gmap as the engine would create it at load + first tick.

    Returns dict(meshes=[(pos, rot, tris, friction)], bodies=[kwargs for body_desc], lasers=[(pos, quat, mask)],
    names=[actor class per body]).  Body order = actor order in the file (MapLoader.c:74-159), lasers after
    (LaserEmitter.c:59-75).  What is approximated: the player capsule is not created (character controller is a later
    row); leafy.gmdl's two convex hulls are represented by the model's bounding box (ModelLoader.c:152)."""
    z = np.load(os.path.join(GOLDEN, "static_test.npz"), allow_pickle=False)
    models = np.load(os.path.join(GOLDEN, "models.npz"))
    meshes = [(z["mesh_pos"][i], (0.0, 0.0, 0.0, 1.0), z["tris"][z["mesh_start"][i]:z["mesh_start"][i + 1]], 4.25)
              for i in range(len(z["mesh_pos"]))]
    params = {13: dict(width=4.0, height=2.0, depth=4.0)}                      # trigger (actor index 12 in file order)
    laser_height = {3: 1, 4: 0, 5: 2, 6: 3}                                   # 'height' params of the four emitters
    leafy_he = tuple(float(v) for v in models["leafy_bb"][3:])
    leafy_off = tuple(float(v) for v in models["leafy_bb"][:3])
    bodies, names, lasers = [], [], []
    S, K, D = 0, 1, 2
    for ai, (cls, xf) in enumerate(zip(z["actors"], z["actor_xf"])):
        pos, q = xf[:3].astype(np.float32), euler_to_quat(xf[3:])
        base = dict(position=tuple(float(v) for v in pos), rotation=tuple(float(v) for v in q), user_data=ai + 1)
        cls = str(cls)
        if cls == "player" or cls.startswith("global_"):
            continue
        if cls == "prop_coin":
            bodies.append(dict(base, half_extents=(0.25, 0.25, 0.25), motion_type=S, layer=3, is_sensor=1, ray_flags=0))
        elif cls == "prop_goal":
            bodies.append(dict(base, half_extents=(0.5, 0.5, 0.5), motion_type=S, layer=3, is_sensor=1, ray_flags=0))
        elif cls == "trigger":
            p = params.get(ai + 1, dict(width=1.0, height=1.0, depth=1.0))
            bodies.append(dict(base, half_extents=(p["width"] / 2, p["height"] / 2, p["depth"] / 2), motion_type=S, layer=3,
                               is_sensor=1, ray_flags=0))
        elif cls == "prop_model_static":
            c = _qrot(q, np.array(leafy_off))
            bodies.append(dict(base, position=tuple(float(v) for v in pos + c), half_extents=leafy_he, motion_type=S, layer=0,
                               ray_flags=0))
        elif cls == "prop_laser_emitter":
            tris = models["laseremitter_tris"].reshape(-1, 3, 3)
            meshes.append((pos, tuple(float(v) for v in q), tris, 0.2))
            fwd = _qrot(q, np.array([0.0, 0.0, 1.0]))
            lp = pos - fwd * float(models["laseremitter_bb"][5])
            h = laser_height.get(ai, 1)
            lp = lp + np.array([0.0, LASER_HEIGHT_OFFSET[h], 0.0])
            lasers.append((lp.astype(np.float32), q, (1 if h == 3 else 3) | (1 << 8)))
            continue
        elif cls == "prop_physbox":
            bodies.append(dict(base, half_extents=(0.2, 0.2, 0.2), motion_type=D, layer=1, mass=10.0, ray_flags=1))
        elif cls == "test_actor":
            c = _qrot(q, np.array(leafy_off))
            bodies.append(dict(base, position=tuple(float(v) for v in pos + c), half_extents=leafy_he, motion_type=D, layer=1,
                               mass=15.0, allowed_dofs=1 | 2 | 4 | 16, ray_flags=0))
        elif cls == "prop_sprite":
            bodies.append(dict(base, shape=0, motion_type=K, layer=0, ray_flags=0))
        else:
            continue
        names.append(cls)
    for lp, q, _ in lasers:                                                   # the Laser actors' own empty bodies
        bodies.append(dict(position=tuple(float(v) for v in lp), rotation=tuple(float(v) for v in q), shape=0, motion_type=S,
                           layer=0, ray_flags=0))
        names.append("prop_laser")
    return dict(meshes=meshes, bodies=bodies, lasers=lasers, names=names)


def laser_rays(lasers, tmax=50.0, world=0) -> np.ndarray:
    """One ray per laser: origin = the laser body's position, direction = its local -Z (Laser.c:142)."""
    rays = np.zeros(len(lasers), RAY_DTYPE)
    for i, (p, q, mask) in enumerate(lasers):
        rays["origin"][i] = p
        d = _qrot(np.asarray(q, np.float32), np.array([0.0, 0.0, -1.0]))
        rays["dir"][i] = (d / np.linalg.norm(d)).astype(np.float32)
        rays["mask"][i] = mask | (world << 16)
    rays["tmax"] = tmax
    return rays
