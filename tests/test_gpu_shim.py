"""GPU parity through the joltc-subset shim (SURVEY §8(b) shape B1, §8(f) rank 3).

`tests/shim/engine_calls.c` is engine-style C: it includes <joltc/...>, registers the engine's layer / ray / body filters
and its character listener as callbacks, loads the map the way MapLoader.c does, creates physboxes, a coin, a door and
two lasers the way the game's actors do, and runs MapFixedUpdate's call order for N ticks.  This test replays the same
scene on the oracle, call for call, and requires every printed observation — body transforms, ray results, character
state, the listener's callback stream — to be bit-identical.
"""
import os
import struct
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CHAR = 0x3FFFFF
TICKS = 260


def _f(hexbits: str) -> np.float32:
    return np.array([int(hexbits, 16)], np.uint32).view(np.float32)[0]


def _scene_file(path, scenes, boxes=None):
    meshes = scenes.load_static("stacked")
    pts = np.load(os.path.join(scenes.GOLDEN, "models.npz"))["cube_hull_points"].astype(np.float32)
    boxes = scenes.stack_positions(8).astype(np.float32) if boxes is None else np.asarray(boxes, np.float32)
    with open(path, "wb") as f:
        f.write(struct.pack("<I", len(meshes)))
        for pos, tris in meshes:
            t = np.ascontiguousarray(tris, np.float32).reshape(-1, 9)
            f.write(np.asarray(pos, np.float32).tobytes())
            f.write(struct.pack("<I", len(t)))
            f.write(t.tobytes())
        f.write(struct.pack("<I", len(pts)))
        f.write(pts.tobytes())
        f.write(struct.pack("<I", len(boxes)))
        f.write(boxes.tobytes())
    return meshes, boxes


def _oracle_run(orc, scenes, meshes, boxes, ticks, max_bodies=64):
    """The driver's scene on the oracle, in the driver's call order.  Returns the lines the driver should print."""
    out = []
    o = orc.World(max_bodies)
    o.character_create((-0.2, 0.0, -0.5))
    first_map = None
    for pos, tris in meshes:
        b = o.add_mesh(pos, tris)
        first_map = b if first_map is None else first_map
    o.commit()
    out.append(f"M {len(meshes)} {first_map:08x}")
    out.append("S 1")                                   # cube.gmdl's hull is its 0.4 m box
    common = dict(allow_sleeping=1)
    box_ids = [o.create(orc.body_desc(position=tuple(p), mass=10.0, ray_flags=1, **common)) for p in boxes]
    coin = o.create(orc.body_desc(half_extents=(0.25, 0.25, 0.25), position=(-1.0, -1.25, -0.5), motion_type=0, layer=3,
                                  is_sensor=1, mass=0.0, ray_flags=0, **common))
    door = o.create(orc.body_desc(half_extents=(0.01, 0.5, 0.5), position=(1.5, -1.0, -1.5), motion_type=1, layer=0, mass=1.0,
                                  ray_flags=1, **common))
    out.append("S 0")                                   # the flat 4-point hull is a stand-in slab
    lasers = [o.create(orc.body_desc(shape=0, half_extents=(0, 0, 0), position=p, rotation=q, motion_type=0, layer=0, mass=0.0,
                                     ray_flags=0, **common))
              for p, q in (((0.0, -1.25, 1.0), (0, 0, 0, 1)), ((0.0, -1.0, -3.0), (0, 1, 0, 0)))]
    out.append(f"I coin {coin:08x} door {door:08x} laser {lasers[0]:08x} laser3 {lasers[1]:08x}")
    names = {b: "prop_physbox" for b in box_ids}
    names.update({coin: "prop_coin", door: "prop_door", lasers[0]: "laser", lasers[1]: "laser3"})
    coin_alive = True
    touching = []
    # JPH_ExtendedUpdateSettings as the driver fills it (PlayerPhysics.c:439-446), cos 75 deg by the C library's cosf
    import ctypes
    libm = ctypes.CDLL("libm.so.6")
    libm.cosf.restype = ctypes.c_float
    libm.cosf.argtypes = [ctypes.c_float]
    ext = (0.25, 0.25, 0.02, 0.15, libm.cosf(np.float32(np.float32(75.0) * np.float32(3.14159265358979323846)) / np.float32(180.0)))
    g_step = np.float32(np.float64(np.float32(-9.81)) * (1.0 / 60.0))

    def ray(origin, direction, tmax, mask):
        r = np.zeros(1, orc.RAY_DTYPE)
        r["origin"][0] = origin
        r["dir"][0] = direction
        r["tmax"] = tmax
        r["mask"] = mask
        return o.raycast(r)[0]

    def bits(x):
        return f"{np.array([x], np.float32).view(np.uint32)[0]:08x}"

    for tick in range(1, ticks + 1):
        move = np.zeros(3, np.float32)
        if tick > 40:
            move[0] = -1.5 if tick <= 140 else 1.5
        _, vel, ground, _ = o.character_get()
        if ground != 0:
            move[1] = np.float32(move[1] + np.float32(vel[1] + g_step))
        o.character_set_velocity([float(x) for x in move])
        h = ray((0.0, -1.25, 2.0), (0, 0, -1), 10.0, 0b0011)
        if tick % 20 == 1:
            hit = int(h["body"] != 0xFFFFFFFF)
            out.append(f"R {tick} camera {hit} {int(h['body']) if hit else 0:08x} {bits(h['fraction']) if hit else bits(0.0)} "
                       f"{int(h['face']) if hit else 0:08x}")
        o.character_update(settings=ext)
        pos = o.character_get()[0]
        # the listener's callbacks run inside ExtendedUpdate: contacts that exist now (added / persisted) by id, bodies
        # before map meshes, then the ones that ended; the coin's handler removes it before this tick's Update
        now = [int(x) for x in o.character_contacts()]
        for other in now:
            if other in touching:
                out.append(f"E {tick} persisted {other:08x}")
            else:
                out.append(f"E {tick} added {other:08x} {names.get(other, '-')}")
                if other == coin and coin_alive:
                    o.destroy(coin)
                    coin_alive = False
        for other in sorted(touching, key=lambda i: (i >= 0x400000, i)):
            if other not in now:
                out.append(f"E {tick} removed {other:08x}")
        touching = now
        if tick == 60:
            o.set_velocity(door, (0.0, 0.0, 1.0))
        if tick == 120:
            o.set_velocity(door, (0.0, 0.0, 0.0))
            o.set_position(door, (1.5, -1.0, -0.5))
        if tick == 200:
            o.set_ray_flags(box_ids[0], 0)
        for i, (origin, d, mask) in enumerate((((0.0, -1.25, 1.0), (0, 0, -1), 0b0011 | 0x100), ((0.0, -1.0, -3.0), (0, 0, 1), 0b0001 | 0x100))):
            h = ray(origin, d, 50.0, mask)
            if tick % 20 == 1 or tick == 200:
                hit = int(h["body"] != 0xFFFFFFFF)
                off = np.float32(d[2]) * (np.float32(h["fraction"]) * np.float32(50.0))
                out.append(f"R {tick} {'laser' if i == 0 else 'laser3'} {hit} {int(h['body']) if hit else 0:08x} "
                           f"{bits(h['fraction']) if hit else bits(0.0)} {int(h['face']) if hit else 0:08x} {bits(off) if hit else bits(0.0)}")
        assert o.step() == 0
        if tick % 20 == 0 or tick == ticks:
            for b in box_ids + [door]:
                xf, _ = o.get(b)
                out.append(f"X {tick} {'box' if b != door else 'door'} {b:08x} " + " ".join(bits(v) for v in xf))
            p, v, ground, _ = o.character_get()
            out.append(f"C {tick} " + " ".join(bits(x) for x in pos) + " " + " ".join(bits(x) for x in v) + f" {ground}")
    xf, _ = o.get(box_ids[0])
    out.append(f"W {bits(xf[0])} {bits(xf[1])} {bits(xf[2])} {bits(1.0)}")
    out.append("done")
    return out, dict(coin=coin, door=door, boxes=box_ids, first_map=first_map)


@pytest.mark.parametrize("max_bodies", [64, 128, 10240], ids=["ensemble-kernel world", "wide-world kernels", "joltc default capacity"])
def test_engine_call_sequence_through_the_shim_matches_the_oracle(orc, scenes, tmp_path, max_bodies):
    """The same engine-style run twice: in the default 64-slot world (the fused ensemble kernel) and, with
    GPX_MAX_BODIES=128, in a wide world (sort-and-sweep, islands, per-tick event kernels) against the oracle's wide mode."""
    import shim_build
    driver = shim_build.build_driver()
    meshes, boxes = _scene_file(tmp_path / "scene.bin", scenes)
    env = dict(os.environ, GPX_MAX_BODIES=str(max_bodies))
    if max_bodies == 10240:
        env.pop("GPX_MAX_BODIES")       # what the shim does when nobody says anything: joltc's default of 10240 bodies
    r = subprocess.run([driver, str(tmp_path / "scene.bin"), str(TICKS)], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, f"driver exit {r.returncode}\n{r.stdout[-2000:]}\n{r.stderr[-2000:]}"
    got = r.stdout.strip().splitlines()
    want, ids = _oracle_run(orc, scenes, meshes, boxes, TICKS, max_bodies)
    for i, (g, w) in enumerate(zip(got, want)):
        assert g == w, f"line {i}: shim printed\n  {g}\noracle says\n  {w}"
    assert len(got) == len(want)

    # the scenario did what it was built to show
    text = "\n".join(got)
    assert f"added {ids['coin']:08x} prop_coin" in text                      # the coin was picked up through the listener ...
    coin_lines = [l for l in got if l.startswith("E") and f"{ids['coin']:08x}" in l]
    assert [l.split()[2] for l in coin_lines] == ["added", "removed"]       # ... and destroyed inside the callback: never persisted
    laser = {int(l.split()[1]): l.split() for l in got if l.startswith("R") and l.split()[2] == "laser"}
    assert int(laser[181][4], 16) == ids["boxes"][0]                         # the beam stops on the bottom box ...
    assert int(laser[200][4], 16) != ids["boxes"][0]                         # ... until that actor loses CAN_BLOCK_LASERS
    triple = [l.split() for l in got if l.startswith("R") and l.split()[2] == "laser3"]
    assert all(int(t[4], 16) >= 0x400000 for t in triple if t[3] == "1")     # the triple laser only ever sees map geometry
    door = {int(l.split()[1]): l.split() for l in got if l.startswith("X") and l.split()[2] == "door"}
    assert abs(_f(door[100][6]) - (-1.5 + 41 / 60)) < 1e-4                   # kinematic: 41 ticks at 1 m/s after tick 60
    assert _f(door[140][6]) == np.float32(-0.5)                            # snapped open by SetPosition at tick 120


@pytest.mark.parametrize("seed,max_bodies", [(1, 64), (2, 64), (3, 128)])
def test_shim_with_a_loose_pile_of_physboxes_matches_the_oracle(orc, scenes, tmp_path, seed, max_bodies):
    """The same engine-style driver with the physboxes thrown in as a loose, overlapping pile next to the player instead of
    the tidy column (boxes shove each other, the player, the door and the laser beams): every printed observation —
    transforms, rays, character state, listener callbacks — line by line."""
    import shim_build
    driver = shim_build.build_driver()
    rng = np.random.default_rng(seed)
    boxes = np.stack([rng.uniform(-1.2, 1.2, 14), rng.uniform(-1.0, 1.5, 14), rng.uniform(-2.4, -0.4, 14)], axis=1)
    meshes, boxes = _scene_file(tmp_path / "scene.bin", scenes, boxes)
    env = dict(os.environ, GPX_MAX_BODIES=str(max_bodies))
    r = subprocess.run([driver, str(tmp_path / "scene.bin"), str(TICKS)], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, f"driver exit {r.returncode}\n{r.stdout[-2000:]}\n{r.stderr[-2000:]}"
    got = r.stdout.strip().splitlines()
    want, _ = _oracle_run(orc, scenes, meshes, boxes, TICKS, max_bodies)
    for i, (g, w) in enumerate(zip(got, want)):
        assert g == w, f"line {i}: shim printed\n  {g}\noracle says\n  {w}"
    assert len(got) == len(want)


def test_shim_rejects_foreign_layer_tables_and_reports_inexact_shapes(tmp_path):
    """The collision matrix is compiled into the kernels: a different ObjectLayerPairFilter must be refused, loudly."""
    import ctypes as C
    import shim_build
    L = C.CDLL(shim_build.build_shim())
    L.JPH_Init.restype = C.c_bool
    assert L.JPH_Init()
    pair_cb = C.CFUNCTYPE(C.c_bool, C.c_uint32, C.c_uint32)
    everything = pair_cb(lambda a, b: True)

    class PairImpl(C.Structure):
        _fields_ = [("ShouldCollide", pair_cb)]

    class Settings(C.Structure):
        _fields_ = [("maxBodies", C.c_uint32), ("numBodyMutexes", C.c_uint32), ("maxBodyPairs", C.c_uint32),
                    ("maxContactConstraints", C.c_uint32), ("_padding", C.c_uint32), ("bpi", C.c_void_p), ("olpf", C.c_void_p),
                    ("ovbpf", C.c_void_p)]

    L.JPH_ObjectLayerPairFilter_Create.restype = C.c_void_p
    L.JPH_PhysicsSystem_Create.restype = C.c_void_p
    L.JPH_PhysicsSystem_Create.argtypes = [C.POINTER(Settings)]
    s = Settings()
    s.olpf = L.JPH_ObjectLayerPairFilter_Create(C.byref(PairImpl(everything)))
    assert L.JPH_PhysicsSystem_Create(C.byref(s)) is None
    s.olpf = None
    sys_ = L.JPH_PhysicsSystem_Create(C.byref(s))
    assert sys_ is not None
    L.JPH_CylinderShape_Create.restype = C.c_void_p
    L.JPH_CylinderShape_Create.argtypes = [C.c_float, C.c_float]
    L.JPH_GPX_ShapeIsExact.argtypes = [C.c_void_p]
    cyl = L.JPH_CylinderShape_Create(0.5, 0.25)                               # NpcJohn.c:29: a bounding box stands in
    assert L.JPH_GPX_ShapeIsExact(cyl) == 0
    L.JPH_Shape_Destroy.argtypes = [C.c_void_p]
    L.JPH_Shape_Destroy(cyl)

    # body ids carry a sequence number: a reused slot gets a new id and the old one goes dead
    class V3(C.Structure):
        _fields_ = [("x", C.c_float), ("y", C.c_float), ("z", C.c_float)]

    class Xfm(C.Structure):
        _fields_ = [("p", V3), ("q", C.c_float * 4)]

    L.JPH_PhysicsSystem_GetBodyInterface.restype = C.c_void_p
    L.JPH_PhysicsSystem_GetBodyInterface.argtypes = [C.c_void_p]
    L.JPH_BoxShape_Create.restype = C.c_void_p
    L.JPH_BoxShape_Create.argtypes = [C.POINTER(V3), C.c_float]
    L.JPH_BodyCreationSettings_Create2_GAME.restype = C.c_void_p
    L.JPH_BodyCreationSettings_Create2_GAME.argtypes = [C.c_void_p, C.POINTER(Xfm), C.c_int, C.c_uint32, C.c_void_p]
    L.JPH_BodyCreationSettings_Destroy.argtypes = [C.c_void_p]
    L.JPH_BodyInterface_CreateAndAddBody.restype = C.c_uint32
    L.JPH_BodyInterface_CreateAndAddBody.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    L.JPH_BodyInterface_RemoveAndDestroyBody.argtypes = [C.c_void_p, C.c_uint32]
    L.JPH_BodyInterface_GetPosition.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(V3)]
    L.JPH_BodyInterface_GetUserData.restype = C.c_uint64
    L.JPH_BodyInterface_GetUserData.argtypes = [C.c_void_p, C.c_uint32]
    bi = L.JPH_PhysicsSystem_GetBodyInterface(sys_)
    box = L.JPH_BoxShape_Create(C.byref(V3(0.2, 0.2, 0.2)), 0.05)

    def make(x, tag):
        st = L.JPH_BodyCreationSettings_Create2_GAME(box, C.byref(Xfm(V3(x, 5.0, 0.0), (C.c_float * 4)(0, 0, 0, 1))), 2, 1, tag)
        b = L.JPH_BodyInterface_CreateAndAddBody(bi, st, 0)
        L.JPH_BodyCreationSettings_Destroy(st)
        return b

    a = make(1.0, 0x1111)
    assert a == 0 and L.JPH_BodyInterface_GetUserData(bi, a) == 0x1111
    L.JPH_BodyInterface_RemoveAndDestroyBody(bi, a)
    b = make(2.0, 0x2222)
    assert b != a and (b & 0x7FFFFF) == (a & 0x7FFFFF) and (b >> 24) == 1      # same slot, next sequence number
    out = V3(-1.0, -1.0, -1.0)
    L.JPH_BodyInterface_GetPosition(bi, a, C.byref(out))                       # the stale id finds nothing
    assert (out.x, out.y, out.z) == (-1.0, -1.0, -1.0) and L.JPH_BodyInterface_GetUserData(bi, a) == 0
    L.JPH_BodyInterface_GetPosition(bi, b, C.byref(out))
    assert (out.x, out.y, out.z) == (2.0, 5.0, 0.0) and L.JPH_BodyInterface_GetUserData(bi, b) == 0x2222
    L.JPH_BodyInterface_RemoveAndDestroyBody(bi, a)                            # and cannot destroy the new tenant
    assert L.JPH_BodyInterface_GetUserData(bi, b) == 0x2222
    L.JPH_Shape_Destroy(box)
    L.JPH_PhysicsSystem_Destroy.argtypes = [C.c_void_p]
    L.JPH_PhysicsSystem_Destroy(sys_)
    L.JPH_Shutdown()
