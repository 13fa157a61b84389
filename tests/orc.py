"""ctypes binding of the CPU oracle (oracle/liborc.so).  Test infrastructure only."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.environ.get("ORC_DIR", os.path.join(ROOT, "oracle"))
STATIC_BASE = 0x400000
INVALID = 0xFFFFFFFF

SHAPE_EMPTY, SHAPE_BOX, SHAPE_SPHERE = 0, 1, 2
MOTION_STATIC, MOTION_KINEMATIC, MOTION_DYNAMIC = 0, 1, 2
LAYER_STATIC, LAYER_DYNAMIC, LAYER_PLAYER, LAYER_SENSOR = 0, 1, 2, 3
RAYMASK_STATIC = 1
RAYMASK_STATIC_DYNAMIC = 3
RAYMASK_REQUIRE_BLOCKS_LASERS = 1 << 8


class BodyDesc(C.Structure):
    """Layout shared by gpx_body_desc (include/gpx.h) and orc_body_desc (oracle/orc.h)."""
    _fields_ = [
        ("shape", C.c_uint32),
        ("half_extents", C.c_float * 3),
        ("convex_radius", C.c_float),
        ("position", C.c_float * 3),
        ("rotation", C.c_float * 4),
        ("linear_velocity", C.c_float * 3),
        ("angular_velocity", C.c_float * 3),
        ("motion_type", C.c_uint32),
        ("layer", C.c_uint32),
        ("mass", C.c_float),
        ("friction", C.c_float),
        ("restitution", C.c_float),
        ("linear_damping", C.c_float),
        ("angular_damping", C.c_float),
        ("gravity_factor", C.c_float),
        ("is_sensor", C.c_uint32),
        ("allowed_dofs", C.c_uint32),
        ("allow_sleeping", C.c_uint32),
        ("ray_flags", C.c_uint32),
        ("user_data", C.c_uint64),
    ]


CAST_DTYPE = np.dtype([("origin", "<f4", 3), ("tmax", "<f4"), ("dir", "<f4", 3), ("mask", "<u4"), ("radius", "<f4"), ("pad", "<f4", 3)])
CAST_HIT_DTYPE = np.dtype([("fraction", "<f4"), ("body", "<u4"), ("face", "<u4"), ("world", "<u4"), ("normal", "<f4", 3), ("pad", "<f4")])
RAY_DTYPE = np.dtype([("origin", "<f4", 3), ("tmax", "<f4"), ("dir", "<f4", 3), ("mask", "<u4")])
HIT_DTYPE = np.dtype([("fraction", "<f4"), ("body", "<u4"), ("face", "<u4"), ("world", "<u4")])


def body_desc(shape=SHAPE_BOX, half_extents=(0.2, 0.2, 0.2), position=(0, 0, 0), rotation=(0, 0, 0, 1),
              linear_velocity=(0, 0, 0), angular_velocity=(0, 0, 0), motion_type=MOTION_DYNAMIC,
              layer=LAYER_DYNAMIC, mass=10.0, friction=0.2, restitution=0.0, linear_damping=0.05,
              angular_damping=0.05, gravity_factor=1.0, is_sensor=0, allowed_dofs=63, allow_sleeping=0,
              ray_flags=1, user_data=0, convex_radius=0.05) -> BodyDesc:
    """Jolt's BodyCreationSettings defaults + the physbox parameters (game/src/actor/prop/Physbox.c:19-38)."""
    d = BodyDesc()
    d.shape = shape
    d.half_extents[:] = half_extents
    d.convex_radius = convex_radius
    d.position[:] = position
    d.rotation[:] = rotation
    d.linear_velocity[:] = linear_velocity
    d.angular_velocity[:] = angular_velocity
    d.motion_type = motion_type
    d.layer = layer
    d.mass = mass
    d.friction = friction
    d.restitution = restitution
    d.linear_damping = linear_damping
    d.angular_damping = angular_damping
    d.gravity_factor = gravity_factor
    d.is_sensor = is_sensor
    d.allowed_dofs = allowed_dofs
    d.allow_sleeping = allow_sleeping
    d.ray_flags = ray_flags
    d.user_data = user_data
    return d


_lib = None


def build():
    subprocess.run(["make", "-s", "-C", ORACLE_DIR], check=True)


def lib():
    global _lib
    if _lib is not None:
        return _lib
    path = os.path.join(ORACLE_DIR, "liborc.so")
    src = [os.path.join(ORACLE_DIR, f) for f in ("orc.c", "orc.h", "orc_math.h")]
    if not os.path.exists(path) or any(os.path.getmtime(s) > os.path.getmtime(path) for s in src):
        build()
    L = C.CDLL(path)
    L.orc_world_create.restype = C.c_void_p
    L.orc_world_create.argtypes = [C.c_uint32, C.c_uint32, C.POINTER(C.c_float), C.c_uint32, C.c_uint32]
    L.orc_world_destroy.argtypes = [C.c_void_p]
    L.orc_world_set_mode.argtypes = [C.c_void_p, C.c_int]
    L.orc_static_add_mesh.restype = C.c_uint32
    L.orc_static_add_mesh.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_void_p,
                                      C.c_uint64, C.c_float]
    L.orc_static_commit.argtypes = [C.c_void_p]
    L.orc_body_create.restype = C.c_uint32
    L.orc_body_create.argtypes = [C.c_void_p, C.c_void_p]  # same layout as gpx_body_desc
    L.orc_body_destroy.argtypes = [C.c_void_p, C.c_uint32]
    L.orc_body_set_velocity.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    L.orc_body_set_position.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.c_float)]
    L.orc_body_set_ray_flags.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32]
    L.orc_body_wake.argtypes = [C.c_void_p, C.c_uint32]
    L.orc_overlap_capsule.restype = C.c_float
    L.orc_overlap_capsule.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.c_float, C.c_float, C.POINTER(C.c_float), C.POINTER(C.c_uint32)]
    L.orc_body_asleep.argtypes = [C.c_void_p, C.c_uint32]
    L.orc_body_asleep.restype = C.c_uint32
    L.orc_step.restype = C.c_int
    L.orc_step.argtypes = [C.c_void_p, C.c_float, C.c_int]
    L.orc_step_mt.restype = C.c_int
    L.orc_step_mt.argtypes = [C.c_void_p, C.c_float, C.c_int]
    L.orc_body_get.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
    L.orc_character_create.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.c_float, C.c_float, C.c_float]
    L.orc_character_destroy.argtypes = [C.c_void_p]
    L.orc_character_set_velocity.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
    L.orc_character_set_position.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
    L.orc_character_update.argtypes = [C.c_void_p, C.c_float]
    L.orc_character_update_ex.argtypes = [C.c_void_p, C.c_float, C.POINTER(C.c_float * 5)]
    L.orc_character_get.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    L.orc_character_contacts.restype = C.c_uint32
    L.orc_character_contacts.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
    L.orc_events.restype = C.c_uint32
    L.orc_events.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
    L.orc_manifold_count.restype = C.c_uint32
    L.orc_manifold_count.argtypes = [C.c_void_p]
    L.orc_raycast.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
    L.orc_spherecast.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
    L.orc_raycast_mt.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
    L.orc_max_threads.restype = C.c_int
    L.orc_step_many.restype = C.c_int
    L.orc_step_many.argtypes = [C.POINTER(C.c_void_p), C.c_uint32, C.c_float, C.c_int, C.c_int]
    L.orc_static_triangles.restype = C.c_uint32
    L.orc_static_triangles.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]
    _lib = L
    return L


class World:
    """One oracle world."""

    def __init__(self, max_bodies=64, max_manifolds=0, gravity=(0.0, -9.81, 0.0), velocity_steps=0,
                 position_steps=0, wide=None):
        self.L = lib()
        g = (C.c_float * 3)(*gravity)
        self.h = C.c_void_p(self.L.orc_world_create(max_bodies, max_manifolds, g, velocity_steps, position_steps))
        self.max_bodies = max_bodies
        # the product switches to its wide-world kernels (hashed-priority colouring) above 64 bodies per world
        if wide if wide is not None else max_bodies > 64:
            self.L.orc_world_set_mode(self.h, 1)

    def close(self):
        if self.h:
            self.L.orc_world_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def add_mesh(self, pos, tris, friction=4.25, rot=(0, 0, 0, 1)) -> int:
        t = np.ascontiguousarray(tris, dtype=np.float32).reshape(-1, 9)
        p = (C.c_float * 3)(*[float(x) for x in pos])
        r = (C.c_float * 4)(*rot)
        return self.L.orc_static_add_mesh(self.h, p, r, t.ctypes.data, len(t), friction)

    def commit(self):
        self.L.orc_static_commit(self.h)

    def create(self, desc: BodyDesc) -> int:
        return self.L.orc_body_create(self.h, C.byref(desc))

    def destroy(self, bid: int):
        self.L.orc_body_destroy(self.h, bid)

    def set_velocity(self, bid: int, v=None, av=None):
        self.L.orc_body_set_velocity(self.h, bid, (C.c_float * 3)(*v) if v is not None else None,
                                     (C.c_float * 3)(*av) if av is not None else None)

    def set_position(self, bid: int, p):
        self.L.orc_body_set_position(self.h, bid, (C.c_float * 3)(*p))

    def overlap_capsule(self, center, half_height, radius):
        n, b = (C.c_float * 3)(), C.c_uint32()
        d = self.L.orc_overlap_capsule(self.h, (C.c_float * 3)(*[float(v) for v in center]), half_height, radius, n, C.byref(b))
        return np.float32(d), np.array(list(n), np.float32), b.value

    def wake(self, bid: int):
        self.L.orc_body_wake(self.h, bid)

    def asleep(self, n: int) -> np.ndarray:
        return np.array([self.L.orc_body_asleep(self.h, i) for i in range(n)], bool)

    def set_ray_flags(self, bid: int, flags: int):
        self.L.orc_body_set_ray_flags(self.h, bid, flags)

    def step(self, dt=1.0 / 60.0, collision_steps=2) -> int:
        return self.L.orc_step(self.h, dt, collision_steps)

    def step_mt(self, dt=1.0 / 60.0, collision_steps=2) -> int:
        """The same tick over host threads (wide mode); bit-identical to step()."""
        return self.L.orc_step_mt(self.h, dt, collision_steps)

    def get(self, bid: int):
        xf = np.zeros(7, np.float32)
        vel = np.zeros(6, np.float32)
        self.L.orc_body_get(self.h, bid, xf.ctypes.data, vel.ctypes.data)
        return xf, vel

    def state(self, n: int):
        xf = np.zeros((n, 7), np.float32)
        vel = np.zeros((n, 6), np.float32)
        for i in range(n):
            self.L.orc_body_get(self.h, i, xf[i].ctypes.data, vel[i].ctypes.data)
        return xf, vel

    # ---- player character (capsule), PlayerPhysics.c:173-194
    def character_create(self, pos, half_height=0.2, radius=0.25, max_slope_deg=50.0):
        self.L.orc_character_create(self.h, (C.c_float * 3)(*pos), half_height, radius, max_slope_deg)

    def character_set_velocity(self, v):
        self.L.orc_character_set_velocity(self.h, (C.c_float * 3)(*v))

    def character_set_position(self, p):
        self.L.orc_character_set_position(self.h, (C.c_float * 3)(*p))

    def character_update(self, dt=1.0 / 60.0, settings=None):
        """settings: (stick_to_floor_step_down, walk_stairs_step_up, min_step_forward, step_forward_test, cos_angle_forward_contact)"""
        if settings is None:
            self.L.orc_character_update(self.h, dt)
        else:
            self.L.orc_character_update_ex(self.h, dt, C.byref((C.c_float * 5)(*settings)))

    def character_get(self):
        p, v = (C.c_float * 3)(), (C.c_float * 3)()
        g, gb = C.c_uint32(), C.c_uint32()
        self.L.orc_character_get(self.h, p, v, C.byref(g), C.byref(gb))
        return np.array(list(p), np.float32), np.array(list(v), np.float32), g.value, gb.value

    def character_contacts(self) -> np.ndarray:
        out = np.zeros(64, np.uint32)
        n = self.L.orc_character_contacts(self.h, out.ctypes.data, 64)
        return out[:n].copy()

    def events(self) -> np.ndarray:
        """(n, 3) triples a, b, kind of the last tick."""
        n = self.L.orc_events(self.h, None, 0)
        out = np.zeros((n, 3), np.uint32)
        self.L.orc_events(self.h, out.ctypes.data, n)
        return out

    def manifolds(self) -> int:
        return self.L.orc_manifold_count(self.h)

    def spherecast(self, casts: np.ndarray) -> np.ndarray:
        casts = np.ascontiguousarray(casts, dtype=CAST_DTYPE)
        hits = np.zeros(len(casts), CAST_HIT_DTYPE)
        self.L.orc_spherecast(self.h, casts.ctypes.data, len(casts), hits.ctypes.data)
        return hits

    def raycast(self, rays: np.ndarray, mt=False) -> np.ndarray:
        rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
        hits = np.zeros(len(rays), HIT_DTYPE)
        (self.L.orc_raycast_mt if mt else self.L.orc_raycast)(self.h, rays.ctypes.data, len(rays), hits.ctypes.data)
        return hits
