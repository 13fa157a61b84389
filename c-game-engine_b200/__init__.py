"""B200 physics-tick / ray-query backend for NBT22/c-game-engine — Python view of the C ABI (include/gpx.h).

This module is a thin ctypes binding used by the tests, bench.py and the multi-GPU harness.  The product is
`libgpx.so` (hand-written sm_100a kernels behind a plain C ABI); there is no CPU implementation here and the
import fails loudly when the library has not been built.
"""
from __future__ import annotations

import ctypes as C
import os
import re
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB_PATH = os.path.join(HERE, "libgpx.so")
HEADER = os.path.join(ROOT, "include", "gpx.h")

STATIC_BODY_BASE = 0x400000
INVALID_BODY = 0xFFFFFFFF

SHAPE_EMPTY, SHAPE_BOX, SHAPE_SPHERE = 0, 1, 2
MOTION_STATIC, MOTION_KINEMATIC, MOTION_DYNAMIC = 0, 1, 2
LAYER_STATIC, LAYER_DYNAMIC, LAYER_PLAYER, LAYER_SENSOR = 0, 1, 2, 3
RAYMASK_STATIC = 1
RAYMASK_STATIC_DYNAMIC = 3
RAYMASK_REQUIRE_BLOCKS_LASERS = 1 << 8
ERR_CUDA = 32

CAST_DTYPE = np.dtype([("origin", "<f4", 3), ("tmax", "<f4"), ("dir", "<f4", 3), ("mask", "<u4"), ("radius", "<f4"), ("pad", "<f4", 3)])
CAST_HIT_DTYPE = np.dtype([("fraction", "<f4"), ("body", "<u4"), ("face", "<u4"), ("world", "<u4"), ("normal", "<f4", 3), ("pad", "<f4")])
RAY_DTYPE = np.dtype([("origin", "<f4", 3), ("tmax", "<f4"), ("dir", "<f4", 3), ("mask", "<u4")])
HIT_DTYPE = np.dtype([("fraction", "<f4"), ("body", "<u4"), ("face", "<u4"), ("world", "<u4")])
TRANSFORM_DTYPE = np.dtype([("position", "<f4", 3), ("rotation", "<f4", 4)])
EVENT_DTYPE = np.dtype([("world", "<u4"), ("body_a", "<u4"), ("body_b", "<u4"), ("kind", "<u4")])
STATS_DTYPE = np.dtype([("kinetic_energy", "<f4"), ("max_speed", "<f4"), ("awake_bodies", "<u4"),
                        ("manifolds", "<u4"), ("position_checksum", "<u8"), ("ticks", "<u4"), ("error", "<u4")])


class WorldConfig(C.Structure):
    _fields_ = [
        ("worlds", C.c_uint32),
        ("max_bodies_per_world", C.c_uint32),
        ("max_manifolds_per_world", C.c_uint32),
        ("max_static_triangles", C.c_uint32),
        ("gravity", C.c_float * 3),
        ("device", C.c_int32),
        ("velocity_steps", C.c_uint32),
        ("position_steps", C.c_uint32),
        ("flags", C.c_uint32),
    ]


class BodyDesc(C.Structure):
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


class ModelCollision(C.Structure):
    """gpx_model_collision (include/gpx.h)."""
    pass


class HullShape(C.Structure):
    _fields_ = [("shape", C.c_uint32), ("half_extents", C.c_float * 3), ("center", C.c_float * 3), ("exact", C.c_uint32)]


ModelCollision._fields_ = [("collision_type", C.c_uint32), ("bb_origin", C.c_float * 3), ("bb_extents", C.c_float * 3),
                           ("n_hulls", C.c_uint32), ("n_triangles", C.c_uint64), ("hull_points", C.c_uint64 * 8),
                           ("hull", HullShape * 8), ("exact", C.c_uint32)]


class CharacterDesc(C.Structure):
    _fields_ = [("half_height", C.c_float), ("radius", C.c_float), ("max_slope_deg", C.c_float), ("mass", C.c_float),
                ("position", C.c_float * 3)]


class CharacterState(C.Structure):
    _fields_ = [("position", C.c_float * 3), ("linear_velocity", C.c_float * 3), ("ground_normal", C.c_float * 3),
                ("ground_state", C.c_uint32), ("ground_body", C.c_uint32)]


class Transform(C.Structure):
    _fields_ = [("position", C.c_float * 3), ("rotation", C.c_float * 4)]


CAPSULE_DTYPE = np.dtype([("center", "<f4", 3), ("half_height", "<f4"), ("radius", "<f4"), ("world", "<u4"), ("reserved", "<u4", 2)])
OVERLAP_DTYPE = np.dtype([("depth", "<f4"), ("normal", "<f4", 3), ("body", "<u4"), ("world", "<u4"), ("reserved", "<u4", 2)])
FIXED_UPDATE_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_double)          # gpx_fixed_update_fn
INPUT_EVENT_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_uint64)  # gpx_input_event_fn


def body_desc(shape=SHAPE_BOX, half_extents=(0.2, 0.2, 0.2), position=(0, 0, 0), rotation=(0, 0, 0, 1),
              linear_velocity=(0, 0, 0), angular_velocity=(0, 0, 0), motion_type=MOTION_DYNAMIC,
              layer=LAYER_DYNAMIC, mass=10.0, friction=0.2, restitution=0.0, linear_damping=0.05,
              angular_damping=0.05, gravity_factor=1.0, is_sensor=0, allowed_dofs=63, allow_sleeping=0,
              ray_flags=1, user_data=0, convex_radius=0.05) -> BodyDesc:
    """Jolt BodyCreationSettings defaults with the physbox parameters (game/src/actor/prop/Physbox.c:19-38)."""
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


def build(verbose: bool = False) -> None:
    """Compile libgpx.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-C", os.path.join(HERE, "csrc"), "-j4"], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout[-4000:], r.stderr[-4000:])
    if r.returncode != 0:
        raise RuntimeError("building libgpx.so failed")


def declared_symbols() -> list[str]:
    """Every function the public header declares."""
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gpx_[a-z0-9_]+)\s*\(", src)))


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, u32, u64, f32, i32 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_float, C.c_int
    sig = {
        "gpx_init": (i32, [i32]),
        "gpx_shutdown": (None, []),
        "gpx_last_error": (C.c_char_p, []),
        "gpx_world_create": (vp, [C.POINTER(WorldConfig)]),
        "gpx_world_destroy": (None, [vp]),
        "gpx_static_add_mesh": (i32, [vp, C.POINTER(Transform), vp, u64, f32, u64, C.POINTER(u32)]),
        "gpx_thread_init": (i32, [vp]),
        "gpx_thread_set_function": (None, [vp]),
        "gpx_thread_queue_input_event": (None, [vp, u64]),
        "gpx_thread_set_input_handler": (None, [vp]),
        "gpx_thread_terminate": (None, []),
        "gpx_thread_lock_tick_mutex": (None, []),
        "gpx_thread_unlock_tick_mutex": (None, []),
        "gpx_thread_set_pinned_delta": (None, [i32]),
        "gpx_thread_frame": (u64, []),
        "gpx_thread_last_tick_ns": (u64, []),
        "gpx_static_commit": (i32, [vp]),
        "gpx_static_remove_mesh": (i32, [vp, u32]),
        "gpx_static_load_gmap": (i32, [vp, vp, u64]),
        "gpx_static_load_gmap_container": (i32, [vp, vp, u64]),
        "gpx_static_load_gmap_file": (i32, [vp, C.c_char_p]),
        "gpx_static_info": (i32, [vp, C.POINTER(u32), C.POINTER(u32), C.POINTER(u32)]),
        "gpx_shape_from_hull": (i32, [vp, u64, f32, C.POINTER(HullShape)]),
        "gpx_model_load_gmdl": (i32, [vp, u64, f32, C.POINTER(ModelCollision)]),
        "gpx_model_load_gmdl_container": (i32, [vp, u64, f32, C.POINTER(ModelCollision)]),
        "gpx_static_add_gmdl": (i32, [vp, C.POINTER(Transform), vp, u64, f32, u32]),
        "gpx_body_create": (u32, [vp, u32, C.POINTER(BodyDesc)]),
        "gpx_body_create_all": (i32, [vp, C.POINTER(BodyDesc), u32, vp, vp, vp]),
        "gpx_body_destroy": (i32, [vp, u32, u32]),
        "gpx_body_set_ray_flags": (i32, [vp, u32, u32, u32]),
        "gpx_body_wake": (i32, [vp, u32, u32]),
        "gpx_overlap_capsule_batch": (i32, [vp, vp, u64, vp]),
        "gpx_spherecast_batch": (i32, [vp, vp, u64, vp]),
        "gpx_read_sleeping": (i32, [vp, vp, u64]),
        "gpx_body_set_linear_velocity": (i32, [vp, u32, u32, C.POINTER(f32)]),
        "gpx_body_set_linear_and_angular_velocity": (i32, [vp, u32, u32, C.POINTER(f32), C.POINTER(f32)]),
        "gpx_body_set_position": (i32, [vp, u32, u32, C.POINTER(f32), i32]),
        "gpx_body_set_rotation": (i32, [vp, u32, u32, C.POINTER(f32), i32]),
        "gpx_body_get_transform": (i32, [vp, u32, u32, C.POINTER(Transform)]),
        "gpx_body_get_world_matrix": (i32, [vp, u32, u32, C.POINTER(f32)]),
        "gpx_body_get_velocity": (i32, [vp, u32, u32, C.POINTER(f32), C.POINTER(f32)]),
        "gpx_body_get_user_data": (u64, [vp, u32, u32]),
        "gpx_body_is_active": (i32, [vp, u32, u32]),
        "gpx_step": (i32, [vp, f32, i32]),
        "gpx_sync_transforms": (i32, [vp]),
        "gpx_read_transforms": (i32, [vp, vp, u64]),
        "gpx_read_velocities": (i32, [vp, vp, u64]),
        "gpx_read_stats": (i32, [vp, vp]),
        "gpx_raycast_batch": (i32, [vp, vp, u64, vp]),
        "gpx_raycast_batch_device": (i32, [vp, vp, u64, vp]),
        "gpx_raycast_batch_async": (i32, [vp, vp, u64, vp]),
        "gpx_raycast_transform": (i32, [vp, u32, C.POINTER(Transform), f32, u32, vp]),
        "gpx_device_alloc": (vp, [u64]),
        "gpx_device_free": (None, [vp]),
        "gpx_host_alloc": (vp, [u64]),
        "gpx_host_free": (None, [vp]),
        "gpx_memcpy_h2d": (i32, [vp, vp, u64]),
        "gpx_memcpy_d2h": (i32, [vp, vp, u64]),
        "gpx_device_sync": (i32, [vp]),
        "gpx_timer_begin": (i32, [vp]),
        "gpx_timer_end": (f32, [vp]),
        "gpx_launch_count": (u64, []),
        "gpx_debug_phase_cycles": (i32, [vp, i32, vp]),
        "gpx_debug_wide_counters": (i32, [vp, vp]),
        "gpx_character_create": (i32, [vp, u32, C.POINTER(CharacterDesc)]),
        "gpx_character_destroy": (i32, [vp, u32]),
        "gpx_character_set_linear_velocity": (i32, [vp, u32, C.POINTER(f32)]),
        "gpx_character_set_position": (i32, [vp, u32, C.POINTER(f32)]),
        "gpx_character_update": (i32, [vp, f32]),
        "gpx_character_update_ex": (i32, [vp, f32, C.POINTER(C.c_float * 5)]),
        "gpx_character_get": (i32, [vp, u32, C.POINTER(CharacterState)]),
        "gpx_character_contacts": (i32, [vp, u32, vp, u32, C.POINTER(C.c_uint32)]),
        "gpx_events_enable": (i32, [vp, i32]),
        "gpx_poll_events": (i32, [vp, vp, u64, C.POINTER(u64)]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


class GpxError(RuntimeError):
    pass


def _check(rc: int, what: str):
    if rc != 0:
        raise GpxError(f"{what} failed with gpx_error {rc}: {lib().gpx_last_error().decode()}")


class World:
    """An ensemble of `worlds` independent world instances sharing one static map (gpx_world)."""

    def __init__(self, worlds=1, max_bodies=8, max_manifolds=0, gravity=(0.0, -9.81, 0.0), device=0,
                 velocity_steps=0, position_steps=0, wide=False):
        L = lib()
        rc = L.gpx_init(device)
        if rc < 0:
            raise GpxError(f"gpx_init({device}) failed ({rc}): {L.gpx_last_error().decode()} — no CPU fallback exists")
        cfg = WorldConfig()
        cfg.worlds = worlds
        cfg.max_bodies_per_world = max_bodies
        cfg.max_manifolds_per_world = max_manifolds
        cfg.max_static_triangles = 0
        cfg.gravity[:] = gravity
        cfg.device = device
        cfg.velocity_steps = velocity_steps
        cfg.position_steps = position_steps
        cfg.flags = 1 if wide else 0        # GPX_WORLD_WIDE
        self.L = L
        self.worlds = worlds
        self.max_bodies = max_bodies
        self.h = C.c_void_p(L.gpx_world_create(C.byref(cfg)))
        if not self.h:
            raise GpxError(f"gpx_world_create failed: {L.gpx_last_error().decode()}")

    def close(self):
        if getattr(self, "h", None):
            self.L.gpx_world_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- static geometry
    def add_mesh(self, pos, tris, friction=4.25, rot=(0, 0, 0, 1), user_data=0) -> int:
        t = np.ascontiguousarray(tris, dtype=np.float32).reshape(-1, 9)
        x = Transform()
        x.position[:] = [float(v) for v in pos]
        x.rotation[:] = rot
        out = C.c_uint32(0)
        _check(self.L.gpx_static_add_mesh(self.h, C.byref(x), t.ctypes.data, len(t), friction, user_data,
                                          C.byref(out)), "gpx_static_add_mesh")
        return out.value

    def add_gmdl(self, pos, body: bytes, friction=4.25, rot=(0, 0, 0, 1), ray_flags=1) -> int:
        """gpx_static_add_gmdl: a static model's triangle mesh (decompressed .gmdl) as one static body."""
        x = Transform()
        x.position[:] = [float(v) for v in pos]
        x.rotation[:] = rot
        buf = np.frombuffer(body, dtype=np.uint8)
        rc = self.L.gpx_static_add_gmdl(self.h, C.byref(x), buf.ctypes.data, len(buf), friction, ray_flags)
        if rc < 0:
            raise GpxError(f"gpx_static_add_gmdl failed with {-rc}")
        return STATIC_BODY_BASE + rc

    def load_gmap(self, body: bytes) -> int:
        buf = np.frombuffer(body, dtype=np.uint8)
        rc = self.L.gpx_static_load_gmap(self.h, buf.ctypes.data, len(buf))
        if rc < 0:
            raise GpxError(f"gpx_static_load_gmap failed ({rc})")
        return rc

    def load_gmap_file(self, path: str) -> int:
        """A .gmap asset file: container header + gzip + map layout, parsed by the library (AssetReader.c, MapLoader.c)."""
        rc = self.L.gpx_static_load_gmap_file(self.h, path.encode())
        if rc < 0:
            raise GpxError(f"gpx_static_load_gmap_file({path}) failed ({rc})")
        return rc

    def commit(self):
        _check(self.L.gpx_static_commit(self.h), "gpx_static_commit")

    def remove_mesh(self, body):
        """Drop one static mesh body (RemoveAndDestroyBody on map / static-model geometry); takes effect at the next commit."""
        _check(self.L.gpx_static_remove_mesh(self.h, body), "gpx_static_remove_mesh")

    def static_info(self):
        a, b, c = C.c_uint32(), C.c_uint32(), C.c_uint32()
        self.L.gpx_static_info(self.h, C.byref(a), C.byref(b), C.byref(c))
        return a.value, b.value, c.value

    # ---- bodies
    def create(self, desc: BodyDesc, world=0) -> int:
        return self.L.gpx_body_create(self.h, world, C.byref(desc))

    def create_all(self, descs, linvel=None, angvel=None):
        arr = (BodyDesc * len(descs))(*descs)
        ids = np.zeros(len(descs), np.uint32)
        lv = None if linvel is None else np.ascontiguousarray(linvel, np.float32)
        av = None if angvel is None else np.ascontiguousarray(angvel, np.float32)
        _check(self.L.gpx_body_create_all(self.h, arr, len(descs), None if lv is None else lv.ctypes.data,
                                          None if av is None else av.ctypes.data, ids.ctypes.data),
               "gpx_body_create_all")
        return ids

    def destroy(self, body, world=0):
        _check(self.L.gpx_body_destroy(self.h, world, body), "gpx_body_destroy")

    def wide_counters(self):
        """Counters of the last sub-step of a wide world (profiling / test aid)."""
        c = np.zeros(8, np.uint32)
        _check(self.L.gpx_debug_wide_counters(self.h, c.ctypes.data), "gpx_debug_wide_counters")
        return dict(manifold_slots=int(c[0]), small_islands=int(c[1]), colours=int(c[2]), error=int(c[3]),
                    medium_islands=int(c[4]), large_island_manifolds=int(c[5]))

    def overlap_capsules(self, queries: np.ndarray) -> np.ndarray:
        """gpx_overlap_capsule_batch: deepest penetration of each upright capsule against the map and the solid bodies."""
        q = np.ascontiguousarray(queries, dtype=CAPSULE_DTYPE)
        out = np.zeros(len(q), OVERLAP_DTYPE)
        _check(self.L.gpx_overlap_capsule_batch(self.h, q.ctypes.data, len(q), out.ctypes.data), "gpx_overlap_capsule_batch")
        return out

    def wake(self, body, world=0):
        _check(self.L.gpx_body_wake(self.h, world, body), "gpx_body_wake")

    def sleeping(self) -> np.ndarray:
        """(worlds, max_bodies) bool: which bodies are asleep."""
        out = np.zeros(self.worlds * self.max_bodies, np.uint8)
        _check(self.L.gpx_read_sleeping(self.h, out.ctypes.data, out.size), "gpx_read_sleeping")
        return out.reshape(self.worlds, self.max_bodies).astype(bool)

    def set_ray_flags(self, body, flags, world=0):
        _check(self.L.gpx_body_set_ray_flags(self.h, world, body, flags), "gpx_body_set_ray_flags")

    def set_velocity(self, body, v, av=None, world=0):
        fv = (C.c_float * 3)(*v)
        if av is None:
            _check(self.L.gpx_body_set_linear_velocity(self.h, world, body, fv), "set_linear_velocity")
        else:
            fa = (C.c_float * 3)(*av)
            _check(self.L.gpx_body_set_linear_and_angular_velocity(self.h, world, body, fv, fa), "set_velocity")

    def set_position(self, body, p, world=0):
        _check(self.L.gpx_body_set_position(self.h, world, body, (C.c_float * 3)(*p), 1), "set_position")

    def set_rotation(self, body, q, world=0):
        _check(self.L.gpx_body_set_rotation(self.h, world, body, (C.c_float * 4)(*q), 1), "set_rotation")

    def get_transform(self, body, world=0):
        t = Transform()
        _check(self.L.gpx_body_get_transform(self.h, world, body, C.byref(t)), "get_transform")
        return np.array(list(t.position) + list(t.rotation), np.float32)

    def user_data(self, body, world=0) -> int:
        return self.L.gpx_body_get_user_data(self.h, world, body)

    # ---- tick
    def step(self, dt=1.0 / 60.0, collision_steps=2) -> int:
        rc = self.L.gpx_step(self.h, dt, collision_steps)
        if rc >= ERR_CUDA:
            raise GpxError(f"gpx_step failed ({rc}): {self.L.gpx_last_error().decode()}")
        return rc

    def sync(self) -> int:
        rc = self.L.gpx_sync_transforms(self.h)
        if rc >= ERR_CUDA:
            raise GpxError(f"gpx_sync_transforms failed ({rc}): {self.L.gpx_last_error().decode()}")
        return rc

    def transforms(self) -> np.ndarray:
        """(worlds, max_bodies, 7) pos + quat, after waiting for the stream."""
        out = np.zeros(self.worlds * self.max_bodies, TRANSFORM_DTYPE)
        rc = self.L.gpx_read_transforms(self.h, out.ctypes.data, len(out))
        if rc >= ERR_CUDA:
            raise GpxError(f"gpx_read_transforms failed ({rc}): {self.L.gpx_last_error().decode()}")
        flat = np.concatenate([out["position"], out["rotation"]], axis=1)
        return flat.reshape(self.worlds, self.max_bodies, 7)

    def velocities(self) -> np.ndarray:
        out = np.zeros((self.worlds * self.max_bodies, 6), np.float32)
        _check(self.L.gpx_read_velocities(self.h, out.ctypes.data, len(out)), "gpx_read_velocities")
        return out.reshape(self.worlds, self.max_bodies, 6)

    def stats(self) -> np.ndarray:
        out = np.zeros(self.worlds, STATS_DTYPE)
        _check(self.L.gpx_read_stats(self.h, out.ctypes.data), "gpx_read_stats")
        return out

    # ---- rays
    def raycast(self, rays: np.ndarray) -> np.ndarray:
        rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
        hits = np.zeros(len(rays), HIT_DTYPE)
        _check(self.L.gpx_raycast_batch(self.h, rays.ctypes.data, len(rays), hits.ctypes.data), "gpx_raycast_batch")
        return hits

    def spherecast(self, casts: np.ndarray) -> np.ndarray:
        """gpx_spherecast_batch: first contact of each swept sphere with the map and the bodies its mask admits."""
        casts = np.ascontiguousarray(casts, dtype=CAST_DTYPE)
        hits = np.zeros(len(casts), CAST_HIT_DTYPE)
        _check(self.L.gpx_spherecast_batch(self.h, casts.ctypes.data, len(casts), hits.ctypes.data), "gpx_spherecast_batch")
        return hits

    def raycast_into(self, rays: np.ndarray, hits: np.ndarray) -> None:
        """gpx_raycast_batch on caller-owned host arrays (e.g. pinned ones from pinned_array): no allocation."""
        assert rays.dtype == RAY_DTYPE and hits.dtype == HIT_DTYPE and len(hits) >= len(rays)
        _check(self.L.gpx_raycast_batch(self.h, rays.ctypes.data, len(rays), hits.ctypes.data), "gpx_raycast_batch")

    def raycast_into_async(self, rays: np.ndarray, hits: np.ndarray) -> None:
        """gpx_raycast_batch_async on pinned arrays: hits are valid after the next sync()."""
        assert rays.dtype == RAY_DTYPE and hits.dtype == HIT_DTYPE and len(hits) >= len(rays)
        _check(self.L.gpx_raycast_batch_async(self.h, rays.ctypes.data, len(rays), hits.ctypes.data), "gpx_raycast_batch_async")

    def raycast_device(self, d_rays: int, n: int, d_hits: int) -> None:
        """gpx_raycast_batch_device: buffers already resident in HBM; asynchronous on the world's stream."""
        _check(self.L.gpx_raycast_batch_device(self.h, C.c_void_p(d_rays), n, C.c_void_p(d_hits)),
               "gpx_raycast_batch_device")

    def timer_begin(self):
        _check(self.L.gpx_timer_begin(self.h), "gpx_timer_begin")

    def timer_end(self) -> float:
        ms = self.L.gpx_timer_end(self.h)
        if ms < 0:
            raise GpxError(f"gpx_timer_end failed: {self.L.gpx_last_error().decode()}")
        return ms

    PHASES = ("load", "forces", "static_narrow", "pair_narrow", "match", "colour", "setup", "warm_start", "velocity",
              "integrate", "position", "cache", "store")

    def phase_cycles(self, enable=True) -> dict:
        """Read (and re-arm) the per-phase SM cycle counters of the tick kernel (profiling aid)."""
        out = np.zeros(16, np.uint64)
        _check(self.L.gpx_debug_phase_cycles(self.h, 1 if enable else 0, out.ctypes.data), "gpx_debug_phase_cycles")
        return {k: int(v) for k, v in zip(self.PHASES, out)}

    # ---- player character (JPH_CharacterVirtual as used by PlayerPhysics.c)
    def character_create(self, pos, half_height=0.2, radius=0.25, max_slope_deg=50.0, world=0):
        d = CharacterDesc()
        d.half_height, d.radius, d.max_slope_deg, d.mass = half_height, radius, max_slope_deg, 10.0
        d.position[:] = pos
        _check(self.L.gpx_character_create(self.h, world, C.byref(d)), "gpx_character_create")

    def character_set_velocity(self, v, world=0):
        _check(self.L.gpx_character_set_linear_velocity(self.h, world, (C.c_float * 3)(*v)), "gpx_character_set_linear_velocity")

    def character_set_position(self, p, world=0):
        _check(self.L.gpx_character_set_position(self.h, world, (C.c_float * 3)(*p)), "gpx_character_set_position")

    def character_update(self, dt=1.0 / 60.0, settings=None):
        """settings: JPH_ExtendedUpdateSettings as (stick_to_floor_step_down, walk_stairs_step_up, min_step_forward,
        step_forward_test, cos_angle_forward_contact); None = plain update."""
        if settings is None:
            _check(self.L.gpx_character_update(self.h, dt), "gpx_character_update")
        else:
            _check(self.L.gpx_character_update_ex(self.h, dt, C.byref((C.c_float * 5)(*settings))), "gpx_character_update_ex")

    def character_get(self, world=0):
        s = CharacterState()
        _check(self.L.gpx_character_get(self.h, world, C.byref(s)), "gpx_character_get")
        return (np.array(list(s.position), np.float32), np.array(list(s.linear_velocity), np.float32), s.ground_state,
                s.ground_body)

    def character_contacts(self, world=0) -> np.ndarray:
        """Ids the character touches after the last character_update (bodies, then static meshes)."""
        out = np.zeros(64, np.uint32)
        n = C.c_uint32()
        _check(self.L.gpx_character_contacts(self.h, world, out.ctypes.data, 64, C.byref(n)), "gpx_character_contacts")
        return out[:n.value].copy()

    def enable_events(self, on=True):
        _check(self.L.gpx_events_enable(self.h, 1 if on else 0), "gpx_events_enable")

    def poll_events(self) -> np.ndarray:
        """Contact events of the last completed tick: records (world, body_a, body_b, kind)."""
        cap = self.worlds * 64
        while True:
            out = np.zeros(cap, EVENT_DTYPE)
            n = C.c_uint64(0)
            rc = self.L.gpx_poll_events(self.h, out.ctypes.data, cap, C.byref(n))
            if rc == 0:
                return out[:n.value]
            if rc != 17:
                raise GpxError(f"gpx_poll_events failed ({rc})")
            cap = int(n.value)

    def device_sync(self):
        _check(self.L.gpx_device_sync(self.h), "gpx_device_sync")

    def raycast_transform(self, pos, rot, max_distance, mask=RAYMASK_STATIC_DYNAMIC, world=0):
        t = Transform()
        t.position[:] = pos
        t.rotation[:] = rot
        hit = np.zeros(1, HIT_DTYPE)
        _check(self.L.gpx_raycast_transform(self.h, world, C.byref(t), max_distance, mask, hit.ctypes.data),
               "gpx_raycast_transform")
        return hit[0]


def pinned_array(n: int, dtype) -> np.ndarray:
    """numpy view of page-locked host memory from gpx_host_alloc (kept alive by the returned array's base)."""
    dtype = np.dtype(dtype)
    L = lib()
    nbytes = max(int(n) * dtype.itemsize, 1)
    p = L.gpx_host_alloc(nbytes)
    if not p:
        raise GpxError("gpx_host_alloc failed")

    class _Owner:
        def __init__(self, ptr):
            self.ptr = ptr
            self.buf = (C.c_uint8 * nbytes).from_address(ptr)

        def __del__(self):
            try:
                L.gpx_host_free(C.c_void_p(self.ptr))
            except Exception:
                pass

    o = _Owner(p)
    a = np.frombuffer(o.buf, dtype=dtype, count=int(n))
    _PINNED.append(o)  # freed at interpreter exit; harness buffers live for the whole run
    return a


_PINNED: list = []


def model_collision(data: bytes, container=False, tolerance=0.03) -> ModelCollision:
    """gpx_model_load_gmdl(_container): the collision section of a .gmdl.  Host-side, needs no device."""
    out = ModelCollision()
    buf = (C.c_uint8 * len(data)).from_buffer_copy(data)
    fn = lib().gpx_model_load_gmdl_container if container else lib().gpx_model_load_gmdl
    _check(fn(C.addressof(buf), len(data), tolerance, C.byref(out)), "gpx_model_load_gmdl")
    return out


def shape_from_hull(points, tolerance=0.03):
    """gpx_shape_from_hull: classify a convex hull's points as BOX / SPHERE.  Host-side, needs no device."""
    pts = np.ascontiguousarray(points, np.float32).reshape(-1, 3)
    out = HullShape()
    _check(lib().gpx_shape_from_hull(pts.ctypes.data, len(pts), tolerance, C.byref(out)), "gpx_shape_from_hull")
    return out.shape, np.array(list(out.half_extents), np.float32), np.array(list(out.center), np.float32), bool(out.exact)
