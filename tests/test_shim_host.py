"""CPU checks of the joltc-subset shim: it builds, exports everything include/joltc_gpx.h declares, engine-style C compiles
against include/joltc/ with -Werror, the host quaternion helpers give known answers, and without a device it refuses to
start instead of computing anything on the host."""
import ctypes as C
import math
import os
import re
import struct
import subprocess

import numpy as np
import pytest

import shim_build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class V3(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("z", C.c_float)]


class Q(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("z", C.c_float), ("w", C.c_float)]


@pytest.fixture(scope="module")
def shim():
    return C.CDLL(shim_build.build_shim())


def test_shim_exports_every_declared_symbol(shim):
    hdr = open(os.path.join(ROOT, "include", "joltc_gpx.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    hdr = "\n".join(l for l in hdr.splitlines() if not l.lstrip().startswith("#"))
    names = sorted(set(re.findall(r"\b((?:JPH|Vector3)_[A-Za-z0-9_]+)\s*\((?!\s*\*)", hdr)))
    assert len(names) >= 95
    missing = [n for n in names if not hasattr(shim, n)]
    assert not missing, f"declared but not exported: {missing}"
    # every JPH_* function the reference calls is declared (the census of SURVEY §8b)
    census = """JPH_BodyCreationSettings_Create2_GAME JPH_BodyCreationSettings_Create_GAME JPH_BodyCreationSettings_Destroy
    JPH_BodyCreationSettings_SetAllowedDOFs JPH_BodyCreationSettings_SetFriction JPH_BodyCreationSettings_SetIsSensor
    JPH_BodyCreationSettings_SetMassPropertiesOverride JPH_BodyCreationSettings_SetOverrideMassProperties JPH_BodyDrawFilter_Create
    JPH_BodyDrawFilter_Destroy JPH_BodyDrawFilter_SetImpl JPH_BodyFilter_Create JPH_BodyFilter_Destroy JPH_BodyInterface_CreateAndAddBody
    JPH_BodyInterface_GetPosition JPH_BodyInterface_GetPositionAndRotation JPH_BodyInterface_GetRotation JPH_BodyInterface_GetUserData
    JPH_BodyInterface_GetWorldTransform JPH_BodyInterface_RemoveAndDestroyBody JPH_BodyInterface_SetLinearAndAngularVelocity
    JPH_BodyInterface_SetLinearVelocity JPH_BodyInterface_SetPosition JPH_BodyInterface_SetRotation JPH_Body_GetObjectLayer
    JPH_Body_GetUserData JPH_BoxShape_Create JPH_BroadPhaseLayerFilter_Create JPH_BroadPhaseLayerFilter_Destroy
    JPH_BroadPhaseLayerInterface_Create JPH_CapsuleShape_Create JPH_CharacterBase_GetGroundState JPH_CharacterContactListener_Create
    JPH_CharacterContactListener_Destroy JPH_CharacterVirtualSettings_Init JPH_CharacterVirtual_Create JPH_CharacterVirtual_Destroy
    JPH_CharacterVirtual_ExtendedUpdate JPH_CharacterVirtual_GetLinearVelocity JPH_CharacterVirtual_GetPosition
    JPH_CharacterVirtual_GetUserData JPH_CharacterVirtual_SetLinearVelocity JPH_CharacterVirtual_SetListener
    JPH_CharacterVirtual_SetPosition JPH_CharacterVirtual_SetRotation JPH_CharacterVirtual_SetUserData JPH_CompoundShapeSettings_AddShape2
    JPH_ConvexHullShape_Create JPH_CylinderShape_Create JPH_DebugRenderer_Create JPH_DebugRenderer_Destroy JPH_DebugRenderer_SetImpl
    JPH_EmptyShapeSettings_Create JPH_Init JPH_JobSystemThreadPool_Create JPH_JobSystem_Destroy JPH_MeshShapeSettings_Create
    JPH_MeshShapeSettings_CreateShape JPH_NarrowPhaseQuery_CastRay2_GAME JPH_NarrowPhaseQuery_CastRay_GAME JPH_ObjectLayerFilter_Create
    JPH_ObjectLayerFilter_Destroy JPH_ObjectLayerPairFilter_Create JPH_ObjectVsBroadPhaseLayerFilter_Create JPH_PhysicsSystem_Create
    JPH_PhysicsSystem_Destroy JPH_PhysicsSystem_DrawBodies JPH_PhysicsSystem_GetBodyInterface JPH_PhysicsSystem_GetNarrowPhaseQuery
    JPH_PhysicsSystem_OptimizeBroadPhase JPH_PhysicsSystem_SetGravity JPH_PhysicsSystem_Update JPH_Quat_FromEulerAngles
    JPH_Quat_GetEulerAngles JPH_Quat_GetRotationAngle JPH_Quat_Lerp JPH_Quat_Multiply JPH_Quat_Normalized JPH_Quat_Rotate
    JPH_Quat_RotateAxisZ JPH_Quat_Rotation JPH_ShapeFilter_Create JPH_ShapeFilter_Destroy JPH_ShapeSettings_Destroy JPH_Shape_Destroy
    JPH_Shutdown JPH_StaticCompoundShapeSettings_Create JPH_StaticCompoundShape_Create""".split()
    assert len(census) == 88
    assert not [n for n in census if n not in names]


def test_forwarding_headers_cover_the_engine_includes():
    wanted = """Math/Quat.h Math/RMat44.h Math/RVec3.h Math/Transform.h Math/Vector3.h Physics/Body/Body.h
    Physics/Body/BodyCreationSettings.h Physics/Body/BodyFilter.h Physics/Body/BodyID.h Physics/Body/BodyInterface.h
    Physics/Body/MassProperties.h Physics/Collision/BroadPhase/BroadPhaseLayer.h Physics/Collision/CastResult.h
    Physics/Collision/NarrowPhaseQuery.h Physics/Collision/ObjectLayer.h Physics/Collision/Shape/Shape.h
    Physics/Collision/Shape/SubShapeID.h Physics/Collision/ShapeFilter.h constants.h enums.h joltc.h types.h""".split()
    for h in wanted:
        assert os.path.exists(os.path.join(ROOT, "include", "joltc", h)), h


def test_engine_style_c_compiles_and_refuses_to_run_without_a_device(tmp_path):
    driver = shim_build.build_driver()          # gcc -std=gnu11 -Wall -Wextra -Werror against include/joltc/
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    scene = tmp_path / "empty.bin"
    scene.write_bytes(struct.pack("<III", 0, 0, 0))
    r = subprocess.run([driver, str(scene), "1"], capture_output=True, text=True)
    assert r.returncode == 4 and "no usable CUDA device" in r.stderr     # JPH_Init said no; nothing was simulated


def test_quaternion_helpers_known_answers(shim):
    shim.JPH_Quat_GetRotationAngle.restype = C.c_float
    shim.Vector3_Length.restype = C.c_float
    q, v, out = Q(), V3(), V3()
    axis_y = V3(0, 1, 0)
    shim.JPH_Quat_Rotation(C.byref(axis_y), C.c_float(math.pi / 2), C.byref(q))       # +90 deg about y
    assert np.allclose([q.x, q.y, q.z, q.w], [0, math.sqrt(0.5), 0, math.sqrt(0.5)], atol=1e-7)
    shim.JPH_Quat_Rotate(C.byref(q), C.byref(V3(0, 0, -1)), C.byref(out))             # forward (-z) turns to -x
    assert np.allclose([out.x, out.y, out.z], [-1, 0, 0], atol=1e-6)
    shim.JPH_Quat_RotateAxisZ(C.byref(q), C.byref(out))
    assert np.allclose([out.x, out.y, out.z], [1, 0, 0], atol=1e-6)
    assert abs(shim.JPH_Quat_GetRotationAngle(C.byref(q), C.byref(axis_y)) - math.pi / 2) < 1e-6
    # euler round trip (MapLoader.c:90 builds actor rotations this way), composition order x then y then z
    e = V3(0.3, -0.7, 1.1)
    shim.JPH_Quat_FromEulerAngles(C.byref(e), C.byref(q))
    qx, qy, qz, t = Q(), Q(), Q(), Q()
    shim.JPH_Quat_Rotation(C.byref(V3(1, 0, 0)), C.c_float(0.3), C.byref(qx))
    shim.JPH_Quat_Rotation(C.byref(axis_y), C.c_float(-0.7), C.byref(qy))
    shim.JPH_Quat_Rotation(C.byref(V3(0, 0, 1)), C.c_float(1.1), C.byref(qz))
    shim.JPH_Quat_Multiply(C.byref(qz), C.byref(qy), C.byref(t))
    shim.JPH_Quat_Multiply(C.byref(t), C.byref(qx), C.byref(t))
    assert np.allclose([q.x, q.y, q.z, q.w], [t.x, t.y, t.z, t.w], atol=1e-6)
    shim.JPH_Quat_GetEulerAngles(C.byref(q), C.byref(out))
    assert np.allclose([out.x, out.y, out.z], [0.3, -0.7, 1.1], atol=1e-5)
    # lerp + normalise as the held-object code uses them (PlayerPhysics.c:374-375)
    a, b = Q(0, 0, 0, 1), Q(0, 1, 0, 0)
    shim.JPH_Quat_Lerp(C.byref(a), C.byref(b), C.c_float(0.2), C.byref(t))
    assert np.allclose([t.x, t.y, t.z, t.w], [0, 0.2, 0, 0.8], atol=1e-7)
    shim.JPH_Quat_Normalized(C.byref(t), C.byref(t))
    assert abs(t.x ** 2 + t.y ** 2 + t.z ** 2 + t.w ** 2 - 1) < 1e-6
    shim.Vector3_Normalized(C.byref(V3(3, 0, 4)), C.byref(out))
    assert np.allclose([out.x, out.y, out.z], [0.6, 0, 0.8], atol=1e-7) and abs(shim.Vector3_Length(C.byref(V3(3, 0, 4))) - 5) < 1e-6
    ay = V3.in_dll(shim, "Vector3_AxisY")
    qi = Q.in_dll(shim, "JPH_Quat_Identity")
    assert (ay.x, ay.y, ay.z) == (0, 1, 0) and (qi.x, qi.y, qi.z, qi.w) == (0, 0, 0, 1)


def test_layer_callbacks_become_masks_and_shapes_are_refcounted(shim):
    """Filter objects and shapes are host bookkeeping: creating and destroying them needs no device."""
    bp_cb = C.CFUNCTYPE(C.c_bool, C.c_uint8)
    ol_cb = C.CFUNCTYPE(C.c_bool, C.c_uint32)

    class BpImpl(C.Structure):
        _fields_ = [("ShouldCollide", bp_cb)]

    class OlImpl(C.Structure):
        _fields_ = [("ShouldCollide", ol_cb)]

    shim.JPH_BroadPhaseLayerFilter_Create.restype = C.c_void_p
    shim.JPH_ObjectLayerFilter_Create.restype = C.c_void_p
    seen = []
    static_only = ol_cb(lambda l: (seen.append(l), l == 0)[1])
    f = shim.JPH_ObjectLayerFilter_Create(C.byref(OlImpl(static_only)))
    assert sorted(seen) == [0, 1, 2, 3]                      # evaluated once per engine layer, at create time
    assert C.cast(f, C.POINTER(C.c_uint32))[0] == 0b0001
    shim.JPH_ObjectLayerFilter_Destroy.argtypes = [C.c_void_p]
    shim.JPH_ObjectLayerFilter_Destroy(f)
    both = bp_cb(lambda l: True)
    g = shim.JPH_BroadPhaseLayerFilter_Create(C.byref(BpImpl(both)))
    assert C.cast(g, C.POINTER(C.c_uint32))[0] == 0b11
    shim.JPH_BroadPhaseLayerFilter_Destroy.argtypes = [C.c_void_p]
    shim.JPH_BroadPhaseLayerFilter_Destroy(g)

    # shapes: ctor returns +1; a compound takes its own reference to each child; exactness is reported
    shim.JPH_BoxShape_Create.restype = C.c_void_p
    shim.JPH_BoxShape_Create.argtypes = [C.POINTER(V3), C.c_float]
    shim.JPH_StaticCompoundShapeSettings_Create.restype = C.c_void_p
    shim.JPH_StaticCompoundShape_Create.restype = C.c_void_p
    shim.JPH_StaticCompoundShape_Create.argtypes = [C.c_void_p]
    shim.JPH_CompoundShapeSettings_AddShape2.argtypes = [C.c_void_p, C.POINTER(V3), C.POINTER(Q), C.c_void_p, C.c_uint32]
    shim.JPH_Shape_Destroy.argtypes = [C.c_void_p]
    shim.JPH_ShapeSettings_Destroy.argtypes = [C.c_void_p]
    shim.JPH_GPX_ShapeIsExact.argtypes = [C.c_void_p]
    box = shim.JPH_BoxShape_Create(C.byref(V3(0.5, 1.0, 2.0)), 0.05)
    refs = C.cast(box, C.POINTER(C.c_int))
    assert refs[0] == 1 and shim.JPH_GPX_ShapeIsExact(box) == 1
    st = shim.JPH_StaticCompoundShapeSettings_Create()
    shim.JPH_CompoundShapeSettings_AddShape2(st, C.byref(V3(0, 0, 0)), C.byref(Q(0, 0, 0, 1)), box, 0)
    assert refs[0] == 2
    comp = shim.JPH_StaticCompoundShape_Create(st)
    assert refs[0] == 3 and shim.JPH_GPX_ShapeIsExact(comp) == 1
    shim.JPH_ShapeSettings_Destroy(st)
    assert refs[0] == 2
    shim.JPH_Shape_Destroy(comp)
    assert refs[0] == 1
    shim.JPH_Shape_Destroy(box)
    # a rotated box inside a compound is only bounded, not represented
    box = shim.JPH_BoxShape_Create(C.byref(V3(0.5, 1.0, 2.0)), 0.05)
    st = shim.JPH_StaticCompoundShapeSettings_Create()
    s = math.sqrt(0.5)
    shim.JPH_CompoundShapeSettings_AddShape2(st, C.byref(V3(0, 0, 0)), C.byref(Q(0, s * math.sin(0.4) / s, 0, math.cos(0.4))), box, 0)
    comp = shim.JPH_StaticCompoundShape_Create(st)
    assert shim.JPH_GPX_ShapeIsExact(comp) == 0
    for obj, fn in ((comp, shim.JPH_Shape_Destroy), (box, shim.JPH_Shape_Destroy), (st, shim.JPH_ShapeSettings_Destroy)):
        fn(obj)
