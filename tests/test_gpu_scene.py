"""GPU parity on BASELINE config 1: the physics content of mapSources/test.json (test.gmap), 600 fixed ticks.

Every tick runs what MapFixedUpdate does for this map without a player (engine/src/physics/MapPhysics.c:58-119): the
four lasers cast their rays (game/src/actor/prop/Laser.c:127-158), then the physics update.  The CUDA path through the
C ABI must match the oracle bit for bit; the north-star tolerances (1e-4 m, 1e-4 rad, ray ids exact, 1e-5 relative
distances) are asserted alongside.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _build(gpx, orc, scenes, cap=16):
    sc = scenes.test_map_scene()
    g = gpx.World(worlds=1, max_bodies=cap)
    o = orc.World(cap)
    for pos, rot, tris, fr in sc["meshes"]:
        g.add_mesh(pos, tris, friction=fr, rot=rot)
        o.add_mesh(pos, tris, friction=fr, rot=rot)
    g.commit()
    for d in sc["bodies"]:
        assert g.create(gpx.body_desc(**d)) == o.create(orc.body_desc(**{k: v for k, v in d.items()}))
    return sc, g, o


def test_test_map_600_ticks_with_lasers(gpx, orc, scenes):
    sc, g, o = _build(gpx, orc, scenes)
    n = len(sc["bodies"])
    nt, nn, nb = g.static_info()
    assert nt == 644 + 4 * 310 and nb == 11 + 4            # map meshes + the four laser-emitter models
    rays = scenes.laser_rays(sc["lasers"])
    phys, actor = sc["names"].index("prop_physbox"), sc["names"].index("test_actor")
    first_hits = None
    for tick in range(1, 601):
        hg = g.raycast(rays)
        if tick % 50 == 1:
            ho = o.raycast(rays)
            assert np.array_equal(hg["body"], ho["body"]) and np.array_equal(hg["face"], ho["face"])
            assert np.array_equal(hg["fraction"].view(np.uint32), ho["fraction"].view(np.uint32))
            if first_hits is None:
                first_hits = hg.copy()
        assert g.step() == 0 and o.step() == 0
        if tick in (1, 60, 300, 600):
            assert g.sync() == 0
            xg = g.transforms()[0, :n]
            xo, vo = o.state(n)
            assert np.abs(xg[:, :3] - xo[:, :3]).max() <= 1e-4
            assert np.array_equal(xg.view(np.uint32), xo.view(np.uint32)), f"tick {tick}: not bit-identical"
            assert np.array_equal(g.velocities()[0, :n].view(np.uint32), vo.view(np.uint32))
    # all four lasers end on something inside their 50 m reach
    assert (first_hits["body"] != gpx.INVALID_BODY).all() and (first_hits["fraction"] < 1.0).all()
    # the 'triple' laser only sees static geometry (Laser.c:64-72)
    assert first_hits["body"][3] >= gpx.STATIC_BODY_BASE
    x = g.transforms()[0]
    v = g.velocities()[0]
    # both dynamic bodies came to rest on the sector floor; sensors, static props and lasers never moved
    assert np.abs(v[[phys, actor], :3]).max() < 0.05
    assert x[phys, 1] < 0.0 and x[actor, 1] < 0.7
    for i, d in enumerate(sc["bodies"]):
        if d.get("motion_type", 2) != 2:
            assert np.allclose(x[i, :3], d["position"], atol=1e-6)
    # DOF lock of the test actor: TX|TY|TZ|RY only (TestActor.c:42-46)
    assert abs(x[actor, 3]) < 1e-6 and abs(x[actor, 5]) < 1e-6
    # the floor-height laser is blocked by the physbox standing in its path once the box has settled (CAN_BLOCK_LASERS)
    h = g.raycast(rays)
    assert h["body"][1] == phys or h["body"][1] >= gpx.STATIC_BODY_BASE


def test_user_data_round_trip_for_ray_targets(gpx, orc, scenes):
    """GetTargetedActor: body id of the hit -> Actor* through GetUserData (PlayerPhysics.c:297-315)."""
    sc, g, o = _build(gpx, orc, scenes)
    phys = sc["names"].index("prop_physbox")
    for _ in range(120):
        g.step()
    g.sync()
    p = g.get_transform(phys)[:3]
    h = g.raycast_transform((float(p[0]), float(p[1]), float(p[2]) + 3.0), (0, 0, 0, 1), 10.0)
    assert h["body"] == phys and abs(h["fraction"] * 10.0 - (3.0 - 0.2)) < 1e-3
    assert g.user_data(int(h["body"])) == sc["bodies"][phys]["user_data"]


def test_static_mesh_removed_at_run_time_and_ray_flags_changed(gpx, orc, scenes):
    """RemoveAndDestroyBody on a static mesh body (Map.c:113, a static-model actor's Actor.c:68) and a laser body filter whose
    answer changes (Laser.c:74-85): after the removal the device world equals an oracle world built without that mesh —
    bodies that stood on it fall, rays pass where it was, later meshes' canonical face ids move down."""
    meshes = scenes.load_static("stacked")
    victim = 0                                            # sector 0's floor/ceiling/walls: the stack stands on it
    g = gpx.World(worlds=1, max_bodies=8)
    ids = [g.add_mesh(pos, tris) for pos, tris in meshes]
    g.commit()
    o_full, o_cut = orc.World(8), orc.World(8)
    for k, (pos, tris) in enumerate(meshes):
        o_full.add_mesh(pos, tris)
        if k != victim:
            o_cut.add_mesh(pos, tris)
    for p in scenes.stack_positions(4):
        d = gpx.body_desc(position=tuple(p))
        assert g.create(d) == o_full.create(d)
    for _ in range(40):
        assert g.step() == 0 and o_full.step() == 0
    x40, v40 = o_full.state(4)
    assert np.array_equal(g.transforms()[0, :4].view(np.uint32), x40.view(np.uint32))
    rays = scenes.shapes_rays(2048, np.array([p for p, _ in meshes]))
    h_before = g.raycast(rays)
    assert np.array_equal(h_before["face"], o_full.raycast(rays)["face"])

    g.remove_mesh(ids[victim])
    # the oracle world without the mesh picks up from the same state
    for k in range(4):
        d = orc.body_desc(position=tuple(x40[k, :3]), rotation=tuple(x40[k, 3:]), linear_velocity=tuple(v40[k, :3]),
                          angular_velocity=tuple(v40[k, 3:]))
        assert o_cut.create(d) == k
    nt, _, nb = g.static_info()
    for tick in range(1, 61):
        assert g.step() == 0 and o_cut.step() == 0
        if tick == 1:
            nt2, _, nb2 = g.static_info()
            assert nt2 == nt - len(meshes[victim][1]) and nb2 == nb     # triangles gone, body ids keep their places
    xg, xo = g.transforms()[0, :4], o_cut.state(4)[0]
    assert xg[:, 1].max() < x40[:, 1].min() - 1.0                       # the floor is gone: everything fell
    # (the contact cache of the cut world starts cold, the device's does not: compare to the physical bar, not bitwise)
    assert np.abs(xg[:, :3] - xo[:, :3]).max() < 5e-2
    h_after, h_cut = g.raycast(rays), o_cut.raycast(rays)
    assert np.array_equal(h_after["face"], h_cut["face"])
    assert np.array_equal(h_after["fraction"].view(np.uint32), h_cut["fraction"].view(np.uint32))
    # body ids of the surviving meshes are unchanged on the device (the cut oracle numbers them one lower past the victim)
    stat = (h_after["body"] >= gpx.STATIC_BODY_BASE) & (h_after["body"] != gpx.INVALID_BODY)
    expect = np.where(h_cut["body"][stat] - gpx.STATIC_BODY_BASE >= victim, h_cut["body"][stat] + 1, h_cut["body"][stat])
    assert np.array_equal(h_after["body"][stat], expect)
    assert (h_before["body"] == ids[victim]).any() and not (h_after["body"] == ids[victim]).any()

    # ray flags: a body stops blocking lasers, then blocks again
    b = g.create(gpx.body_desc(position=(2.0, -1.0, 2.0), motion_type=0, layer=0, half_extents=(0.5, 0.5, 0.5)))
    ob = o_cut.create(orc.body_desc(position=(2.0, -1.0, 2.0), motion_type=0, layer=0, half_extents=(0.5, 0.5, 0.5)))
    assert b == ob
    r = np.zeros(1, gpx.RAY_DTYPE)
    r["origin"][0], r["dir"][0], r["tmax"], r["mask"] = (2.0, -1.0, 4.0), (0, 0, -1), 50.0, 0b11 | 0x100
    assert g.raycast(r)["body"][0] == b == o_cut.raycast(r)["body"][0]
    g.set_ray_flags(b, 0)
    o_cut.set_ray_flags(ob, 0)
    hg, ho = g.raycast(r)[0], o_cut.raycast(r)[0]
    assert hg["body"] != b and hg["face"] == ho["face"] and hg["fraction"] == ho["fraction"]
    g.set_ray_flags(b, 1)
    assert g.raycast(r)["body"][0] == b
