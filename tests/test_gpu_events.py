"""GPU parity: contact events (added / persisted / removed) through gpx_poll_events vs the oracle.

The engine consumes such events as actor callbacks (OnPlayerContactAdded/Persisted/Removed,
engine/src/physics/PlayerPhysics.c:89-152; coins, goals and triggers are sensor bodies: Coin.c:41-55, Trigger.c:33-50).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_box_falls_through_a_sensor_onto_the_floor(gpx, orc, scenes):
    g = gpx.World(worlds=1, max_bodies=8)
    o = orc.World(8)
    for pos, tris in scenes.load_static("stacked"):
        g.add_mesh(pos, tris)
        o.add_mesh(pos, tris)
    g.commit()
    g.enable_events()
    descs = [dict(position=(0.0, 0.2, -1.5)),                                                        # falling box
             dict(half_extents=(0.5, 0.15, 0.5), position=(0.0, -0.2, -1.5), layer=3, motion_type=0, is_sensor=1),  # trigger
             dict(position=(0.0, -1.29, -1.5)),                                                      # box resting on the floor
             dict(shape=2, half_extents=(0.2, 0, 0), position=(0.9, -1.0, -1.5))]                    # sphere beside it
    for d in descs:
        assert g.create(gpx.body_desc(**d)) == o.create(orc.body_desc(**d))
    seen = {1: set(), 2: set(), 3: set()}
    for tick in range(1, 121):
        assert g.step() == 0 and o.step() == 0
        eg = g.poll_events()
        eo = o.events()
        got = np.stack([eg["body_a"], eg["body_b"], eg["kind"]], axis=1) if len(eg) else np.zeros((0, 3), np.uint32)
        assert np.array_equal(got, eo), f"tick {tick}: events differ\n{got}\n{eo}"
        assert (eg["world"] == 0).all()
        for a, b, k in got:
            seen[int(k)].add((int(a), int(b)))
    # the falling box entered the sensor, left it again, landed on the resting box
    assert (0, 1) in seen[1] and (0, 1) in seen[2] and (0, 1) in seen[3]
    assert (0, 2) in seen[1] and (0, 2) in seen[2]
    # resting box and sphere touch the floor mesh of sector 0 from the first tick on
    assert any(a == 2 and b >= gpx.STATIC_BODY_BASE for a, b in seen[1])
    assert any(a == 3 and b >= gpx.STATIC_BODY_BASE for a, b in seen[1])
    # state parity is unaffected by event generation
    assert np.array_equal(g.transforms()[0, :4].view(np.uint32), o.state(4)[0].view(np.uint32))


def test_events_per_world_in_an_ensemble(gpx, orc, scenes):
    W = 6
    g = gpx.World(worlds=W, max_bodies=8)
    os_ = [orc.World(8) for _ in range(W)]
    for pos, tris in scenes.load_static("stacked"):
        g.add_mesh(pos, tris)
        for o in os_:
            o.add_mesh(pos, tris)
    g.commit()
    g.enable_events()
    vel = scenes.ensemble_velocities(W, 8)
    pos = scenes.stack_positions(8)
    g.create_all([gpx.body_desc(position=tuple(p)) for p in pos], linvel=vel)
    for wi, o in enumerate(os_):
        for k in range(8):
            o.create(orc.body_desc(position=tuple(pos[k]), linear_velocity=tuple(vel[wi, k])))
    for tick in range(1, 31):
        assert g.step() == 0
        eg = g.poll_events()
        for wi, o in enumerate(os_):
            assert o.step() == 0
            mine = eg[eg["world"] == wi]
            got = np.stack([mine["body_a"], mine["body_b"], mine["kind"]], axis=1) if len(mine) else np.zeros((0, 3), np.uint32)
            assert np.array_equal(got, o.events()), f"tick {tick} world {wi}"
    assert len(eg) == W * 8 and (eg["kind"] == 2).all()     # settled column: 8 persisted contacts per world


def test_events_must_be_enabled(gpx):
    g = gpx.World(worlds=1, max_bodies=8)
    with pytest.raises(gpx.GpxError):
        g.poll_events()


def test_wide_world_events_sensors_and_character_contacts(gpx, orc, scenes):
    """The wide-world path (> 64 bodies): touching pairs are sorted and diffed once per tick.  A lattice settling on the
    shipped map, a box falling through a sensor, the player capsule walking into the columns: event streams, bodies and
    character state identical to the oracle's wide mode."""
    n = 80
    g = gpx.World(worlds=1, max_bodies=n)
    o = orc.World(n)
    for pos, tris in scenes.load_static("stacked"):
        g.add_mesh(pos, tris)
        o.add_mesh(pos, tris)
    g.commit()
    g.enable_events()
    descs = []
    for ix in range(5):
        for iz in range(4):
            for k in range(3):
                descs.append(dict(position=(-1.6 + 0.8 * ix, -1.25 + 0.43 * k, -2.6 + 0.7 * iz)))
    descs.append(dict(position=(0.3, 0.6, 0.8)))                                                               # 60: falls ...
    descs.append(dict(half_extents=(0.5, 0.15, 0.5), position=(0.3, -0.2, 0.8), layer=3, motion_type=0, is_sensor=1))  # 61: ... through this
    descs.append(dict(half_extents=(0.25, 0.25, 0.25), position=(0.0, -1.25, 1.6), layer=3, motion_type=0, is_sensor=1))  # 62: a coin
    for d in descs:
        assert g.create(gpx.body_desc(**d)) == o.create(orc.body_desc(**d))
    g.character_create((0.0, -0.9, 2.6))
    o.character_create((0.0, -0.9, 2.6))
    seen = {1: set(), 2: set(), 3: set()}
    CH = 0x3FFFFF
    for tick in range(1, 151):
        for side in (g, o):
            p, vel, ground, _ = side.character_get()
            vy = 0.0 if ground == 0 else float(vel[1]) + (-9.81 / 60.0)
            side.character_set_velocity((0.0, vy, -1.5 if tick > 20 else 0.0))
            side.character_update()
        assert g.step() == 0 and o.step() == 0
        eg = g.poll_events()
        got = np.stack([eg["body_a"], eg["body_b"], eg["kind"]], axis=1) if len(eg) else np.zeros((0, 3), np.uint32)
        eo = o.events()
        assert np.array_equal(got, eo), f"tick {tick}: {len(got)} vs {len(eo)} events"
        for a, b, k in got:
            seen[int(k)].add((int(a), int(b)))
    assert (60, 61) in seen[1] and (60, 61) in seen[3]                       # through the sensor
    assert (62, CH) in seen[1] and (62, CH) in seen[3]                       # the character crossed the coin
    assert any(a == CH and b >= gpx.STATIC_BODY_BASE for a, b in seen[2])    # and stands on the map
    assert any(b == CH and a < 60 for a, b in seen[1])                       # and reached the columns
    assert len(seen[2]) > 60
    assert g.sync() == 0
    assert np.array_equal(g.transforms()[0, :63].view(np.uint32), o.state(63)[0].view(np.uint32))
    pg, po = g.character_get(), o.character_get()
    assert np.array_equal(pg[0].view(np.uint32), po[0].view(np.uint32)) and pg[2:] == po[2:]


def test_events_of_an_ensemble_of_wide_worlds(gpx, orc, scenes):
    """Three worlds of 80 slots each (the wide kernels, bodies of all worlds in one array): every world has its own pile, a
    sensor, a coin and a walking character; the event stream of each world, tick by tick, is its own oracle's."""
    n, W = 80, 3
    g = gpx.World(worlds=W, max_bodies=n)
    os_ = [orc.World(n) for _ in range(W)]
    for pos, tris in scenes.load_static("stacked"):
        g.add_mesh(pos, tris)
        for o in os_:
            o.add_mesh(pos, tris)
    g.commit()
    g.enable_events()
    for wi, o in enumerate(os_):
        descs = []
        for ix in range(3 + wi):
            for k in range(3):
                descs.append(dict(position=(-1.2 + 0.8 * ix, -1.25 + 0.43 * k, -2.6 + 0.3 * wi)))
        descs.append(dict(position=(0.3, 0.6, 0.8)))
        descs.append(dict(half_extents=(0.5, 0.15, 0.5), position=(0.3, -0.2, 0.8), layer=3, motion_type=0, is_sensor=1))
        descs.append(dict(half_extents=(0.25, 0.25, 0.25), position=(0.0, -1.25, 1.6 - 0.2 * wi), layer=3, motion_type=0, is_sensor=1))
        for d in descs:
            assert g.create(gpx.body_desc(**d), world=wi) == o.create(orc.body_desc(**d))
        g.character_create((0.0, -0.9, 2.6), world=wi)
        o.character_create((0.0, -0.9, 2.6))
    kinds = [set() for _ in range(W)]
    for tick in range(1, 121):
        for wi, o in enumerate(os_):
            p, vel, ground, _ = o.character_get()
            vy = 0.0 if ground == 0 else float(vel[1]) + (-9.81 / 60.0)
            v = (0.0, vy, (-1.5 - 0.3 * wi) if tick > 20 else 0.0)
            o.character_set_velocity(v)
            g.character_set_velocity(v, world=wi)
            o.character_update()
        g.character_update()
        assert g.step() == 0
        for o in os_:
            assert o.step() == 0
        eg = g.poll_events()
        for wi, o in enumerate(os_):
            e = eg[eg["world"] == wi]
            got = np.stack([e["body_a"], e["body_b"], e["kind"]], axis=1) if len(e) else np.zeros((0, 3), np.uint32)
            eo = o.events()
            assert np.array_equal(got, eo), f"tick {tick} world {wi}: {len(got)} vs {len(eo)} events\n{got[:6]}\n{eo[:6]}"
            kinds[wi] |= {int(k) for k in got[:, 2]}
            pg, vg, gg, bg = g.character_get(world=wi)
            po, vo, go, bo = o.character_get()
            assert np.array_equal(pg.view(np.uint32), po.view(np.uint32)) and (gg, bg) == (go, bo)
        # the worlds' records come world after world
        assert np.all(np.diff(eg["world"].astype(np.int64)) >= 0)
    assert all(k == {1, 2, 3} for k in kinds)
