"""GPU parity: the player character (capsule collide-and-slide, ground state, contact callbacks) through the C ABI vs
the oracle.  Call order per tick is MapFixedUpdate's: MovePlayer sets the velocity, UpdatePlayer advances the
character, then the physics update (engine/src/physics/MapPhysics.c:66-108, PlayerPhysics.c:283-294,447)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _pair(gpx, orc, scenes, static="stacked", cap=8, worlds=1):
    g = gpx.World(worlds=worlds, max_bodies=cap)
    os_ = [orc.World(cap) for _ in range(worlds)]
    for pos, tris in scenes.load_static(static):
        g.add_mesh(pos, tris)
        for o in os_:
            o.add_mesh(pos, tris)
    g.commit()
    return g, os_


def _move(side, v, gravity=True):
    p, vel, ground, _ = side.character_get()
    vy = 0.0
    if gravity and ground != 0:
        vy = float(vel[1]) + (-9.81 / 60.0)
    side.character_set_velocity((v[0], vy, v[2]))


def _same(g, o, what):
    pg, vg, gg, bg = g.character_get()
    po, vo, go, bo = o.character_get()
    assert np.array_equal(pg.view(np.uint32), po.view(np.uint32)), f"{what}: position {pg} vs {po}"
    assert np.array_equal(vg.view(np.uint32), vo.view(np.uint32)), f"{what}: velocity {vg} vs {vo}"
    assert (gg, bg) == (go, bo), f"{what}: ground {gg}/{bg:#x} vs {go}/{bo:#x}"


def test_character_walk_matches_oracle_on_stacked_map(gpx, orc, scenes):
    g, (o,) = _pair(gpx, orc, scenes)
    descs = [dict(position=(1.0, -1.3, -1.5), motion_type=0, layer=0),                                        # crate
             dict(position=(-1.0, -1.25, -1.5), half_extents=(0.25, 0.25, 0.25), motion_type=0, layer=3, is_sensor=1),  # coin
             dict(position=(0.3, -1.2, 0.5)),                                                                 # dynamic box
             dict(shape=2, half_extents=(0.3, 0, 0), position=(0.0, -1.2, -3.0), motion_type=0, layer=0)]      # pillar cap
    for d in descs:
        assert g.create(gpx.body_desc(**d)) == o.create(orc.body_desc(**d))
    g.enable_events()
    g.character_create((0.0, 0.0, -1.5))
    o.character_create((0.0, 0.0, -1.5))
    rng = np.random.default_rng(3)
    heading = np.array([1.5, 0.0, 0.0])
    seen = set()
    for tick in range(1, 361):
        if tick % 45 == 0:
            a = rng.uniform(0, 2 * np.pi)
            heading = np.array([2.5 * np.cos(a), 0.0, 2.5 * np.sin(a)])
        for side in (g, o):
            _move(side, heading if tick > 40 else (0.0, 0.0, 0.0))
            side.character_update()
        assert g.step() == 0 and o.step() == 0
        _same(g, o, f"tick {tick}")
        eg = g.poll_events()
        got = np.stack([eg["body_a"], eg["body_b"], eg["kind"]], axis=1) if len(eg) else np.zeros((0, 3), np.uint32)
        assert np.array_equal(got, o.events()), f"tick {tick}: events differ"
        for a_, b_, k_ in got:
            if gpx.lib() and (a_ == 0x3FFFFF or b_ == 0x3FFFFF):
                seen.add((int(a_), int(b_), int(k_)))
    p, v, ground, gb = g.character_get()
    assert ground == 0 and gb >= gpx.STATIC_BODY_BASE                    # standing on some sector's floor mesh
    assert any(a == 0x3FFFFF and b >= gpx.STATIC_BODY_BASE for a, b, k in seen)
    # bodies keep matching too (the dynamic box gets pushed around by the character on both sides)
    assert np.array_equal(g.transforms()[0, :4].view(np.uint32), o.state(4)[0].view(np.uint32))


def test_character_lands_slides_and_is_blocked(gpx, orc, scenes):
    g, (o,) = _pair(gpx, orc, scenes)
    d = dict(position=(1.0, -1.3, -1.5), motion_type=0, layer=0)
    assert g.create(gpx.body_desc(**d)) == o.create(orc.body_desc(**d))
    g.character_create((0.0, 0.0, -1.5))
    o.character_create((0.0, 0.0, -1.5))
    for tick in range(90):
        for side in (g, o):
            _move(side, (0.0, 0.0, 0.0))
            side.character_update()
    _same(g, o, "landing")
    p, v, ground, gb = g.character_get()
    assert abs(p[1] - (-1.5 + 0.45)) < 1e-3 and ground == 0 and gb >= gpx.STATIC_BODY_BASE
    for tick in range(60):
        for side in (g, o):
            _move(side, (1.5, 0.0, 0.0))
            side.character_update()
    _same(g, o, "crate")
    assert abs(g.character_get()[0][0] - (1.0 - 0.2 - 0.25)) < 2e-3       # stopped at the crate's face
    for tick in range(150):
        for side in (g, o):
            _move(side, (0.7, 0.0, -3.0))
            side.character_update()
    _same(g, o, "wall")
    assert abs(g.character_get()[0][2] - (-4.0 + 0.25)) < 2e-3            # held at radius distance from the z = -4 wall


def test_characters_of_an_ensemble_are_independent(gpx, orc, scenes):
    W = 5
    g, os_ = _pair(gpx, orc, scenes, worlds=W)
    for wi, o in enumerate(os_):
        g.character_create((0.2 * wi, -1.0, -1.5), world=wi)
        o.character_create((0.2 * wi, -1.0, -1.5))
    for tick in range(80):
        for wi, o in enumerate(os_):
            p, vel, ground, _ = o.character_get()
            v = (0.5 * (wi - 2), 0.0 if ground == 0 else float(vel[1]) - 9.81 / 60.0, -0.4 * wi)
            o.character_set_velocity(v)
            o.character_update()
            g.character_set_velocity(v, world=wi)
        g.character_update()
    for wi, o in enumerate(os_):
        pg, vg, gg, bg = g.character_get(world=wi)
        po, vo, go, bo = o.character_get()
        assert np.array_equal(pg.view(np.uint32), po.view(np.uint32)) and (gg, bg) == (go, bo), f"world {wi}"


def test_batched_capsule_overlap_queries_match_the_oracle(gpx, orc, scenes):
    """gpx_overlap_capsule_batch: the collide-shape query of the character controller for arbitrary upright capsules —
    4096 of them scattered through stacked.gmap and a few bodies, depth / normal / body bit-identical to the oracle."""
    g, (o,) = _pair(gpx, orc, scenes, cap=8)
    descs = [dict(position=(1.0, -1.3, -1.5), motion_type=0, layer=0),
             dict(position=(-0.5, -1.0, 0.5), half_extents=(0.3, 0.5, 0.2), rotation=(0.0, 0.38268343, 0.0, 0.92387953)),
             dict(shape=2, half_extents=(0.35, 0, 0), position=(0.0, -1.1, -2.5)),
             dict(position=(0.5, -1.25, 1.0), half_extents=(0.25, 0.25, 0.25), motion_type=0, layer=3, is_sensor=1)]   # ignored
    for d in descs:
        assert g.create(gpx.body_desc(**d)) == o.create(orc.body_desc(**d))
    rng = np.random.default_rng(5)
    n = 4096
    q = np.zeros(n, gpx.CAPSULE_DTYPE)
    q["center"] = rng.uniform((-4.5, -1.9, -4.5), (4.5, 0.5, 4.5), (n, 3)).astype(np.float32)
    q["half_height"] = rng.uniform(0.0, 0.4, n).astype(np.float32)
    q["radius"] = rng.uniform(0.05, 0.4, n).astype(np.float32)
    q[:8]["center"] = [(1.0, -1.3, -1.5), (0.0, -1.1, -2.5), (-0.5, -1.0, 0.5), (0.5, -1.25, 1.0),
                       (0.0, 50.0, 0.0), (0.0, -1.45, 0.0), (1.0, -0.85, -1.5), (0.0, -1.05, -1.5)]
    q[:8]["half_height"] = 0.0
    q[:8]["radius"] = 0.1
    out = g.overlap_capsules(q)
    hits = 0
    for i in range(n):
        d, nrm, body = o.overlap_capsule(q["center"][i], float(q["half_height"][i]), float(q["radius"][i]))
        assert out["body"][i] == body, f"query {i}: body {out['body'][i]:#x} vs {body:#x}"
        assert np.array([out["depth"][i]], np.float32).view(np.uint32)[0] == np.array([d], np.float32).view(np.uint32)[0], f"query {i}: depth"
        assert np.array_equal(out["normal"][i].view(np.uint32), nrm.view(np.uint32)), f"query {i}: normal"
        hits += body != 0xFFFFFFFF
    assert n // 10 < hits < n                                      # a real mix of touching and free capsules
    assert out["body"][0] == 0 and out["body"][1] == 2 and out["body"][2] == 1      # inside the crate, the sphere, the turned box
    assert out["body"][4] == gpx.INVALID_BODY and out["depth"][4] == 0.0             # far above everything
    assert out["body"][5] >= gpx.STATIC_BODY_BASE and abs(out["normal"][5][1] - 1.0) < 1e-6   # on the floor: pushed up
    assert (out["body"] != 3).all()                                                  # the sensor is never reported


def test_character_wakes_and_pushes_a_sleeping_box(gpx, orc, scenes):
    """Walking into a dynamic body applies CharacterVirtual's contact impulse: the box (asleep by then) wakes, is shoved
    along and tumbles; the character is held up by it.  Bodies, sleep states and the character match the oracle bit for bit."""
    g, (o,) = _pair(gpx, orc, scenes)
    d = dict(position=(1.0, -1.3, -1.5), allow_sleeping=1)
    assert g.create(gpx.body_desc(**d)) == o.create(orc.body_desc(**d)) == 0
    g.character_create((0.0, -1.0, -1.5))
    o.character_create((0.0, -1.0, -1.5))
    x_box = []
    for tick in range(1, 161):
        for side in (g, o):
            _move(side, (1.5, 0.0, 0.0) if tick > 60 else (0.0, 0.0, 0.0))
            side.character_update()
        assert g.step() == 0 and o.step() == 0
        if tick == 60:
            assert g.sleeping()[0, 0] and o.asleep(1)[0]
        if tick % 10 == 0:
            _same(g, o, f"tick {tick}")
            xo, vo = o.state(1)
            assert np.array_equal(g.transforms()[0, :1].view(np.uint32), xo.view(np.uint32)), f"tick {tick}: box differs"
            assert np.array_equal(g.velocities()[0, :1].view(np.uint32), vo.view(np.uint32))
            assert g.sleeping()[0, 0] == o.asleep(1)[0]
        x_box.append(float(g.get_transform(0)[0]) if tick % 10 == 0 else (x_box[-1] if x_box else 1.0))
    assert x_box[5] == 1.0                       # untouched while the character stands still
    assert max(x_box) > 1.5                      # shoved more than half a metre along +x


@pytest.mark.parametrize("height,start,vx", [(0.24, (0.0, 0.5, 0.0), 1.5), (0.30, (0.0, 0.5, 0.0), 1.5),
                                             (0.20, (2.0, 0.66, 0.0), -1.5), (0.12, (0.0, 0.5, 0.3), 2.5)])
def test_extended_update_stairs_and_stick_to_floor_match_the_oracle(gpx, orc, height, start, vx):
    """JPH_CharacterVirtual_ExtendedUpdate with the engine's settings (PlayerPhysics.c:439-446) on a floor with one step:
    up a 0.24 m step, stopped by a 0.30 m one, set down when walking off a 0.20 m one — bit for bit like the oracle,
    every tick; the contact lists agree too."""
    from test_oracle import ENGINE_EXTENDED_UPDATE, stair_scene
    g = gpx.World(worlds=1, max_bodies=8)
    o = orc.World(8)
    for side in (g, o):
        side.add_mesh((0, 0, 0), stair_scene(height))
        side.commit()
        side.character_create(start)
    for tick in range(1, 121):
        for side in (g, o):
            _move(side, (vx if tick > 20 else 0.0, 0.0, 0.0))
            side.character_update(settings=ENGINE_EXTENDED_UPDATE)
        _same(g, o, f"step {height} tick {tick}")
        assert list(g.character_contacts()) == list(o.character_contacts())
    p = o.character_get()[0]
    if height == 0.24:
        assert p[0] > 1.5 and abs(p[1] - (height + 0.45)) < 1e-3
    if height == 0.30:
        assert abs(p[0] - 0.75) < 1e-3


def test_fast_moves_are_swept_like_the_oracles(gpx, orc):
    """A tick's move longer than half the capsule radius is taken in pieces (swept motion): a fall at 40 m/s ends on the
    floor, a run at 40 m/s ends at the wall, and every tick is the oracle's, bit for bit."""
    from test_oracle import ENGINE_EXTENDED_UPDATE, stair_scene
    g = gpx.World(worlds=1, max_bodies=8)
    o = orc.World(8)
    for side in (g, o):
        side.add_mesh((0, 0, 0), stair_scene(3.0))
        side.commit()
        side.character_create((-2.0, 10.3, 0.0))
    for tick in range(1, 31):
        for side in (g, o):
            side.character_set_velocity([0.0, -40.0, 0.0])
            side.character_update(settings=ENGINE_EXTENDED_UPDATE)
        _same(g, o, f"fall tick {tick}")
    assert abs(o.character_get()[0][1] - 0.45) < 1e-3
    for tick in range(1, 41):
        for side in (g, o):
            side.character_set_velocity([40.0, 0.0, 1.0])
            side.character_update(settings=ENGINE_EXTENDED_UPDATE)
        _same(g, o, f"run tick {tick}")
    assert abs(o.character_get()[0][0] - 0.75) < 1e-3
