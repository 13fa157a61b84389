"""GPU parity: sleeping (the sleep test of JPH_PhysicsSystem_Update) through the C ABI vs the oracle.

Bodies created with allow_sleeping go to sleep island by island once their three test points have stayed within 15 mm for
0.5 s; a sleeper is static for the tick until an active body touches it or the host wakes it; a world in which nothing is
active skips its tick.  Every state — transforms, velocities, who sleeps — must match the oracle bit for bit.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _pair(gpx, orc, scenes, cap, worlds=1):
    g = gpx.World(worlds=worlds, max_bodies=cap)
    os_ = [orc.World(cap) for _ in range(worlds)]
    for pos, tris in scenes.load_static("stacked"):
        g.add_mesh(pos, tris)
        for o in os_:
            o.add_mesh(pos, tris)
    g.commit()
    return g, os_


def _same(g, os_, n, what):
    xg, vg, sg = g.transforms(), g.velocities(), g.sleeping()
    for wi, o in enumerate(os_):
        xo, vo = o.state(n)
        assert np.array_equal(xg[wi, :n].view(np.uint32), xo.view(np.uint32)), f"{what}: world {wi} transforms differ"
        assert np.array_equal(vg[wi, :n].view(np.uint32), vo.view(np.uint32)), f"{what}: world {wi} velocities differ"
        assert np.array_equal(sg[wi, :n], o.asleep(n)), f"{what}: world {wi} sleep states differ"


def test_stack_falls_asleep_is_woken_by_a_falling_box_and_sleeps_again(gpx, orc, scenes):
    g, (o,) = _pair(gpx, orc, scenes, 16)
    for p in scenes.stack_positions(8):
        d = dict(position=tuple(p), allow_sleeping=1)
        assert g.create(gpx.body_desc(**d)) == o.create(orc.body_desc(**d))
    asleep_at = None
    for tick in range(1, 121):
        assert g.step() == 0 and o.step() == 0
        if tick % 10 == 0 or g.sleeping()[0, :8].all() and asleep_at is None:
            _same(g, [o], 8, f"settling tick {tick}")
            if asleep_at is None and g.sleeping()[0, :8].all():
                asleep_at = tick
    assert asleep_at is not None and asleep_at >= 30             # at least the 0.5 s of the test
    x0 = g.transforms()[0, :8].copy()
    for _ in range(60):
        assert g.step() == 0 and o.step() == 0
    assert np.array_equal(g.transforms()[0, :8], x0) and np.abs(g.velocities()[0, :8]).max() == 0.0
    # a ninth box dropped on the column: the top box wakes on contact, the rest in a cascade, then all sleep again
    d = dict(position=(0.0, 3.0, -1.5), allow_sleeping=1)
    assert g.create(gpx.body_desc(**d)) == o.create(orc.body_desc(**d)) == 8
    counts = []
    for tick in range(1, 241):
        assert g.step() == 0 and o.step() == 0
        counts.append(int(g.sleeping()[0, :9].sum()))
        if tick % 8 == 0:
            _same(g, [o], 9, f"after the drop, tick {tick}")
    assert counts[0] == 8 and min(counts) == 0 and counts[-1] == 9
    first_wake = next(i for i, c in enumerate(counts) if c < 8)
    assert counts[first_wake] == 7                                # exactly the touched box first


def test_host_wake_setters_kinematic_contact_and_bodies_that_may_not_sleep(gpx, orc, scenes):
    g, (o,) = _pair(gpx, orc, scenes, 16)
    descs = [dict(position=(0.0, -1.25, -1.5), allow_sleeping=1),                          # 0 sleeps
             dict(position=(1.2, -1.25, -1.5), allow_sleeping=0),                          # 1 never sleeps
             dict(position=(-1.2, -1.25, -0.5), allow_sleeping=1),                         # 2 sleeps, woken by velocity
             dict(position=(-1.2, -1.25, -2.5), allow_sleeping=1),                         # 3 sleeps, hit by the platform
             dict(half_extents=(0.3, 0.2, 0.3), position=(-2.6, -1.25, -2.5), motion_type=1, layer=0)]  # 4 kinematic
    for d in descs:
        assert g.create(gpx.body_desc(**d)) == o.create(orc.body_desc(**d))
    for tick in range(1, 61):
        assert g.step() == 0 and o.step() == 0
    _same(g, [o], 5, "settled")
    s = g.sleeping()[0]
    assert s[0] and not s[1] and s[2] and s[3] and not s[4]
    g.set_velocity(2, (0.0, 2.0, 0.0))                             # a non-zero velocity activates
    o.set_velocity(2, (0.0, 2.0, 0.0))
    g.set_velocity(4, (1.0, 0.0, 0.0))                             # the platform starts moving towards body 3
    o.set_velocity(4, (1.0, 0.0, 0.0))
    g.wake(0)
    o.wake(0)
    assert g.step() == 0 and o.step() == 0
    _same(g, [o], 5, "after the wake calls")
    s = g.sleeping()[0]
    assert not s[0] and not s[2] and s[3]
    hit = None
    for tick in range(1, 121):
        assert g.step() == 0 and o.step() == 0
        if tick % 6 == 0:
            _same(g, [o], 5, f"platform tick {tick}")
        if hit is None and not g.sleeping()[0, 3]:
            hit = tick
    assert hit is not None and 40 < hit < 80                        # 1 m/s over the 0.9 m gap
    assert g.transforms()[0, 3, 0] > -1.15                          # and the platform pushed it along


def test_ensemble_worlds_sleep_independently_and_idle_worlds_cost_nothing(gpx, orc, scenes):
    """64 worlds: the even ones start at rest and fall asleep after the 0.5 s of the test, the odd ones get the C5 initial
    velocities and keep swaying.  Each world decides for itself, and the sleeping half drops out of the tick."""
    worlds = 64
    g, os_ = _pair(gpx, orc, scenes, 8, worlds)
    vel = scenes.ensemble_velocities(worlds, 8).copy()
    vel[0::2] = 0.0
    pos = scenes.stack_positions(8)
    g.create_all([gpx.body_desc(position=tuple(p), allow_sleeping=1) for p in pos], linvel=vel)
    for wi, o in enumerate(os_):
        for k, p in enumerate(pos):
            o.create(orc.body_desc(position=tuple(p), linear_velocity=tuple(vel[wi, k]), allow_sleeping=1))
    for tick in range(1, 181):
        assert g.step() == 0
        for o in os_:
            assert o.step() == 0
        if tick in (1, 30, 40, 60, 120, 180):
            _same(g, os_, 8, f"ensemble tick {tick}")
    world_asleep = g.sleeping().all(axis=1)
    assert world_asleep[0::2].all()                                  # every world that started at rest sleeps
    assert not world_asleep[1::2].all()                              # the kicked columns are mostly still moving
    x0 = g.transforms()[0::2].copy()
    for _ in range(30):
        assert g.step() == 0
    assert np.array_equal(g.transforms()[0::2], x0)                  # a sleeping world is not touched at all


def test_a_sleeping_ensemble_skips_its_ticks(gpx, scenes):
    """4096 worlds of the C5 column started at rest: once everything sleeps a tick is a load and a vote per world."""
    worlds = 4096
    g = gpx.World(worlds=worlds, max_bodies=8)
    for pos, tris in scenes.load_static("stacked"):
        g.add_mesh(pos, tris)
    g.commit()
    g.create_all([gpx.body_desc(position=tuple(p), allow_sleeping=1) for p in scenes.stack_positions(8)])

    def timed_tick():
        g.sync()
        g.timer_begin()
        assert g.step() == 0
        return g.timer_end()

    for _ in range(10):
        assert g.step() == 0
    awake_ms = min(timed_tick() for _ in range(5))
    for _ in range(60):
        assert g.step() == 0
    assert g.sleeping().all()
    idle_ms = min(timed_tick() for _ in range(5))
    assert idle_ms < 0.25 * awake_ms, f"idle tick {idle_ms:.3f} ms vs awake {awake_ms:.3f} ms"
    assert g.stats()["error"].max() == 0


def test_wide_world_islands_sleep_and_wake(gpx, orc, scenes):
    """The wide-world kernels (> 64 bodies): 36 separate 3-box columns fall asleep island by island; a sphere rolled into
    one column wakes that island only; a kinematic slab swept through another wakes it too."""
    pos = scenes.lattice_positions(6, 3, 6)
    n0 = len(pos)
    n = n0 + 2
    g = gpx.World(worlds=1, max_bodies=n)
    o = orc.World(n)
    for p, t in scenes.box_map():
        g.add_mesh(p, t)
        o.add_mesh(p, t)
    g.commit()
    for p in pos:
        d = dict(position=tuple(p), allow_sleeping=1)
        assert g.create(gpx.body_desc(**d)) == o.create(orc.body_desc(**d))

    def same(what):
        assert g.sync() == 0
        xo, vo = o.state(n)
        assert np.array_equal(g.transforms()[0, :n].view(np.uint32), xo.view(np.uint32)), f"{what}: transforms differ"
        assert np.array_equal(g.velocities()[0, :n].view(np.uint32), vo.view(np.uint32)), f"{what}: velocities differ"
        assert np.array_equal(g.sleeping()[0, :n], o.asleep(n)), f"{what}: sleep states differ"

    for tick in range(1, 151):
        assert g.step() == 0 and o.step() == 0
        if tick in (1, 30, 60, 90, 120, 150):
            same(f"settling tick {tick}")
    s = g.sleeping()[0, :n0]
    assert s.all(), f"{s.sum()} of {n0} asleep"
    x = g.transforms()[0, :n0]
    target = int(np.argmin(np.abs(x[:, 0] - x[:, 0].max()) + np.abs(x[:, 1] - x[:, 1].min())))   # a bottom box at the +x edge
    ball = dict(shape=2, half_extents=(0.15, 0, 0), position=(float(x[target, 0]) + 1.5, float(x[target, 1]), float(x[target, 2])),
                linear_velocity=(-3.0, 0.0, 0.0), mass=5.0, allow_sleeping=1)
    assert g.create(gpx.body_desc(**ball)) == o.create(orc.body_desc(**ball)) == n0
    slab = dict(half_extents=(0.1, 0.3, 0.3), position=(float(x[:, 0].min()) - 1.0, float(x[:, 1].min()) + 0.5, float(x[0, 2])),
                motion_type=1, layer=0, linear_velocity=(1.5, 0.0, 0.0))
    assert g.create(gpx.body_desc(**slab)) == o.create(orc.body_desc(**slab)) == n0 + 1
    woke = np.zeros(n0, bool)
    for tick in range(1, 91):
        assert g.step() == 0 and o.step() == 0
        if tick % 10 == 0:
            same(f"after the hits, tick {tick}")
        woke |= ~g.sleeping()[0, :n0]
    assert woke[target]                                             # the ball's column woke
    assert 3 <= woke.sum() < n0                                     # ... and not the whole world
