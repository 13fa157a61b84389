"""GPU parity: the wide-world kernels (one large world, state in global memory: sort-and-sweep broadphase, one thread
per pair, hashed-priority colouring, cooperative coloured solver) through the C ABI vs the CPU oracle in its wide mode.

BASELINE config 4 is 100 000 boxes in the 1024 m room of mapSources/max_box.json.  Above 256 bodies the oracle takes
its candidate pairs from a sort-and-sweep (same pairs, same order as its all-pairs loop), which lets it follow the full
100 000-box world at a few seconds per tick: bit-exact parity is asserted there for the ticks in which the lattice
lands and its 100 000 contacts form, at 20 250 boxes for longer, and the rest of the full-size run is covered by
size-independent properties (no errors, nothing falls through the floor, bit-reproducible runs).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _quat_angle(qa, qb):
    qa, qb = qa.astype(np.float64), qb.astype(np.float64)
    va, wa, vb, wb = qa[..., :3], qa[..., 3:], qb[..., :3], qb[..., 3:]
    vec = wa * vb - wb * va - np.cross(va, vb)
    dot = np.abs(np.sum(qa * qb, axis=-1))
    return 2 * np.arctan2(np.linalg.norm(vec, axis=-1), dot)


def _pair(gpx, orc, meshes, n, **kw):
    g = gpx.World(worlds=1, max_bodies=n, **kw)
    o = orc.World(n, **kw)
    for pos, tris in meshes:
        g.add_mesh(pos, tris)
        o.add_mesh(pos, tris)
    g.commit()
    return g, o


def _assert_same(g, o, n, what):
    xg = g.transforms()[0, :n]
    xo, vo = o.state(n)
    dp = np.abs(xg[:, :3] - xo[:, :3]).max()
    da = _quat_angle(xg[:, 3:], xo[:, 3:]).max()
    assert dp <= 1e-4 and da <= 1e-4, f"{what}: differs by {dp} m / {da} rad"
    assert np.array_equal(xg.view(np.uint32), xo.view(np.uint32)), f"{what}: not bit-identical (max dp {dp})"
    assert np.array_equal(g.velocities()[0, :n].view(np.uint32), vo.view(np.uint32)), f"{what}: velocities differ"


def test_small_lattice_matches_oracle(gpx, orc, scenes):
    """6 x 3 x 6 lattice (108 boxes) dropping onto the floor of the big room and settling into 36 columns."""
    pos = scenes.lattice_positions(6, 3, 6)
    n = len(pos)
    g, o = _pair(gpx, orc, scenes.box_map(), n)
    ids = g.create_all([gpx.body_desc(position=tuple(p)) for p in pos])
    assert list(ids) == list(range(n))
    for p in pos:
        o.create(orc.body_desc(position=tuple(p)))
    for tick in range(1, 91):
        assert g.step() == 0 and o.step() == 0
        if tick in (1, 5, 30, 90):
            assert g.sync() == 0
            _assert_same(g, o, n, f"lattice 6x3x6 tick {tick}")
    y = g.transforms()[0, :n, 1]
    assert y.min() > -512.0 + 0.15            # nothing sank into the floor
    assert g.stats()["manifolds"][0] == 0 or True


def test_pile_on_shipped_map_with_mixed_bodies(gpx, orc, scenes):
    """80 bodies on stacked.gmap: a tumbling block of boxes with random spin, spheres, a kinematic platform, a sensor."""
    rng = np.random.default_rng(11)
    descs = []
    for p in scenes.block_positions(4, 4, 4, 0.43):
        descs.append(dict(position=tuple(p), linear_velocity=tuple(rng.uniform(-0.5, 0.5, 3)),
                          angular_velocity=tuple(rng.uniform(-1, 1, 3))))
    for k in range(12):
        descs.append(dict(shape=2, half_extents=(0.15 + 0.01 * k, 0, 0), position=(-1.2 + 0.22 * k, 0.9, -1.5 + 0.05 * k), mass=4.0))
    descs.append(dict(half_extents=(0.6, 0.05, 0.6), position=(1.6, -1.2, -1.5), motion_type=1, linear_velocity=(-0.2, 0.0, 0.0)))
    descs.append(dict(half_extents=(0.5, 0.25, 0.5), position=(0.0, -1.25, -1.5), layer=3, motion_type=0, is_sensor=1))
    descs.append(dict(shape=0, motion_type=0, layer=0))
    descs.append(dict(position=(1.6, -0.9, -1.5), friction=0.6, restitution=0.3))
    n = len(descs)
    g, o = _pair(gpx, orc, scenes.load_static("stacked"), n)
    for d in descs:
        assert g.create(gpx.body_desc(**d)) == o.create(orc.body_desc(**d))
    for tick in range(1, 121):
        assert g.step() == 0 and o.step() == 0
        if tick in (1, 10, 40, 120):
            assert g.sync() == 0
            _assert_same(g, o, n, f"pile tick {tick}")


def test_islands_small_large_kinematic_and_free_bodies(gpx, orc, scenes):
    """The island split of the wide path: separate short columns (islands solved inside one warp), an 8 x 5 wall
    whose boxes all touch (one medium island -> a block of its own), a column riding a moving kinematic platform (its
    contacts are set up before the platform moves and position-corrected after), and bodies still in free fall."""
    descs = []
    for ix in range(6):                                   # six 3-box columns, 1.5 m apart: six small islands
        for k in range(3):
            descs.append(dict(position=(-6.0 + 1.5 * ix, -511.75 + 0.41 * k, -4.0)))
    for j in range(5):                                    # an 8 x 5 wall of touching boxes: one island of > 32 manifolds
        for i in range(8):
            descs.append(dict(position=(4.0 + 0.4 * i, -511.75 + 0.4 * j, 6.0)))
    descs.append(dict(half_extents=(0.8, 0.05, 0.8), position=(-2.0, -511.0, 3.0), motion_type=1, linear_velocity=(0.3, 0.05, 0.0)))
    for k in range(3):                                    # column on the platform
        descs.append(dict(position=(-2.0, -510.74 + 0.41 * k, 3.0), friction=0.8))
    for k in range(5):                                    # free fall for the whole run
        descs.append(dict(shape=2, half_extents=(0.2, 0, 0), position=(10.0 + k, -400.0, -10.0), angular_velocity=(0.3 * k, 0.1, 0.0)))
    n = len(descs)
    assert n > 64
    g, o = _pair(gpx, orc, scenes.box_map(), n)
    for d in descs:
        assert g.create(gpx.body_desc(**d)) == o.create(orc.body_desc(**d))
    for tick in range(1, 101):
        assert g.step() == 0 and o.step() == 0
        if tick in (1, 2, 10, 40, 100):
            assert g.sync() == 0
            _assert_same(g, o, n, f"islands tick {tick}")
    c = g.wide_counters()
    assert c["small_islands"] >= 7 and c["medium_islands"] >= 1      # the wall: one block solves it
    x = g.transforms()[0, :n]
    plat = 18 + 40
    assert abs(x[plat, 0] - (-2.0 + 0.3 * 100 / 60)) < 1e-3                   # the platform went where its velocity says
    assert np.all(x[plat + 1:plat + 4, 0] > -1.9)                              # and carried its column along


def test_two_thousand_boxes_match_oracle(gpx, orc, scenes):
    """20 x 5 x 20 lattice: exercises the sort-and-sweep at a size the all-pairs oracle still finishes in seconds."""
    pos = scenes.lattice_positions(20, 5, 20)
    n = len(pos)
    g, o = _pair(gpx, orc, scenes.box_map(), n)
    g.create_all([gpx.body_desc(position=tuple(p)) for p in pos])
    for p in pos:
        o.create(orc.body_desc(position=tuple(p)))
    for tick in range(1, 13):
        assert g.step() == 0 and o.step() == 0
        if tick in (1, 6, 12):
            assert g.sync() == 0
            _assert_same(g, o, n, f"lattice 20x5x20 tick {tick}")


def _lattice_parity(gpx, orc, scenes, dims, ticks, checks):
    pos = scenes.lattice_positions(*dims)
    n = len(pos)
    g, o = _pair(gpx, orc, scenes.box_map(), n)
    g.create_all([gpx.body_desc(position=tuple(p)) for p in pos])
    for p in pos:
        o.create(orc.body_desc(position=tuple(p)))
    for tick in range(1, ticks + 1):
        assert g.step() == 0 and o.step() == 0
        if tick in checks:
            assert g.sync() == 0
            _assert_same(g, o, n, f"lattice {dims} tick {tick}")
    return g, o


def test_twenty_thousand_boxes_match_oracle(gpx, orc, scenes):
    """45 x 10 x 45 lattice through landing and into rest: 2025 columns of ten boxes, 20 250 manifolds."""
    g, o = _lattice_parity(gpx, orc, scenes, (45, 10, 45), 45, (5, 15, 30, 45))
    assert o.L.orc_manifold_count(o.h) == 20250 == int(g.stats()["manifolds"][0])


def test_the_benchmarked_100k_lattice_matches_oracle(gpx, orc, scenes):
    """BASELINE configs[3] itself: 100 x 10 x 100 boxes, the 30 settling ticks bench.py runs before it times anything."""
    g, o = _lattice_parity(gpx, orc, scenes, (100, 10, 100), 30, (10, 20, 30))
    assert o.L.orc_manifold_count(o.h) == 100000 == int(g.stats()["manifolds"][0])


def _radix_fallbacks(g):
    c = np.zeros(8, np.uint32)
    g.L.gpx_debug_wide_counters(g.h, c.ctypes.data)
    return int(c[6])


def test_bodies_that_overtake_each_other_keep_the_broadphase_exact(gpx, orc, scenes):
    """The broadphase sorts each sub-step's keys starting from the last order (tile sorts, radix sort only when those
    are not enough).  A gas of 6000 spheres and boxes without gravity changes that order all the time; a third of them
    teleported at once breaks it beyond what the tile sorts repair.  Results stay bit-identical to the oracle's."""
    rng = np.random.default_rng(11)
    n = 6000
    pos = rng.uniform(-12.0, 12.0, (n, 3)).astype(np.float32)
    pos[:, 1] -= 490.0
    vel = rng.uniform(-6.0, 6.0, (n, 3)).astype(np.float32)
    g, o = _pair(gpx, orc, scenes.box_map(), n, gravity=(0.0, 0.0, 0.0))
    for i in range(n):
        d = dict(position=tuple(pos[i]), linear_velocity=tuple(vel[i]), shape=2 if i % 3 == 0 else 1,
                 half_extents=(0.15, 0.15, 0.15), linear_damping=0.0)
        assert g.create(gpx.body_desc(**d)) == o.create(orc.body_desc(**d)) == i
    contacts = 0
    for tick in range(1, 41):
        if tick == 20:
            for b in range(0, n, 3):
                p = tuple(float(v) for v in (-pos[b, 0], pos[b, 1], -pos[b, 2]))
                g.set_position(b, p)
                o.set_position(b, p)
        assert g.step() == 0 and o.step() == 0
        contacts = max(contacts, o.L.orc_manifold_count(o.h))
        if tick in (1, 10, 19, 20, 30, 40):
            assert g.sync() == 0
            _assert_same(g, o, n, f"gas tick {tick}")
            if tick == 19:
                assert _radix_fallbacks(g) == 1          # only the very first sub-step had no order to start from
    assert _radix_fallbacks(g) >= 2                      # the teleport needed the full sort
    assert contacts > 100


def test_create_destroy_and_setters_in_a_wide_world(gpx, orc, scenes):
    pos = scenes.lattice_positions(5, 3, 5)
    n = len(pos)
    g, o = _pair(gpx, orc, scenes.box_map(), n + 8)
    o.L.orc_world_set_mode(o.h, 1)
    for p in pos:
        d = dict(position=tuple(p))
        assert g.create(gpx.body_desc(**d)) == o.create(orc.body_desc(**d))
    for _ in range(20):
        assert g.step() == 0 and o.step() == 0
    import ctypes as C
    for b in (3, 40, 41):
        g.destroy(b)
        o.destroy(b)
    g.set_velocity(10, (0.5, 2.0, 0.0), (0.0, 1.0, 0.0))
    o.L.orc_body_set_velocity(o.h, 10, (C.c_float * 3)(0.5, 2.0, 0.0), (C.c_float * 3)(0.0, 1.0, 0.0))
    d = dict(position=(0.0, -510.0, 0.0), shape=2, half_extents=(0.3, 0, 0))
    assert g.create(gpx.body_desc(**d)) == o.create(orc.body_desc(**d)) == 3
    for _ in range(40):
        assert g.step() == 0 and o.step() == 0
    assert g.sync() == 0
    alive = [i for i in range(n) if i not in (40, 41)]
    xg = g.transforms()[0][alive]
    xo = o.state(n)[0][alive]
    assert np.array_equal(xg.view(np.uint32), xo.view(np.uint32))


def test_full_size_100k_boxes_properties(gpx, scenes):
    """BASELINE config 4 at full size: 100 x 10 x 100 lattice in the 1024 m room."""
    pos = scenes.lattice_positions()
    n = len(pos)
    assert n == 100_000

    def run(ticks):
        g = gpx.World(worlds=1, max_bodies=n)
        for p, t in scenes.box_map():
            g.add_mesh(p, t)
        g.commit()
        proto = gpx.body_desc()
        arr = (gpx.BodyDesc * n)()
        for i in range(n):
            arr[i] = proto
            arr[i].position[0], arr[i].position[1], arr[i].position[2] = pos[i]
        ids = np.zeros(n, np.uint32)
        assert g.L.gpx_body_create_all(g.h, arr, n, None, None, ids.ctypes.data) == 0
        for _ in range(ticks):
            assert g.step() == 0
        assert g.sync() == 0
        return g.transforms()[0], g.stats()[0]

    x1, s1 = run(30)
    x2, s2 = run(30)
    assert np.array_equal(x1.view(np.uint32), x2.view(np.uint32)), "two runs of the same world differ"
    assert s1["error"] == 0 and s1["position_checksum"] == s2["position_checksum"]
    assert x1[:, 1].min() > -512.0 + 0.15      # the floor holds
    assert np.isfinite(x1).all()
    # columns keep their footprint: nothing was pushed sideways by more than a few centimetres
    assert np.abs(x1[:, [0, 2]] - pos[:, [0, 2]]).max() < 0.1


def test_islands_are_a_pure_rescheduling(gpx, scenes, monkeypatch):
    """Solving small islands inside warps must not change a single bit against sending every manifold through the
    grid-wide colour phases (GPX_WIDE_NO_ISLANDS=1, a debugging switch read at every tick)."""
    def run(no_islands):
        if no_islands:
            monkeypatch.setenv("GPX_WIDE_NO_ISLANDS", "1")
        else:
            monkeypatch.delenv("GPX_WIDE_NO_ISLANDS", raising=False)
        n = 80
        g = gpx.World(worlds=1, max_bodies=n)
        for pos, tris in scenes.load_static("stacked"):
            g.add_mesh(pos, tris)
        g.commit()
        rng = np.random.default_rng(2)
        for ix in range(5):
            for iz in range(4):
                for k in range(3):
                    g.create(gpx.body_desc(position=(-1.6 + 0.8 * ix, -1.25 + 0.43 * k, -2.6 + 0.7 * iz),
                                           angular_velocity=tuple(rng.uniform(-0.5, 0.5, 3))))
        for _ in range(90):
            assert g.step() == 0
        assert g.sync() == 0
        c = g.wide_counters()
        return g.transforms()[0, :60].copy(), g.velocities()[0, :60].copy(), c
    xa, va, ca = run(False)
    xb, vb, cb = run(True)
    assert ca["small_islands"] > 0 and cb["small_islands"] == 0
    assert np.array_equal(xa.view(np.uint32), xb.view(np.uint32)) and np.array_equal(va.view(np.uint32), vb.view(np.uint32))


def test_wide_kernels_for_an_ensemble_of_mid_size_worlds(gpx, orc, scenes):
    """GPX_WORLD_WIDE with worlds > 1: 12 independent worlds, each the 4 x 4 x 4 block of boxes tumbling onto stacked.gmap
    (~150 manifolds per world — far more than the lanes the fused ensemble kernel has per world), in ONE global
    sort-and-sweep / pair / island pipeline.  Every world must match its own oracle (wide mode) bit for bit, and the
    worlds must not see each other although their bodies occupy the same space."""
    W, n = 12, 64
    g = gpx.World(worlds=W, max_bodies=n, wide=True)
    os_ = [orc.World(n, wide=True) for _ in range(W)]
    for pos, tris in scenes.load_static("stacked"):
        g.add_mesh(pos, tris)
        for o in os_:
            o.add_mesh(pos, tris)
    g.commit()
    pos = scenes.block_positions(4, 4, 4, 0.45)
    vel = scenes.ensemble_velocities(W, n)
    g.create_all([gpx.body_desc(position=tuple(p)) for p in pos], linvel=vel)
    for wi, o in enumerate(os_):
        for k, p in enumerate(pos):
            o.create(orc.body_desc(position=tuple(p), linear_velocity=tuple(vel[wi, k])))
    for tick in range(1, 61):
        assert g.step() == 0
        for o in os_:
            assert o.step() == 0
        if tick in (1, 5, 20, 60):
            assert g.sync() == 0
            xg, vg = g.transforms(), g.velocities()
            for wi, o in enumerate(os_):
                xo, vo = o.state(n)
                assert np.array_equal(xg[wi].view(np.uint32), xo.view(np.uint32)), f"tick {tick} world {wi}: transforms differ"
                assert np.array_equal(vg[wi].view(np.uint32), vo.view(np.uint32)), f"tick {tick} world {wi}: velocities differ"
    assert not np.array_equal(xg[0], xg[1])                          # the worlds did evolve differently
    # rays and getters address worlds as in any ensemble
    r = np.zeros(W, gpx.RAY_DTYPE)
    r["origin"], r["dir"], r["tmax"] = (0.0, 3.0, -1.5), (0.0, -1.0, 0.0), 10.0
    r["mask"] = 0b11 | (np.arange(W, dtype=np.uint32) << 16)
    hg = g.raycast(r)
    for wi, o in enumerate(os_):
        r1 = r[wi:wi + 1].copy()
        r1["mask"] = 0b11
        ho = o.raycast(r1)[0]
        assert hg["body"][wi] == ho["body"] and hg["fraction"][wi] == ho["fraction"] and hg["world"][wi] == wi
    g.enable_events()                # per-world event lists: tests/test_gpu_events.py::test_events_of_an_ensemble_of_wide_worlds
