"""GPU parity: the fused tick kernel through the C ABI vs the CPU oracle on the same seeded inputs.

Bars (north_star): positions within 1e-4 m and orientations within 1e-4 rad after 600 ticks on non-chaotic scenes;
stacked scenes settle to the same rest state within 1e-3 m.  Oracle and kernel evaluate identical fp32 expressions
in the same constraint order, so trajectories are additionally required to be bit-identical.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

POS_TOL, ROT_TOL, REST_TOL = 1e-4, 1e-4, 1e-3


def _quat_angle(qa, qb):
    """Rotation angle between two orientations (atan2 form: exact near zero, unlike acos of a float32 dot)."""
    qa, qb = qa.astype(np.float64), qb.astype(np.float64)
    va, wa, vb, wb = qa[..., :3], qa[..., 3:], qb[..., :3], qb[..., 3:]
    vec = wa * vb - wb * va - np.cross(va, vb)
    dot = np.abs(np.sum(qa * qb, axis=-1))
    return 2 * np.arctan2(np.linalg.norm(vec, axis=-1), dot)


def _assert_state(xg, xo, what, exact=True):
    dp = np.abs(xg[..., :3] - xo[..., :3]).max()
    da = _quat_angle(xg[..., 3:], xo[..., 3:]).max()
    assert dp <= POS_TOL, f"{what}: position differs by {dp} m (bar {POS_TOL})"
    assert da <= ROT_TOL, f"{what}: orientation differs by {da} rad (bar {ROT_TOL})"
    if exact:
        assert np.array_equal(xg.view(np.uint32), xo.view(np.uint32)), f"{what}: not bit-identical (max dp {dp})"


def _pair(gpx, orc, scenes, static="stacked", max_bodies=8, worlds=1, **kw):
    g = gpx.World(worlds=worlds, max_bodies=max_bodies, **kw)
    os_ = [orc.World(max_bodies, **{k: v for k, v in kw.items() if k != "max_manifolds"}) for _ in range(worlds)]
    if static:
        for pos, tris in scenes.load_static(static):
            g.add_mesh(pos, tris)
            for o in os_:
                o.add_mesh(pos, tris)
    g.commit()
    return g, os_


def test_free_fall_matches_oracle_and_closed_form(gpx, orc, scenes):
    g, (o,) = _pair(gpx, orc, scenes, static=None)
    d = gpx.body_desc(position=(0, 10, 0), linear_velocity=(1, 0, -2), angular_velocity=(0.5, 1.0, -0.25))
    assert g.create(d) == 0 and o.create(d) == 0
    for _ in range(60):
        assert g.step() == 0 and o.step() == 0
    xg = g.transforms()[0, :1]
    xo, _ = o.state(1)
    _assert_state(xg, xo, "free fall 60 ticks")
    v, y, h = 0.0, 10.0, 1.0 / 120.0
    for _ in range(120):
        v = (v - 9.81 * h) * (1 - 0.05 * h)
        y += v * h
    assert abs(xg[0, 1] - y) < 1e-4


def test_stack8_600_ticks_matches_oracle_and_golden(gpx, orc, scenes):
    """BASELINE config 2: 8-box column on stacked.gmap, 600 ticks."""
    g, (o,) = _pair(gpx, orc, scenes)
    for p in scenes.stack_positions(8):
        d = gpx.body_desc(position=tuple(p))
        assert g.create(d) == o.create(d)
    gold = np.load(scenes.GOLDEN + "/oracle_stack8.npz")
    for tick in range(1, 601):
        assert g.step() == 0 and o.step() == 0
        if tick in (1, 10, 60, 600):
            xg = g.transforms()[0]
            xo, vo = o.state(8)
            _assert_state(xg, xo, f"stack8 tick {tick}")
            _assert_state(xg, gold[f"xf_{tick}"], f"stack8 golden tick {tick}")
            assert np.array_equal(g.velocities()[0].view(np.uint32), vo.view(np.uint32))
    # rest state: each box sits 0.4 above the one below, bottom box on the floor at y = -1.5
    y = g.transforms()[0, :, 1]
    assert np.abs(y - (-1.3 + 0.4 * np.arange(8))).max() < 5e-3
    assert np.abs(g.velocities()[0]).max() < 1e-3
    assert np.abs(xg[:, 1] - xo[:, 1]).max() <= REST_TOL


def test_ensemble_worlds_match_independent_oracles(gpx, orc, scenes):
    """BASELINE config 5 in miniature: worlds with different initial velocities stay independent and each one
    matches its own oracle world."""
    W = 12
    g, os_ = _pair(gpx, orc, scenes, worlds=W)
    vel = scenes.ensemble_velocities(W, 8)
    descs = [gpx.body_desc(position=tuple(p)) for p in scenes.stack_positions(8)]
    ids = g.create_all(descs, linvel=vel)
    assert list(ids) == list(range(8))
    for wi, o in enumerate(os_):
        for k, d in enumerate(descs):
            dd = gpx.body_desc(position=tuple(scenes.stack_positions(8)[k]), linear_velocity=tuple(vel[wi, k]))
            assert o.create(dd) == k
    for tick in range(1, 241):
        assert g.step() == 0
        for o in os_:
            assert o.step() == 0
        if tick in (1, 30, 240):
            xg = g.transforms()
            xo = np.stack([o.state(8)[0] for o in os_])
            _assert_state(xg, xo, f"ensemble tick {tick}")
    st = g.stats()
    assert (st["error"] == 0).all() and (st["ticks"] == 240).all() and (st["awake_bodies"] == 8).all()
    assert len(set(st["position_checksum"].tolist())) == W   # different worlds, different trajectories


@pytest.mark.parametrize("cap,tile_case", [(8, "tile8"), (16, "tile16"), (32, "tile32"), (40, "tile32-strided")])
def test_tile_widths_give_identical_results(gpx, orc, scenes, cap, tile_case):
    """The same scene run with every lane-per-world width must produce the same bits (fixed solve order)."""
    g, (o,) = _pair(gpx, orc, scenes, max_bodies=cap, max_manifolds=96)
    pos = scenes.block_positions(2, 2, 2, 0.42)
    rng = np.random.default_rng(5)
    for p in pos:
        d = gpx.body_desc(position=tuple(p), linear_velocity=tuple(rng.uniform(-0.5, 0.5, 3)),
                          angular_velocity=tuple(rng.uniform(-1, 1, 3)))
        assert g.create(d) == o.create(d)
    for _ in range(90):
        assert g.step() == 0 and o.step() == 0
    _assert_state(g.transforms()[0, :len(pos)], o.state(len(pos))[0], f"{tile_case} 90 ticks")


def test_spheres_boxes_kinematic_and_restitution(gpx, orc, scenes):
    g, (o,) = _pair(gpx, orc, scenes, max_bodies=16)
    descs = [
        gpx.body_desc(shape=gpx.SHAPE_SPHERE, half_extents=(0.4, 0, 0), position=(0.5, 0.0, -1.5), mass=15, restitution=0.5),
        gpx.body_desc(shape=gpx.SHAPE_SPHERE, half_extents=(0.3, 0, 0), position=(0.55, 1.0, -1.45), mass=5),
        gpx.body_desc(position=(-0.5, -1.0, -1.5), rotation=(0.1305262, 0, 0.1305262, 0.9828), friction=0.5),
        gpx.body_desc(half_extents=(0.5, 0.1, 0.5), position=(-0.5, -0.3, -1.5), mass=20),
        gpx.body_desc(half_extents=(0.6, 0.05, 0.6), position=(1.5, -1.2, -1.5), motion_type=gpx.MOTION_KINEMATIC,
                      layer=gpx.LAYER_DYNAMIC, linear_velocity=(-0.3, 0.0, 0.0)),
        gpx.body_desc(position=(1.5, -0.8, -1.5)),
        gpx.body_desc(position=(0.0, -1.0, 0.5), allowed_dofs=1 | 2 | 4 | 16, mass=15, angular_velocity=(1, 2, 3)),
        gpx.body_desc(half_extents=(0.25, 0.25, 0.25), position=(0.0, -1.0, -1.5), layer=gpx.LAYER_SENSOR,
                      motion_type=gpx.MOTION_STATIC, is_sensor=1),
        gpx.body_desc(shape=gpx.SHAPE_EMPTY, position=(0, 0, 0), motion_type=gpx.MOTION_STATIC, layer=gpx.LAYER_STATIC),
    ]
    for d in descs:
        assert g.create(d) == o.create(d)
    for tick in range(1, 301):
        assert g.step() == 0 and o.step() == 0
        if tick in (1, 20, 120, 300):
            _assert_state(g.transforms()[0, :len(descs)], o.state(len(descs))[0], f"mixed scene tick {tick}")
    x = g.transforms()[0]
    assert x[7, 1] == np.float32(-1.0)          # static sensor never moves
    assert abs(x[4, 0] - (1.5 - 0.3 * 5.0)) < 1e-3   # kinematic platform moved by its velocity
    assert abs(x[6, 3]) < 1e-6 and abs(x[6, 5]) < 1e-6  # DOF-locked body only turns about Y


def test_create_destroy_set_between_ticks(gpx, orc, scenes):
    """Body create/destroy and the velocity/position setters arrive between ticks (Actor.c:32-78, Door.c:59-96)."""
    g, (o,) = _pair(gpx, orc, scenes)
    a = gpx.body_desc(position=(0, -1.0, -1.5))
    b = gpx.body_desc(position=(0, -0.5, -1.5), user_data=0xABCDEF)
    assert g.create(a) == o.create(a) == 0
    assert g.create(b) == o.create(b) == 1
    assert g.user_data(1) == 0xABCDEF
    for _ in range(30):
        assert g.step() == 0 and o.step() == 0
    g.destroy(0)
    o.destroy(0)
    g.set_velocity(1, (0.2, 1.0, 0.0), (0, 0.5, 0))
    import ctypes as C
    o.L.orc_body_set_velocity(o.h, 1, (C.c_float * 3)(0.2, 1.0, 0.0), (C.c_float * 3)(0, 0.5, 0))
    for _ in range(30):
        assert g.step() == 0 and o.step() == 0
    c = gpx.body_desc(position=(0.1, 0.5, -1.5))
    assert g.create(c) == o.create(c) == 0        # slot 0 is reused
    for _ in range(60):
        assert g.step() == 0 and o.step() == 0
    _assert_state(g.transforms()[0, :2], o.state(2)[0], "after destroy/create/set")
    # getters come from the host mirror
    g.sync()
    assert np.array_equal(g.get_transform(1), g.transforms()[0, 1])


def test_gmap_loader_feeds_the_same_static_soup(gpx, orc, scenes):
    """The C loader (container body -> device soup) against the fixture decoded by the Python tool."""
    import gasset
    for name in ("stacked", "test", "shapes"):
        _, _, body = gasset.read_container(open(f"{scenes.GOLDEN}/{name}_min.gmap", "rb").read())
        g = gpx.World(worlds=1, max_bodies=8)
        meshes = scenes.load_static(name)
        assert g.load_gmap(body) == sum(1 for _, t in meshes if len(t))
        g.commit()
        o = orc.World(8)
        for pos, tris in meshes:
            if len(tris):
                o.add_mesh(pos, tris)
        rays = scenes.shapes_rays(4000, np.array([p for p, _ in meshes]))
        hg, ho = g.raycast(rays), o.raycast(rays)
        assert np.array_equal(hg.view(np.uint8), ho.view(np.uint8))


def test_gmap_asset_file_is_decoded_by_the_library(gpx, orc, scenes, tmp_path):
    """Container header + gzip + map layout in C (AssetReader.c:150-257 + MapLoader.c:200-273), including the
    container's integrity checks."""
    path = f"{scenes.GOLDEN}/shapes_min.gmap"
    g = gpx.World(worlds=1, max_bodies=8)
    meshes = scenes.load_static("shapes")
    assert g.load_gmap_file(path) == sum(1 for _, t in meshes if len(t))
    g.commit()
    o = orc.World(8)
    for pos, tris in meshes:
        if len(tris):
            o.add_mesh(pos, tris)
    rays = scenes.shapes_rays(4000, np.array([p for p, _ in meshes]))
    assert np.array_equal(g.raycast(rays).view(np.uint8), o.raycast(rays).view(np.uint8))
    blob = open(path, "rb").read()
    for bad in (b"XXXX" + blob[4:], blob[:-5], blob[:4] + b"\x03" + blob[5:]):
        p = tmp_path / "bad.gmap"
        p.write_bytes(bad)
        with pytest.raises(gpx.GpxError):
            gpx.World(worlds=1, max_bodies=8).load_gmap_file(str(p))
    with pytest.raises(gpx.GpxError):
        g.load_gmap_file(str(tmp_path / "missing.gmap"))


def test_manifold_capacity_overflow_reports_error(gpx, scenes):
    """JPH_PhysicsUpdateError_ContactConstraintsFull analogue: the step reports instead of corrupting memory."""
    g = gpx.World(worlds=1, max_bodies=16, max_manifolds=4)
    for pos, tris in scenes.load_static("stacked"):
        g.add_mesh(pos, tris)
    g.commit()
    for p in scenes.block_positions(2, 2, 4, 0.38):
        g.create(gpx.body_desc(position=tuple(p)))
    g.step()
    assert g.sync() & 4


def test_sixty_four_bodies_in_one_world(gpx, orc, scenes):
    """The largest ensemble world: 64 bodies on 32 lanes (two bodies per lane), 4 x 4 x 4 block collapsing."""
    g, (o,) = _pair(gpx, orc, scenes, max_bodies=64, max_manifolds=384)
    rng = np.random.default_rng(21)
    for p in scenes.block_positions(4, 4, 4, 0.43):
        d = gpx.body_desc(position=tuple(p), linear_velocity=tuple(rng.uniform(-0.3, 0.3, 3)))
        assert g.create(d) == o.create(d)
    for tick in range(1, 61):
        assert g.step() == 0 and o.step() == 0
        if tick in (1, 15, 60):
            _assert_state(g.transforms()[0], o.state(64)[0], f"64 bodies tick {tick}")


def test_empty_and_ragged_worlds(gpx, orc, scenes):
    """Worlds of one ensemble may hold different numbers of bodies, including none; an empty map is legal too."""
    W = 5
    g, os_ = _pair(gpx, orc, scenes, worlds=W)
    pos = scenes.stack_positions(8)
    for wi, o in enumerate(os_):
        for k in range(wi * 2):                       # 0, 2, 4, 6, 8 bodies
            d = gpx.body_desc(position=tuple(pos[k]))
            assert g.create(d, world=wi) == o.create(d) == k
    for tick in range(1, 91):
        assert g.step() == 0
        for o in os_:
            assert o.step() == 0
    x = g.transforms()
    for wi, o in enumerate(os_):
        n = wi * 2
        if n:
            _assert_state(x[wi, :n], o.state(n)[0], f"ragged world {wi}")
    st = g.stats()
    assert list(st["awake_bodies"]) == [0, 2, 4, 6, 8] and (st["error"] == 0).all()
    # no static geometry at all: bodies fall freely, rays miss
    e = gpx.World(worlds=2, max_bodies=8)
    e.commit()
    assert e.create(gpx.body_desc(position=(0, 0, 0)), world=1) == 0
    for _ in range(30):
        assert e.step() == 0
    y = e.transforms()[:, 0, 1]
    assert y[0] == 0.0 and y[1] < -1.0


def test_destroy_everything_then_reuse(gpx, orc, scenes):
    g, (o,) = _pair(gpx, orc, scenes)
    for p in scenes.stack_positions(8):
        d = gpx.body_desc(position=tuple(p))
        assert g.create(d) == o.create(d)
    for _ in range(20):
        assert g.step() == 0 and o.step() == 0
    for b in range(8):
        g.destroy(b)
        o.destroy(b)
    for _ in range(5):
        assert g.step() == 0 and o.step() == 0
    assert g.stats()["awake_bodies"][0] == 0 and g.stats()["manifolds"][0] == 0
    assert g.create(gpx.body_desc(position=(0.0, -1.0, -1.5))) == o.create(orc.body_desc(position=(0.0, -1.0, -1.5))) == 0
    with pytest.raises(gpx.GpxError):
        g.destroy(5)                                  # already gone
    for _ in range(60):
        assert g.step() == 0 and o.step() == 0
    _assert_state(g.transforms()[0, :1], o.state(1)[0], "after wipe and reuse")


def _oracle_columns(orc, scenes, worlds, allow_sleeping=0):
    """Independent oracle worlds for the given GLOBAL world indices of BASELINE config 5."""
    import ctypes as C
    meshes = scenes.load_static("stacked")
    pos = scenes.stack_positions(8)
    ws = []
    for wi in worlds:
        vel = scenes.ensemble_velocities(1, 8, first_world=int(wi))[0]
        o = orc.World(8)
        for p, t in meshes:
            o.add_mesh(p, t)
        for k in range(8):
            o.create(orc.body_desc(position=tuple(pos[k]), linear_velocity=tuple(vel[k]), allow_sleeping=allow_sleeping))
        ws.append(o)
    return ws, (C.c_void_p * len(ws))(*[o.h for o in ws])


def test_the_benchmarked_ensemble_matches_independent_oracle_worlds(gpx, orc, scenes):
    """The configuration bench.py times — 4096 kicked columns, 8 body slots per world, 600 ticks, so the 8-lane launch plus
    the routed 32-lane launch for worlds that topple — compared bit for bit with independent oracle worlds: the first 64
    worlds and every world that showed more than 8 manifolds at a sampled tick (the ones that change launch)."""
    W = 4096
    g = gpx.World(worlds=W, max_bodies=8)
    for pos, tris in scenes.load_static("stacked"):
        g.add_mesh(pos, tris)
    g.commit()
    vel = scenes.ensemble_velocities(W, 8)
    g.create_all([gpx.body_desc(position=tuple(p)) for p in scenes.stack_positions(8)], linvel=vel)
    busy = set()
    snaps = {}
    for tick in range(1, 601):
        assert g.step() == 0
        if tick % 20 == 0:
            st = g.stats()
            assert (st["error"] == 0).all()
            busy.update(np.nonzero(st["manifolds"] > 8)[0].tolist())
        if tick in (200, 400, 600):
            snaps[tick] = (g.transforms().copy(), g.velocities().copy())
    assert busy, "no column toppled: the routed launch was never exercised"
    sample = sorted(set(range(64)) | set(sorted(busy)[:48]))
    ws, arr = _oracle_columns(orc, scenes, sample)
    for tick in (200, 400, 600):
        assert orc.lib().orc_step_many(arr, len(ws), 1.0 / 60.0, 2, 200) == 0
        xg, vg = snaps[tick]
        for o, wi in zip(ws, sample):
            xo, vo = o.state(8)
            _assert_state(xg[wi], xo, f"world {wi} tick {tick}")
            assert np.array_equal(vg[wi].view(np.uint32), vo.view(np.uint32)), f"world {wi} tick {tick}: velocities"


def test_kicked_ensemble_dissipates_and_falls_asleep(gpx, scenes):
    """Size-independent properties of the full ensemble: without sleeping the median of the worlds' fastest body slows down
    from one second to the next and ends below 2 cm/s; with sleeping every column that still stands is asleep after
    five seconds and every world after ten."""
    W = 4096
    for allow in (0, 1):
        g = gpx.World(worlds=W, max_bodies=8)
        for pos, tris in scenes.load_static("stacked"):
            g.add_mesh(pos, tris)
        g.commit()
        vel = scenes.ensemble_velocities(W, 8)
        g.create_all([gpx.body_desc(position=tuple(p), allow_sleeping=allow) for p in scenes.stack_positions(8)], linvel=vel)
        p50 = []
        for tick in range(1, 601):
            assert g.step() == 0
            if tick % 60 == 0:
                st = g.stats()
                p50.append(float(np.median(st["max_speed"])))
                if allow and tick == 300:
                    standing = st["manifolds"] <= 8
                    assert (st["awake_bodies"][standing] == 0).mean() > 0.9
        if allow:
            assert (st["awake_bodies"] == 0).all() and (st["max_speed"] == 0).all()
        else:
            assert all(b <= a for a, b in zip(p50[1:], p50[2:])), p50
            assert p50[-1] < 0.02, p50
