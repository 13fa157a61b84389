"""GPU parity: batched closest-hit rays through the C ABI vs the brute-force oracle.

Bar (north_star): hit/miss and body/face ids bit-exact; distances within 1e-5 relative.  Both sides evaluate the
same Moller-Trumbore expression without FMA contraction, so the distances are in fact required to be bit-identical
here; the 1e-5 bound is asserted as well so a future kernel change reports which bar it broke.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _worlds(gpx, orc, scenes, name, max_bodies=8):
    meshes = scenes.load_static(name)
    g = gpx.World(worlds=1, max_bodies=max_bodies)
    o = orc.World(max_bodies)
    for pos, tris in meshes:
        g.add_mesh(pos, tris)
        o.add_mesh(pos, tris)
    g.commit()
    return g, o, meshes


def _assert_hits_equal(h_gpu, h_ref):
    assert np.array_equal(h_gpu["body"], h_ref["body"]), "hit/miss or body ids differ"
    assert np.array_equal(h_gpu["face"], h_ref["face"]), "face ids differ"
    hit = h_ref["body"] != 0xFFFFFFFF
    rel = np.abs(h_gpu["fraction"][hit] - h_ref["fraction"][hit]) / np.maximum(h_ref["fraction"][hit], 1e-12)
    assert rel.size == 0 or rel.max() <= 1e-5, f"hit distance off by {rel.max()} relative (bar 1e-5)"
    assert np.array_equal(h_gpu["fraction"].view(np.uint32), h_ref["fraction"].view(np.uint32)), "fractions not bit-identical"


@pytest.mark.parametrize("name", ["shapes", "stacked", "test", "orb"])
def test_static_rays_match_oracle(gpx, orc, scenes, name):
    g, o, meshes = _worlds(gpx, orc, scenes, name)
    nt, nn, nb = g.static_info()
    assert nt == sum(len(t) for _, t in meshes) and nb == len(meshes) and nn == max(nt - 1, 1)
    rays = scenes.shapes_rays(20000, np.array([p for p, _ in meshes]))
    _assert_hits_equal(g.raycast(rays), o.raycast(rays, mt=True))


def test_shapes_rays_match_committed_golden(gpx, orc, scenes):
    g, o, meshes = _worlds(gpx, orc, scenes, "shapes")
    rays = scenes.shapes_rays(8192, np.array([p for p, _ in meshes]))
    gold = np.load(scenes.GOLDEN + "/oracle_rays_shapes.npz")["hits"]
    _assert_hits_equal(g.raycast(rays), gold)


def test_empty_batch_empty_map_and_single_triangle(gpx, orc):
    g = gpx.World(worlds=1, max_bodies=8)
    g.commit()
    rays = np.zeros(4, gpx.RAY_DTYPE)
    rays["dir"] = (0, 0, -1)
    rays["tmax"] = 10
    rays["mask"] = gpx.RAYMASK_STATIC
    h = g.raycast(rays)
    assert (h["body"] == gpx.INVALID_BODY).all() and (h["fraction"] == 2.0).all()
    assert len(g.raycast(rays[:0])) == 0
    tri = np.array([[[-1, -1, -5], [1, -1, -5], [0, 1, -5]]], np.float32)
    g.add_mesh((0, 0, 0), tri)
    g.commit()
    o = orc.World(8)
    o.add_mesh((0, 0, 0), tri)
    rays["origin"][1] = (5, 0, 0)        # misses
    rays["dir"][2] = (0, 0, 1)           # points away
    rays["tmax"][3] = 4.0                # too short
    hg, ho = g.raycast(rays), o.raycast(rays)
    _assert_hits_equal(hg, ho)
    assert hg["body"][0] == gpx.STATIC_BODY_BASE and hg["face"][0] == 0 and hg["fraction"][0] == np.float32(0.5)
    assert (hg["body"][1:] == gpx.INVALID_BODY).all()


def test_axis_aligned_and_degenerate_directions(gpx, orc, scenes):
    """Zero direction components exercise the reciprocal-direction handling of the slab test."""
    g, o, meshes = _worlds(gpx, orc, scenes, "shapes")
    pos = np.array([p for p, _ in meshes])
    dirs = np.array([[1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1],
                     [1, 1, 0], [0, -1, 1], [1, 0, -1]], np.float32)
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    rays = np.zeros(len(pos) * len(dirs), gpx.RAY_DTYPE)
    rays["origin"] = np.repeat(pos, len(dirs), axis=0)
    rays["dir"] = np.tile(dirs, (len(pos), 1))
    rays["tmax"] = 50
    rays["mask"] = gpx.RAYMASK_STATIC
    _assert_hits_equal(g.raycast(rays), o.raycast(rays))


def test_rays_hit_dynamic_bodies_and_respect_masks(gpx, orc, scenes):
    """Crosshair ray sees STATIC+DYNAMIC layers; the 'triple' laser sees STATIC only; the body filter needs
    GPX_BODY_BLOCKS_LASERS (engine/src/physics/PlayerPhysics.c:55-77, game/src/actor/prop/Laser.c:40-85)."""
    g, o, meshes = _worlds(gpx, orc, scenes, "stacked")
    descs = [gpx.body_desc(position=(0.0, -1.0, -1.5), rotation=(0, 0.3826834, 0, 0.9238795)),
             gpx.body_desc(position=(1.0, -1.0, -1.5), ray_flags=0),
             gpx.body_desc(shape=gpx.SHAPE_SPHERE, half_extents=(0.4, 0, 0), position=(-1.0, -1.0, -1.5)),
             gpx.body_desc(position=(0.0, -0.2, -1.5), layer=gpx.LAYER_SENSOR, motion_type=gpx.MOTION_STATIC, is_sensor=1)]
    for d in descs:
        assert g.create(d) == o.create(d)
    rng = np.random.default_rng(7)
    n = 6000
    rays = np.zeros(n, gpx.RAY_DTYPE)
    rays["origin"] = np.array([0, -0.5, -1.5], np.float32) + rng.uniform(-1.5, 1.5, (n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    rays["dir"] = d / np.linalg.norm(d, axis=1, keepdims=True)
    rays["tmax"] = 10
    for mask in (gpx.RAYMASK_STATIC_DYNAMIC, gpx.RAYMASK_STATIC, 2,
                 gpx.RAYMASK_STATIC_DYNAMIC | gpx.RAYMASK_REQUIRE_BLOCKS_LASERS):
        rays["mask"] = mask
        hg, ho = g.raycast(rays), o.raycast(rays)
        _assert_hits_equal(hg, ho)
        if mask == gpx.RAYMASK_STATIC_DYNAMIC:
            assert set(np.unique(hg["body"][hg["body"] < 8])) == {0, 1, 2}  # boxes + sphere hit, sensor never
        if mask & gpx.RAYMASK_REQUIRE_BLOCKS_LASERS:
            assert 1 not in hg["body"]


def test_full_size_batch_properties(gpx, orc, scenes):
    """BASELINE config 3 at full size (2^20 rays): checked by sampling against the oracle and by invariants."""
    g, o, meshes = _worlds(gpx, orc, scenes, "shapes")
    n = 1 << 20
    rays = scenes.shapes_rays(n, np.array([p for p, _ in meshes]))
    h = g.raycast(rays)
    hit = h["body"] != gpx.INVALID_BODY
    assert hit.mean() > 0.99                      # closed rooms: almost every ray hits
    assert (h["fraction"][hit] >= 0).all() and (h["fraction"][hit] <= 1).all()
    assert (h["fraction"][~hit] == 2.0).all()
    assert (h["face"][hit] < 512).all() and ((h["body"][hit] - gpx.STATIC_BODY_BASE) < 24).all()
    sel = np.random.default_rng(3).choice(n, 30000, replace=False)
    _assert_hits_equal(h[sel], o.raycast(rays[sel], mt=True))
    # determinism: same batch again, identical bytes
    assert np.array_equal(g.raycast(rays).view(np.uint8), h.view(np.uint8))


def test_engine_style_single_ray(gpx, orc, scenes):
    """CastRay_GAME: origin transform, direction = local -Z, fraction * maxDistance = metres (PlayerPhysics.c:305,404)."""
    g, o, meshes = _worlds(gpx, orc, scenes, "stacked")
    h = g.raycast_transform((0, 0, 0), (0, 0, 0, 1), 10.0)
    assert h["body"] == gpx.STATIC_BODY_BASE + 0 and abs(h["fraction"] * 10.0 - 4.0) < 1e-5  # wall z = -4
    h = g.raycast_transform((0, 0, 0), (0, 0, 0, 1), 3.0)
    assert h["body"] == gpx.INVALID_BODY


def test_async_batch_is_valid_after_the_tick_sync(gpx, orc, scenes):
    """gpx_raycast_batch_async: enqueue only; hits land in the pinned buffer by the next gpx_sync_transforms."""
    g, o, meshes = _worlds(gpx, orc, scenes, "stacked")
    d = gpx.body_desc(position=(0.0, -1.0, -1.5))
    assert g.create(d) == o.create(d)
    rays = scenes.shapes_rays(3000, np.array([p for p, _ in meshes]), mask=gpx.RAYMASK_STATIC_DYNAMIC)
    h_rays = gpx.pinned_array(len(rays), gpx.RAY_DTYPE)
    h_hits = gpx.pinned_array(len(rays), gpx.HIT_DTYPE)
    h_rays[:] = rays
    for _ in range(5):
        ref = o.raycast(rays)                 # rays see the state BEFORE this tick's step, as in MapFixedUpdate
        g.raycast_into_async(h_rays, h_hits)
        assert g.step() == 0 and o.step() == 0
        assert g.sync() == 0
        _assert_hits_equal(np.asarray(h_hits), ref)


def test_async_batches_back_to_back_and_device_sync(gpx, orc, scenes):
    """The hits of an async batch travel back on a copy stream: two batches in a row (the second must not overwrite the
    first one's device buffer before it has left), a step in between, and gpx_device_sync as the joining call."""
    g, o, meshes = _worlds(gpx, orc, scenes, "stacked")
    d = gpx.body_desc(position=(0.0, -1.0, -1.5))
    assert g.create(d) == o.create(d)
    rays_a = scenes.shapes_rays(4000, np.array([p for p, _ in meshes]), mask=gpx.RAYMASK_STATIC_DYNAMIC)
    rays_b = np.ascontiguousarray(rays_a[::-1])
    ha, hb = gpx.pinned_array(len(rays_a), gpx.RAY_DTYPE), gpx.pinned_array(len(rays_b), gpx.RAY_DTYPE)
    oa, ob = gpx.pinned_array(len(rays_a), gpx.HIT_DTYPE), gpx.pinned_array(len(rays_b), gpx.HIT_DTYPE)
    ha[:] = rays_a
    hb[:] = rays_b
    for k in range(4):
        ref_a = o.raycast(rays_a)
        g.raycast_into_async(ha, oa)
        g.raycast_into_async(hb, ob)              # same pre-step state, same device staging buffer
        ref_b = o.raycast(rays_b)
        assert g.step() == 0 and o.step() == 0
        if k & 1:
            g.device_sync()
        else:
            assert g.sync() == 0
        _assert_hits_equal(np.asarray(oa), ref_a)
        _assert_hits_equal(np.asarray(ob), ref_b)


def test_static_model_mesh_from_a_gmdl_is_hit_like_the_same_triangles(gpx, orc, scenes):
    """gpx_static_add_gmdl: laseremitter.gmdl's 310-triangle collision mesh placed like the map's emitters
    (CreateStaticModelShape, ModelLoader.c:345-351) answers rays exactly like the oracle given the same triangles."""
    import gasset
    m = np.load(scenes.GOLDEN + "/models.npz")
    body = gasset.build_gmdl_body(1, m["laseremitter_bb"][:3], m["laseremitter_bb"][3:], tris=m["laseremitter_tris"])
    rot = (0.0, float(np.sin(0.4)), 0.0, float(np.cos(0.4)))
    g = gpx.World(worlds=1, max_bodies=8)
    o = orc.World(8)
    sb = g.add_gmdl((1.0, 0.5, -2.0), body, friction=4.25, rot=rot)
    assert sb == o.add_mesh((1.0, 0.5, -2.0), m["laseremitter_tris"], friction=4.25, rot=rot)
    g.commit()
    o.commit()
    rng = np.random.default_rng(11)
    n = 4096
    rays = np.zeros(n, gpx.RAY_DTYPE)
    rays["origin"] = np.float32([1.0, 0.5, -2.0]) + rng.uniform(-2, 2, (n, 3)).astype(np.float32)
    d = np.float32([1.0, 0.5, -2.0]) + rng.uniform(-0.3, 0.3, (n, 3)).astype(np.float32) - rays["origin"]
    rays["dir"] = d / np.linalg.norm(d, axis=1, keepdims=True)
    rays["tmax"] = 10.0
    rays["mask"] = gpx.RAYMASK_STATIC
    hg, ho = g.raycast(rays), o.raycast(rays)
    assert (hg["body"] != gpx.INVALID_BODY).mean() > 0.5
    assert np.array_equal(hg.view(np.uint8), ho.view(np.uint8))


def test_sphere_casts_match_the_oracle_on_the_shapes_map_and_on_bodies(gpx, orc, scenes):
    """gpx_spherecast_batch (north_star: "ray and shape queries"): 16384 random sphere casts through shapes.gmap with boxes
    and spheres in the way — hit / miss, ids, fractions and normals bit-identical to the oracle's brute force."""
    meshes = scenes.load_static("shapes")
    g = gpx.World(worlds=1, max_bodies=8)
    o = orc.World(8)
    for pos, tris in meshes:
        g.add_mesh(pos, tris)
        o.add_mesh(pos, tris)
    g.commit()
    centre = np.mean([p for p, _ in meshes], axis=0)
    rng = np.random.default_rng(5)
    for k in range(6):
        d = dict(position=tuple(float(x) for x in centre + rng.uniform(-3, 3, 3)), motion_type=0, layer=1,
                 rotation=tuple(float(x) for x in (lambda q: q / np.linalg.norm(q))(rng.normal(size=4))))
        if k % 2:
            d.update(shape=2, half_extents=(0.5, 0, 0))
        else:
            d.update(half_extents=(0.4, 0.7, 0.3))
        assert g.create(gpx.body_desc(**d)) == o.create(orc.body_desc(**d))
    assert g.step() == 0   # the creates reach the device
    o.step()
    n = 16384
    rays = scenes.shapes_rays(n, np.array([p for p, _ in meshes]))
    c = np.zeros(n, gpx.CAST_DTYPE)
    c["origin"] = rays["origin"]
    c["origin"][::3] = (centre + rng.uniform(-4, 4, (len(c[::3]), 3))).astype(np.float32)
    c["dir"] = rays["dir"]
    c["tmax"] = 25.0
    c["mask"] = gpx.RAYMASK_STATIC_DYNAMIC
    c["mask"][::5] = gpx.RAYMASK_STATIC
    c["radius"] = rng.choice(np.float32([0.0, 0.05, 0.25, 0.6]), n)
    hg, ho = g.spherecast(c), o.spherecast(c)
    assert (hg["body"] != gpx.INVALID_BODY).mean() > 0.9 and (hg["body"] < 8).mean() > 0.01 and (hg["fraction"] == 0).mean() > 0.001
    bad = np.nonzero(hg.view(np.uint8).reshape(n, 32) != ho.view(np.uint8).reshape(n, 32))[0]
    assert len(bad) == 0, (bad[:5], hg[bad[:3]], ho[bad[:3]])
