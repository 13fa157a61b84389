"""CPU suite: the oracle (oracle/orc.c) against analytic known answers and the committed golden vectors.

The reference has no tests or golden files for this path and its Jolt dependency is not buildable here
(PARITY UNPINNED, see oracle/orc.h and DESIGN.md), so the oracle is pinned by physics that has a closed form and by
regression vectors generated with tests/golden/make_golden.py.
"""
import numpy as np
import pytest


def _stacked_world(orc, scenes, n=8, **kw):
    o = orc.World(n, **kw)
    for pos, tris in scenes.load_static("stacked"):
        o.add_mesh(pos, tris)
    return o


# ------------------------------------------------------------------------------------------------ rays

def test_moller_trumbore_known_answers(orc):
    o = orc.World(8)
    tri = np.array([[[-1, -1, -5], [1, -1, -5], [0, 1, -5]]], np.float32)
    o.add_mesh((0, 0, 0), tri)
    rays = np.zeros(5, orc.RAY_DTYPE)
    rays["dir"] = (0, 0, -1)
    rays["tmax"] = 10
    rays["mask"] = orc.RAYMASK_STATIC
    rays["origin"][1] = (0.25, -0.5, 1.0)   # 6 m away
    rays["origin"][2] = (5, 0, 0)           # passes beside the triangle
    rays["dir"][3] = (0, 0, 1)              # points away
    rays["tmax"][4] = 4.0                   # too short
    h = o.raycast(rays)
    assert h["body"][0] == orc.STATIC_BASE and h["face"][0] == 0 and h["fraction"][0] == np.float32(0.5)
    assert h["body"][1] == orc.STATIC_BASE and abs(h["fraction"][1] * 10 - 6.0) < 1e-6
    assert (h["body"][2:] == orc.INVALID).all() and (h["fraction"][2:] == 2.0).all()


def test_ray_hits_box_and_sphere_at_analytic_distance(orc):
    o = orc.World(8)
    o.create(orc.body_desc(position=(0, 0, -3), half_extents=(0.5, 0.5, 0.5), motion_type=orc.MOTION_STATIC,
                           layer=orc.LAYER_DYNAMIC))
    o.create(orc.body_desc(shape=orc.SHAPE_SPHERE, half_extents=(0.4, 0, 0), position=(2, 0, -3),
                           motion_type=orc.MOTION_STATIC, layer=orc.LAYER_DYNAMIC))
    rays = np.zeros(3, orc.RAY_DTYPE)
    rays["dir"] = (0, 0, -1)
    rays["tmax"] = 10
    rays["mask"] = orc.RAYMASK_STATIC_DYNAMIC
    rays["origin"][1] = (2, 0, 0)
    rays["origin"][2] = (2, 0, 0)
    rays["mask"][2] = orc.RAYMASK_STATIC            # dynamic layer filtered out (Laser.c:64-72 "triple" filter)
    h = o.raycast(rays)
    assert h["body"][0] == 0 and abs(h["fraction"][0] * 10 - 2.5) < 1e-6
    assert h["body"][1] == 1 and abs(h["fraction"][1] * 10 - 2.6) < 1e-6
    assert h["body"][2] == orc.INVALID


def test_rays_on_shapes_map_match_committed_golden(orc, scenes):
    meshes = scenes.load_static("shapes")
    o = orc.World(8)
    for pos, tris in meshes:
        o.add_mesh(pos, tris)
    rays = scenes.shapes_rays(8192, np.array([p for p, _ in meshes]))
    gold = np.load(scenes.GOLDEN + "/oracle_rays_shapes.npz")["hits"]
    h = o.raycast(rays, mt=True)
    assert np.array_equal(h.view(np.uint8), gold.view(np.uint8))
    # independent numpy Moller-Trumbore over every triangle for a subset: ids and distances agree
    tris = np.concatenate([t + p for p, t in meshes]).astype(np.float64)
    sub = slice(0, 256)
    o3, d3 = rays["origin"][sub].astype(np.float64), rays["dir"][sub].astype(np.float64)
    e1, e2 = tris[:, 1] - tris[:, 0], tris[:, 2] - tris[:, 0]
    best = np.full(256, np.inf)
    face = np.full(256, -1)
    for i in range(256):
        p = np.cross(d3[i], e2)
        det = np.einsum("ij,ij->i", e1, p)
        ok = np.abs(det) > 1e-12
        inv = np.where(ok, 1.0 / np.where(ok, det, 1), 0)
        tv = o3[i] - tris[:, 0]
        u = np.einsum("ij,ij->i", tv, p) * inv
        q = np.cross(tv, e1)
        v = (q @ d3[i]) * inv
        t = np.einsum("ij,ij->i", e2, q) * inv
        hit = ok & (u >= 0) & (v >= 0) & (u + v <= 1) & (t >= 0) & (t <= 50)
        if hit.any():
            k = np.argmin(np.where(hit, t, np.inf))
            best[i], face[i] = t[k], k
    got_hit = h["body"][sub] != orc.INVALID
    assert np.array_equal(got_hit, face >= 0)
    rel = np.abs(h["fraction"][sub][got_hit] * 50.0 - best[got_hit]) / best[got_hit]
    assert rel.max() < 1e-5
    # ids agree except where two triangles tie within rounding
    same = h["face"][sub][got_hit] == face[got_hit]
    assert same.mean() > 0.98


# ------------------------------------------------------------------------------------------------ tick

def test_free_fall_closed_form(orc):
    o = orc.World(8)
    o.create(orc.body_desc(position=(0, 10, 0), linear_velocity=(1, 0, -2)))
    for _ in range(60):
        assert o.step() == 0
    v, y, x, vx, h = 0.0, 10.0, 0.0, 1.0, 1.0 / 120.0
    for _ in range(120):
        v = (v - 9.81 * h) * (1 - 0.05 * h)
        vx = vx * (1 - 0.05 * h)
        y += v * h
        x += vx * h
    xf, vel = o.get(0)
    assert abs(xf[1] - y) < 1e-4 and abs(xf[0] - x) < 1e-5 and abs(vel[1] - v) < 1e-4


def test_box_rests_on_floor_at_half_extent(orc, scenes):
    o = _stacked_world(orc, scenes)
    o.create(orc.body_desc(position=(0.0, -1.0, -1.5)))
    for _ in range(240):
        assert o.step() == 0
    xf, vel = o.get(0)
    assert abs(xf[1] - (-1.5 + 0.2)) < 5e-3          # floor y = -1.5, half extent 0.2 (penetration slop 0.02 allowed)
    assert np.abs(vel).max() < 0.02
    assert abs(xf[0]) < 1e-3 and abs(xf[2] + 1.5) < 1e-3


def test_stack8_matches_committed_golden(orc, scenes):
    o = _stacked_world(orc, scenes)
    for p in scenes.stack_positions(8):
        o.create(orc.body_desc(position=tuple(p)))
    gold = np.load(scenes.GOLDEN + "/oracle_stack8.npz")
    for tick in range(1, 61):
        assert o.step() == 0
        if tick in (1, 10, 60):
            xf, vel = o.state(8)
            assert np.array_equal(xf.view(np.uint32), gold[f"xf_{tick}"].view(np.uint32))
            assert np.array_equal(vel.view(np.uint32), gold[f"vel_{tick}"].view(np.uint32))


def test_two_body_collision_conserves_linear_momentum(orc):
    """No gravity, no damping: sequential impulses are equal and opposite, so total momentum is conserved."""
    o = orc.World(8, gravity=(0, 0, 0))
    a = orc.body_desc(position=(-0.5, 0, 0), linear_velocity=(1.0, 0, 0), mass=10, linear_damping=0, angular_damping=0)
    b = orc.body_desc(position=(0.5, 0.05, 0.02), linear_velocity=(-0.5, 0, 0), mass=5, linear_damping=0, angular_damping=0)
    o.create(a)
    o.create(b)
    p0 = 10 * 1.0 + 5 * (-0.5)
    for _ in range(90):
        assert o.step() == 0
    _, va = o.get(0)
    _, vb = o.get(1)
    assert abs(10 * va[0] + 5 * vb[0] - p0) < 1e-3
    assert abs(10 * va[1] + 5 * vb[1]) < 1e-3
    assert va[0] < 1.0 and vb[0] > -0.5       # they did collide


def test_kinematic_moves_and_sensor_ignored(orc, scenes):
    o = _stacked_world(orc, scenes)
    o.create(orc.body_desc(half_extents=(0.6, 0.05, 0.6), position=(1.5, -1.2, -1.5), motion_type=orc.MOTION_KINEMATIC,
                           linear_velocity=(-0.3, 0, 0)))
    o.create(orc.body_desc(position=(0.0, -1.25, -1.5), layer=orc.LAYER_SENSOR, motion_type=orc.MOTION_STATIC, is_sensor=1,
                           half_extents=(0.5, 0.25, 0.5)))
    o.create(orc.body_desc(position=(0.0, -0.9, -1.5)))     # falls through the sensor onto the floor
    for _ in range(120):
        assert o.step() == 0
    assert abs(o.get(0)[0][0] - (1.5 - 0.3 * 2.0)) < 1e-4
    assert abs(o.get(2)[0][1] - (-1.3)) < 5e-3


def test_dof_lock_and_mass_override(orc):
    o = orc.World(8, gravity=(0, 0, 0))
    o.create(orc.body_desc(allowed_dofs=1 | 2 | 4 | 16, mass=15, angular_velocity=(1, 2, 3), angular_damping=0))
    for _ in range(30):
        o.step()
    xf, vel = o.get(0)
    assert abs(xf[3]) < 1e-7 and abs(xf[5]) < 1e-7 and abs(xf[4]) > 0.1     # rotates about Y only (TestActor.c:42-46)
    assert vel[3] == 0 and vel[5] == 0 and abs(vel[4] - 2) < 1e-6


def test_contact_capacity_reports_error(orc, scenes):
    o = _stacked_world(orc, scenes, n=16, max_manifolds=4)
    for p in scenes.block_positions(2, 2, 4, 0.38):
        o.create(orc.body_desc(position=tuple(p)))
    assert o.step() & 4      # JPH_PhysicsUpdateError_ContactConstraintsFull analogue (MapPhysics.c:109-113)


def test_step_many_equals_sequential_steps(orc, scenes):
    import ctypes as C
    vel = scenes.ensemble_velocities(6, 8)
    pos = scenes.stack_positions(8)

    def build():
        ws = []
        for wi in range(6):
            o = _stacked_world(orc, scenes)
            for k in range(8):
                o.create(orc.body_desc(position=tuple(pos[k]), linear_velocity=tuple(vel[wi, k])))
            ws.append(o)
        return ws
    a, b = build(), build()
    arr = (C.c_void_p * 6)(*[o.h for o in a])
    assert orc.lib().orc_step_many(arr, 6, 1.0 / 60.0, 2, 20) == 0
    for o in b:
        for _ in range(20):
            o.step()
    for x, y in zip(a, b):
        assert np.array_equal(x.state(8)[0].view(np.uint32), y.state(8)[0].view(np.uint32))


# ------------------------------------------------------------------------------------------------ player character

def _walk(o, v, ticks, gravity=True):
    """MovePlayer + UpdatePlayer for `ticks` ticks: horizontal velocity v, gravity while not on the ground
    (engine/src/physics/PlayerPhysics.c:283-294)."""
    for _ in range(ticks):
        p, vel, ground, _ = o.character_get()
        vy = 0.0
        if gravity and ground != 0:
            vy = float(vel[1]) + (-9.81 / 60.0)
        o.character_set_velocity((v[0], vy, v[2]))
        o.character_update()
    return o.character_get()


def test_character_lands_and_stands_on_the_sector_floor(orc, scenes):
    o = _stacked_world(orc, scenes)
    o.character_create((0.0, 0.0, -1.5))
    p, v, ground, gb = o.character_get()
    assert ground == 3                                   # in air until the first update finds the floor
    p, v, ground, gb = _walk(o, (0.0, 0.0, 0.0), 90)
    # capsule: half height 0.2 + radius 0.25 above the floor at y = -1.5
    assert abs(p[1] - (-1.5 + 0.45)) < 1e-3 and ground == 0 and gb >= orc.STATIC_BASE
    assert abs(p[0]) < 1e-6 and abs(p[2] + 1.5) < 1e-6 and abs(v[1]) < 1e-6


def test_character_walks_and_slides_along_a_wall(orc, scenes):
    o = _stacked_world(orc, scenes)
    o.character_create((0.0, -1.05, -1.5))
    _walk(o, (0.0, 0.0, 0.0), 5)
    # sector 0 is the triangle (0,4) (4,-4) (-4,-4) in xz (mapSources/stacked.json): its z = -4 wall stops the capsule
    p, v, ground, _ = _walk(o, (0.7, 0.0, -3.0), 120)
    assert ground == 0
    assert abs(p[2] - (-4.0 + 0.25)) < 2e-3              # held at radius distance from the wall
    assert p[0] > 0.7                                    # kept sliding along it
    assert abs(p[1] - (-1.05)) < 2e-3


def test_character_is_blocked_by_boxes_and_reports_contacts(orc, scenes):
    o = _stacked_world(orc, scenes)
    o.create(orc.body_desc(position=(1.0, -1.3, -1.5), motion_type=orc.MOTION_STATIC, layer=orc.LAYER_STATIC))          # crate
    o.create(orc.body_desc(position=(-1.0, -1.25, -1.5), half_extents=(0.25, 0.25, 0.25), motion_type=orc.MOTION_STATIC,
                           layer=orc.LAYER_SENSOR, is_sensor=1))                                                        # coin
    o.character_create((0.0, -1.05, -1.5))
    for _ in range(3):
        _walk(o, (0.0, 0.0, 0.0), 1)
        o.step()
    seen = {1: set(), 2: set(), 3: set()}

    def tick(v):
        _walk(o, v, 1)
        o.step()
        for a, b, k in o.events():
            seen[int(k)].add((int(a), int(b)))
    for _ in range(60):
        tick((1.5, 0.0, 0.0))
    p, _, _, _ = o.character_get()
    assert abs(p[0] - (1.0 - 0.2 - 0.25)) < 2e-3         # stopped at the crate's face
    assert (0, 0x3FFFFF) in seen[1] and (0, 0x3FFFFF) in seen[2]
    for _ in range(90):
        tick((-1.5, 0.0, 0.0))
    assert (0, 0x3FFFFF) in seen[3]                      # left the crate
    assert (1, 0x3FFFFF) in seen[1]                      # walked into the coin sensor (not blocked by it)
    p, _, _, _ = o.character_get()
    assert p[0] < -1.0
    assert any(a == 0x3FFFFF and b >= orc.STATIC_BASE for a, b in seen[2])   # standing on the floor mesh throughout


# ---- sleeping (the sleep test)

def test_sleep_test_known_answers(orc):
    """A body drifting in a straight line keeps its test points inside spheres of radius (distance travelled) / 2, so
    it becomes a sleep candidate iff it covers less than 2 x 15 mm in 0.5 s: 0.05 m/s sleeps, 0.07 m/s never does."""
    def run(speed, ticks=90, **kw):
        o = orc.World(4)
        o.create(orc.body_desc(position=(0, 0, 0), linear_velocity=(speed, 0, 0), gravity_factor=0.0, linear_damping=0.0,
                               angular_damping=0.0, allow_sleeping=1, **kw))
        first = None
        for t in range(1, ticks + 1):
            assert o.step() == 0
            if first is None and o.asleep(1)[0]:
                first = t
        return first, o
    first, o = run(0.05)
    assert first == 31                      # tick 1 sets the spheres, then 30 ticks of 1/60 s reach 0.5 s
    x, v = o.state(1)
    assert np.all(v == 0.0) and abs(x[0, 0] - 0.05 * 31 / 60) < 1e-6          # stopped where it fell asleep
    assert run(0.07)[0] is None
    assert run(0.05, is_sensor=1)[0] is None                                   # sensors never sleep
    o2 = orc.World(4)
    o2.create(orc.body_desc(position=(0, 0, 0), gravity_factor=0.0, allow_sleeping=0))
    for _ in range(90):
        o2.step()
    assert not o2.asleep(1)[0]                                                 # allow_sleeping = 0 keeps it awake
    # a spinning body moves its off-centre test points: 1 rad/s on a 0.2 m box sweeps them well past 15 mm
    o3 = orc.World(4)
    o3.create(orc.body_desc(position=(0, 0, 0), angular_velocity=(0, 0, 1.0), gravity_factor=0.0, angular_damping=0.0,
                            allow_sleeping=1))
    for _ in range(90):
        o3.step()
    assert not o3.asleep(1)[0]


def test_island_sleeps_together_and_wakes_in_a_cascade(orc, scenes):
    o = orc.World(16)
    for pos, tris in scenes.load_static("stacked"):
        o.add_mesh(pos, tris)
    for p in scenes.stack_positions(8):
        o.create(orc.body_desc(position=tuple(p), allow_sleeping=1))
    states = []
    for _ in range(60):
        assert o.step() == 0
        states.append(o.asleep(8).sum())
    assert set(states) == {0, 8}                                               # the column is one island: all or nothing
    x0 = o.state(8)[0].copy()
    for _ in range(30):
        o.step()
    assert np.array_equal(o.state(8)[0], x0)
    o.create(orc.body_desc(position=(0.0, 3.0, -1.5), allow_sleeping=1))
    counts = []
    for _ in range(240):
        o.step()
        counts.append(int(o.asleep(9).sum()))
    woke = [c for c in counts if c < 8]
    assert woke[0] == 7 and min(counts) == 0 and counts[-1] == 9               # top box first, then everything, then rest
    assert sorted(woke[:woke.index(0) + 1], reverse=True) == woke[:woke.index(0) + 1]   # one way down the column


# ------------------------------------------------------------------------------------------------ dissipation / rest
# Known answers the restated solver has to meet before any trajectory of it is worth comparing (the first
# restatement left kicked columns whirling at 0.2 m/s for ever): energy leaves a kicked column and does not come back,
# the column ends at rest at the stacked heights, and sliding friction stops a box where Coulomb's law says.

BOX_MASS, BOX_SIDE = 10.0, 0.4


def _kinetic_energy(vel):
    inertia = BOX_MASS * BOX_SIDE * BOX_SIDE / 6.0
    return 0.5 * BOX_MASS * float((vel[:, :3].astype(np.float64) ** 2).sum()) + \
        0.5 * inertia * float((vel[:, 3:].astype(np.float64) ** 2).sum())


def _kicked_column(orc, scenes, world, allow_sleeping):
    """One world of BASELINE config 5: the 8-box column with the bench's Philox kicks of that world index."""
    o = _stacked_world(orc, scenes)
    vel = scenes.ensemble_velocities(1, 8, first_world=world)[0]
    for k, p in enumerate(scenes.stack_positions(8)):
        o.create(orc.body_desc(position=tuple(p), linear_velocity=tuple(vel[k]), allow_sleeping=allow_sleeping))
    return o


@pytest.mark.parametrize("world", [0, 1, 2, 3])
def test_kicked_column_loses_its_energy_and_comes_to_rest(orc, scenes, world):
    """Bodies that may not sleep (the bench's setting): the kinetic energy of the swaying column, taken as the maximum
    over one-second windows (the sway trades kinetic for potential energy within a window), never grows by more than
    5 % from one window to the next after the first second, drops at least a hundredfold within ten seconds and is
    below a millijoule at fifteen; the column still stands."""
    o = _kicked_column(orc, scenes, world, allow_sleeping=0)
    ke = []
    for _ in range(900):
        assert o.step() == 0
        ke.append(_kinetic_energy(o.state(8)[1]))
    win = np.array([max(ke[a:a + 60]) for a in range(0, 900, 60)])
    assert (win[2:] <= 1.05 * win[1:-1]).all(), f"kinetic energy grows again: {win}"
    assert win[9] < 0.01 * win[1], f"too little dissipation: {win}"
    assert win[14] < 1.0e-3, f"still moving after 15 s: {win}"
    xf, vel = o.state(8)
    assert np.abs(vel[:, :3]).max() < 0.02 and np.abs(vel[:, 3:]).max() < 0.02
    assert np.abs(xf[:, 1] - (-1.3 + 0.4 * np.arange(8))).max() < 5e-3


@pytest.mark.parametrize("world", [0, 1, 5, 10, 14])
def test_kicked_column_falls_asleep_at_rest(orc, scenes, world):
    """With sleeping allowed (Jolt's and therefore the engine's default): every kicked column is asleep within five
    seconds — exactly at rest, at the stacked heights — and stays that way."""
    o = _kicked_column(orc, scenes, world, allow_sleeping=1)
    for _ in range(300):
        assert o.step() == 0
    assert o.asleep(8).all()
    xf, vel = o.state(8)
    assert (vel == 0).all() and _kinetic_energy(vel) == 0.0
    assert np.abs(xf[:, 1] - (-1.3 + 0.4 * np.arange(8))).max() < 5e-3
    for _ in range(300):
        assert o.step() == 0
    xf2, vel2 = o.state(8)
    assert o.asleep(8).all() and (vel2 == 0).all() and np.array_equal(xf, xf2)


def test_undisturbed_column_is_at_rest_to_a_millimetre_per_second(orc, scenes):
    """BASELINE config 2: the column set down without kicks is at rest (|v|, |w| < 1e-3) after 600 ticks."""
    o = _stacked_world(orc, scenes)
    for p in scenes.stack_positions(8):
        o.create(orc.body_desc(position=tuple(p)))
    for _ in range(600):
        assert o.step() == 0
    xf, vel = o.state(8)
    assert np.abs(vel).max() < 1.0e-3
    assert np.abs(xf[:, 1] - (-1.3 + 0.4 * np.arange(8))).max() < 5e-3


@pytest.mark.parametrize("friction,v0", [(0.2, 1.0), (0.2, 2.0), (0.05, 1.0)])
def test_sliding_box_stops_at_the_coulomb_distance(orc, scenes, friction, v0):
    """A box pushed to v0 on the map floor (friction 4.25, combined sqrt(f1 f2)) stops after v0^2 / (2 mu g), less the
    half step v0 h / 2 the semi-implicit integrator takes off; it does not turn and does not drift sideways."""
    o = _stacked_world(orc, scenes)
    o.create(orc.body_desc(position=(0.0, -1.3, -1.5), friction=friction))
    for _ in range(60):
        assert o.step() == 0
    x0 = o.state(1)[0][0, :3].copy()
    o.set_velocity(0, v=(v0, 0.0, 0.0))
    for _ in range(240):
        assert o.step() == 0
    xf, vel = o.state(1)
    mu = np.sqrt(friction * 4.25)
    want = v0 * v0 / (2.0 * mu * 9.81) - 0.5 * v0 / 120.0
    slid = float(xf[0, 0] - x0[0])
    assert abs(slid - want) < 0.02 * want + 1.0e-3, (slid, want)
    assert np.abs(vel).max() < 1.0e-6
    assert abs(float(xf[0, 2] - x0[2])) < 1.0e-4 and abs(float(xf[0, 4])) < 1.0e-4


@pytest.mark.parametrize("mu_box,slides", [(0.01, True), (0.2, False)])
def test_box_on_an_incline_slides_with_g_sin_minus_mu_g_cos_or_stays(orc, mu_box, slides):
    """A 20 degree ramp (mesh friction 4.25, combined sqrt(f1 f2)): with mu = 0.206 < tan 20 the box accelerates down the
    slope at g (sin - mu cos) (less the 0.05 linear damping), without tumbling; with mu = 0.92 it stays where it was put."""
    th = np.radians(20.0)
    c, s_ = float(np.cos(th)), float(np.sin(th))
    # ramp: a 12 m x 4 m quad through the origin, descending along +x
    a, b = np.array([-6 * c, 6 * s_, -2.0]), np.array([6 * c, -6 * s_, -2.0])
    quad = np.array([[a, b, b + [0, 0, 4.0]], [b + [0, 0, 4.0], a + [0, 0, 4.0], a]], np.float32)
    o = orc.World(8)
    o.add_mesh((0, 0, 0), quad)
    o.commit()
    n = np.array([s_, c, 0.0])                      # the ramp's normal
    half = np.sqrt(0.5 * (1 - c))                   # rotation about z by -20 degrees
    start = (-4.0 * np.array([c, -s_, 0.0])) + n * 0.2
    o.create(orc.body_desc(position=tuple(float(v) for v in start), rotation=(0.0, 0.0, -float(np.sin(th / 2)), float(np.cos(th / 2))),
                           friction=mu_box, linear_velocity=(0.0, 0.0, 0.0)))
    x0 = o.state(1)[0][0, :3].astype(np.float64)
    for _ in range(60):
        assert o.step() == 0
    xf, vel = o.state(1)
    along = np.array([c, -s_, 0.0])
    moved = float((xf[0, :3] - x0) @ along)
    mu = float(np.sqrt(mu_box * 4.25))
    if slides:
        acc = 9.81 * (s_ - mu * c)
        # v' = acc - 0.05 v  =>  x(1) = acc / k (1 - (1 - exp(-k)) / k)
        k = 0.05
        want = acc / k * (1.0 - (1.0 - np.exp(-k)) / k)
        assert abs(moved - want) < 0.03 * want, (moved, want)
        assert abs(float(vel[0, :3] @ along) - acc / k * (1 - np.exp(-k))) < 0.03 * acc
        assert np.abs(vel[0, 3:]).max() < 0.05                      # slides, does not tumble
        assert abs(float((xf[0, :3] - x0) @ n)) < 2e-3              # stays on the ramp
    else:
        assert abs(moved) < 2e-3 and np.abs(vel).max() < 1e-3


def test_restitution_returns_e_squared_of_the_drop_height(orc, scenes):
    """A sphere with restitution 0.8 dropped 1 m onto the map floor (restitution is the larger of the two, 0.8) leaves it
    with 0.8 of its impact speed and climbs back to 0.64 of the drop height, damping and the discrete contact apart."""
    o = _stacked_world(orc, scenes)
    floor_y = -1.5
    o.create(orc.body_desc(shape=orc.SHAPE_SPHERE, half_extents=(0.2, 0, 0), position=(0.0, floor_y + 0.2 + 1.0, -1.5), restitution=0.8,
                           linear_damping=0.0, angular_damping=0.0))
    ys, vys = [], []
    for _ in range(150):
        assert o.step() == 0
        x, v = o.state(1)
        ys.append(float(x[0, 1]) - (floor_y + 0.2))
        vys.append(float(v[0, 1]))
    hit = int(np.argmax(np.array(vys) > 0))                     # first tick moving up again
    v_in, v_out = -min(vys[:hit]), vys[hit]
    assert abs(v_in - np.sqrt(2 * 9.81 * 1.0)) < 0.05 * v_in
    assert abs(v_out / v_in - 0.8) < 0.03
    apex = max(ys[hit:])
    assert abs(apex - 0.64) < 0.05


def test_free_spin_keeps_its_angular_velocity_as_jolt_does_without_gyroscopic_forces(orc):
    """No gravity, no contacts, no damping: a spinning box keeps its world-frame angular velocity for ten seconds, about a
    principal axis and about a skew one alike — Jolt's default (mApplyGyroscopicForce = false) integrates w without the
    gyroscopic term, so for the skew spin it is w that is constant, not the angular momentum I w; the orientation stays a
    unit quaternion."""
    o = orc.World(8, gravity=(0.0, 0.0, 0.0))
    he = (0.1, 0.2, 0.4)
    o.create(orc.body_desc(position=(0, 0, 0), half_extents=he, angular_velocity=(0.0, 3.0, 0.0), linear_damping=0.0, angular_damping=0.0))
    o.create(orc.body_desc(position=(5, 0, 0), half_extents=he, angular_velocity=(1.0, 2.0, 0.5), linear_damping=0.0, angular_damping=0.0))
    m = 10.0
    inertia = np.array([m / 3 * (he[1] ** 2 + he[2] ** 2), m / 3 * (he[0] ** 2 + he[2] ** 2), m / 3 * (he[0] ** 2 + he[1] ** 2)])

    def momentum(xf, w):
        q = xf[3:].astype(np.float64)
        x, y, z, s = q
        R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * s), 2 * (x * z + y * s)],
                      [2 * (x * y + z * s), 1 - 2 * (x * x + z * z), 2 * (y * z - x * s)],
                      [2 * (x * z - y * s), 2 * (y * z + x * s), 1 - 2 * (x * x + y * y)]])
        return R @ (inertia * (R.T @ w.astype(np.float64)))
    xf, vel = o.state(2)
    l0 = momentum(xf[1], vel[1, 3:])
    w_first = vel[1, 3:].copy()
    for _ in range(600):
        assert o.step() == 0
    xf, vel = o.state(2)
    assert np.allclose(vel[0, 3:], (0.0, 3.0, 0.0), atol=1e-6)
    l1 = momentum(xf[1], vel[1, 3:])
    # Jolt (and this restatement) integrates w without the gyroscopic term: w stays constant in the world frame, so
    # |L| is NOT conserved for a skew spin — what is pinned here is that documented behaviour
    assert np.allclose(vel[1, 3:], w_first, atol=1e-6)
    assert abs(np.linalg.norm(xf[1, 3:]) - 1.0) < 1e-5
    assert np.linalg.norm(l1 - l0) > 1e-3


# ------------------------------------------------------------------------------------------------ stairs (ExtendedUpdate)

ENGINE_EXTENDED_UPDATE = (0.25, 0.25, 0.02, 0.15, float(np.cos(np.radians(75.0))))  # PlayerPhysics.c:439-446


def stair_scene(h):
    """A floor (y = 0, x in [-4, 4]) with a step of height h from x = 1 on; normals face up / towards -x."""
    def quad(a, b, c, d):
        return [[a, b, c], [c, d, a]]
    tris = quad((4, 0, 2), (4, 0, -2), (-4, 0, -2), (-4, 0, 2))
    tris += quad((4, h, 2), (4, h, -2), (1, h, -2), (1, h, 2))
    tris += quad((1, 0, -2), (1, h, -2), (1, h, 2), (1, 0, 2))
    return np.array(tris, np.float32)


def walk(world, ticks, vx, settings, record=None):
    """The engine's MovePlayer + UpdatePlayer loop (PlayerPhysics.c:203-295): horizontal speed vx, gravity while in the air."""
    g = np.float32(-9.81 / 60.0)
    for _ in range(ticks):
        _, v, ground, _ = world.character_get()
        mv = np.zeros(3, np.float32)
        mv[0] = vx
        if ground != 0:
            mv[1] = np.float32(v[1] + g)
        world.character_set_velocity([float(x) for x in mv])
        world.character_update(settings=settings)
        if record is not None:
            record.append(world.character_get())


@pytest.mark.parametrize("height,settings,climbs", [(0.24, None, False), (0.24, ENGINE_EXTENDED_UPDATE, True),
                                                    (0.30, ENGINE_EXTENDED_UPDATE, False)])
def test_character_walks_up_steps_no_higher_than_the_step_height(orc, height, settings, climbs):
    """walkStairsStepUp = 0.25 (PlayerPhysics.c:441): a 0.24 m step stops the plain update, ExtendedUpdate's walk-stairs
    takes it, and a 0.30 m step stays a wall."""
    o = orc.World(8)
    o.add_mesh((0, 0, 0), stair_scene(height))
    o.commit()
    o.character_create((0.0, 0.5, 0.0))
    walk(o, 90, 1.5, settings)
    p, _, ground, _ = o.character_get()
    if climbs:
        assert p[0] > 1.5 and abs(p[1] - (height + 0.45)) < 1e-3 and ground == 0
    else:
        assert abs(p[0] - 0.75) < 1e-3 and abs(p[1] - 0.45) < 1e-3 and ground == 0


def test_character_sticks_to_the_floor_when_walking_down_a_step(orc):
    """stickToFloorStepDown = 0.25 (PlayerPhysics.c:440): walking off a 0.2 m step the character is set down at once and
    never reports being in the air; without it there are airborne ticks."""
    airborne = {}
    for name, settings in (("plain", None), ("extended", ENGINE_EXTENDED_UPDATE)):
        o = orc.World(8)
        o.add_mesh((0, 0, 0), stair_scene(0.2))
        o.commit()
        o.character_create((2.0, 0.2 + 0.45 + 0.01, 0.0))
        rec = []
        walk(o, 20, 0.0, settings)
        walk(o, 80, -1.5, settings, rec)
        p = rec[-1][0]
        assert abs(p[1] - 0.45) < 1e-3 and p[0] < 0.5
        airborne[name] = sum(1 for _, _, ground, _ in rec if ground == 3)
    assert airborne["extended"] == 0 and airborne["plain"] > 0


def test_a_fast_character_does_not_pass_through_a_floor_or_a_wall(orc):
    """Swept motion: at 40 m/s a tick's move (0.67 m) is longer than the capsule's radius; taken in one piece the centre
    would land beyond a zero-thickness floor or wall and be pushed out on the far side.  The move is taken in pieces of
    at most half a radius: the character lands on the floor (centre at half height + radius) and stops at the wall."""
    o = orc.World(8)
    o.add_mesh((0, 0, 0), stair_scene(3.0))           # floor y = 0, a 3 m wall at x = 1 facing -x
    o.commit()
    # falling: starts 10.3 m up so that, at 40 m/s, whole ticks would carry the centre from above the floor to below it
    o.character_create((-2.0, 10.3, 0.0))
    for _ in range(30):
        o.character_set_velocity([0.0, -40.0, 0.0])
        o.character_update()
    p, _, ground, _ = o.character_get()
    assert abs(p[1] - 0.45) < 1e-3 and ground == 0
    # running at the wall
    o.character_set_position([-3.05, 0.45, 0.0])
    for _ in range(30):
        o.character_set_velocity([40.0, 0.0, 0.0])
        o.character_update()
    p, _, ground, _ = o.character_get()
    assert abs(p[0] - 0.75) < 1e-3 and abs(p[1] - 0.45) < 1e-3


# ------------------------------------------------------------------------------------------------ sphere casts

def test_sphere_cast_known_answers(orc):
    """A sphere of radius 0.25 cast down -z at a triangle in the plane z = -5, a 1 m box and a 0.4 m sphere: face hits stop
    a radius early, edge / vertex / corner hits at sqrt(r^2 - offset^2) from the feature, an overlapping start reports 0."""
    o = orc.World(8)
    o.add_mesh((0, 0, 0), np.array([[[-1, -1, -5], [1, -1, -5], [0, 1, -5]]], np.float32))
    o.create(orc.body_desc(position=(3, 0, -5), half_extents=(0.5, 0.5, 0.5), motion_type=0, layer=1))
    o.create(orc.body_desc(shape=2, half_extents=(0.4, 0, 0), position=(6, 0, -5), motion_type=0, layer=1))
    c = np.zeros(9, orc.CAST_DTYPE)
    c["dir"] = (0, 0, -1)
    c["tmax"] = 10
    c["mask"] = orc.RAYMASK_STATIC_DYNAMIC
    c["radius"] = 0.25
    c["origin"] = [(0, 0, 0), (1.1, -1, 0), (3, 0, 0), (3.6, 0, 0), (6, 0, 0), (0, 0, -4.9), (10, 0, 0), (3.6, 0.6, 0), (3, 0, 0)]
    c["mask"][8] = orc.RAYMASK_STATIC           # the box is filtered out by the layer mask
    h = o.spherecast(c)
    t = h["fraction"] * 10
    edge = 0.25 ** 2 - 0.1 ** 2
    assert abs(t[0] - 4.75) < 1e-5 and h["body"][0] == orc.STATIC_BASE and np.allclose(h["normal"][0], (0, 0, 1))
    assert abs(t[1] - (5 - np.sqrt(edge))) < 1e-5 and abs(np.linalg.norm(h["normal"][1]) - 1) < 1e-5
    assert abs(t[2] - 4.25) < 1e-5 and h["body"][2] == 0 and h["face"][2] == 5
    assert abs(t[3] - (4.5 - np.sqrt(edge))) < 1e-5 and h["body"][3] == 0
    assert abs(t[4] - (5 - 0.65)) < 1e-5 and h["body"][4] == 1
    assert t[5] == 0 and h["body"][5] == orc.STATIC_BASE
    assert h["body"][6] == orc.INVALID and h["fraction"][6] == 2.0
    assert abs(t[7] - (4.5 - np.sqrt(0.25 ** 2 - 0.02))) < 1e-5
    assert h["body"][8] == orc.INVALID
    # a cast with radius 0 is a ray
    c["radius"] = 0.0
    r = np.zeros(9, orc.RAY_DTYPE)
    for k in ("origin", "dir", "tmax", "mask"):
        r[k] = c[k]
    hr, hc = o.raycast(r), o.spherecast(c)
    hit = hr["body"] != orc.INVALID
    assert np.array_equal(hr["body"], hc["body"]) and np.allclose(hr["fraction"][hit], hc["fraction"][hit], atol=1e-6)


def test_sweep_candidates_reproduce_the_all_pairs_loop(orc, scenes):
    """Above 256 body slots the oracle takes candidate pairs from a sort-and-sweep; the same 200 bodies in a world of
    200 slots (all-pairs) and one of 300 slots (sweep) must evolve bit-identically."""
    rng = np.random.default_rng(7)
    n = 200
    pos = np.stack([rng.uniform(-1.5, 1.5, n), -512.0 + 0.3 + 0.45 * np.arange(n) / 8, rng.uniform(-1.5, 1.5, n)], axis=1)
    spin = rng.uniform(-3, 3, (n, 3))
    worlds = []
    for slots in (n, 300):
        o = orc.World(slots, wide=True)
        for p, t in scenes.box_map():
            o.add_mesh(p, t)
        for i in range(n):
            kind = orc.SHAPE_SPHERE if i % 5 == 0 else orc.SHAPE_BOX
            o.create(orc.body_desc(shape=kind, position=tuple(pos[i]), linear_velocity=(0.0, -2.0, 0.0),
                                   angular_velocity=tuple(spin[i])))
        worlds.append(o)
    seen = 0
    for tick in range(60):
        for o in worlds:
            assert o.step() == 0
        seen = max(seen, worlds[0].L.orc_manifold_count(worlds[0].h))
        xa, va = worlds[0].state(n)
        xb, vb = worlds[1].state(n)
        assert np.array_equal(xa.view(np.uint32), xb.view(np.uint32)), f"tick {tick}"
        assert np.array_equal(va.view(np.uint32), vb.view(np.uint32)), f"tick {tick}"
    assert seen > 150            # the pile really is in contact


def test_threaded_step_reproduces_the_serial_step(orc, scenes):
    """orc_step_mt (bench.py's CPU baseline for the one large world) against orc_step: a gas of 6000 spheres and boxes
    that collide all over the place, and a lattice that lands in columns — the same bits, tick after tick."""
    rng = np.random.default_rng(11)
    n = 6000
    pos = rng.uniform(-12.0, 12.0, (n, 3)).astype(np.float32)
    pos[:, 1] -= 490.0
    vel = rng.uniform(-6.0, 6.0, (n, 3)).astype(np.float32)
    lat = scenes.lattice_positions(24, 6, 24)
    for kind in ("gas", "lattice"):
        worlds = []
        for _ in range(2):
            count = n if kind == "gas" else len(lat)
            o = orc.World(count, gravity=(0.0, 0.0, 0.0) if kind == "gas" else (0.0, -9.81, 0.0))
            for p, t in scenes.box_map():
                o.add_mesh(p, t)
            for i in range(count):
                if kind == "gas":
                    o.create(orc.body_desc(position=tuple(pos[i]), linear_velocity=tuple(vel[i]), shape=2 if i % 3 == 0 else 1,
                                           half_extents=(0.15, 0.15, 0.15), linear_damping=0.0, allow_sleeping=1))
                else:
                    o.create(orc.body_desc(position=tuple(lat[i]), allow_sleeping=1))
            worlds.append((o, count))
        (a, count), (b, _) = worlds
        contacts = 0
        for tick in range(30):
            assert a.step() == 0 and b.step_mt() == 0
            contacts = max(contacts, a.L.orc_manifold_count(a.h))
            assert a.L.orc_manifold_count(a.h) == b.L.orc_manifold_count(b.h)
            xa, va = a.state(count)
            xb, vb = b.state(count)
            assert np.array_equal(xa.view(np.uint32), xb.view(np.uint32)), f"{kind} tick {tick}"
            assert np.array_equal(va.view(np.uint32), vb.view(np.uint32)), f"{kind} tick {tick}"
        assert contacts > 100


def test_floor_and_wall_manifolds_of_one_body_do_not_feed_each_other(orc, scenes):
    """Fuzz seed 40275 on the CPU: a thin plate knocked into the corner between the floor and a sloping wall of
    stacked.gmap has two manifolds against the SAME static body whose contact points lie a hair apart.  Matching the
    warm start by body pair and point distance alone let each manifold pick up the other's impulses, sub-step after
    sub-step (2.4e6, 2.7e6, 3.4e6 ...), until the plate left at 38 km/s with a NaN orientation — on the GPU and here
    alike.  The slot's opening triangle is part of the key now (Jolt keys its cache by sub-shape pair)."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("fuzz_parity", os.path.join(os.path.dirname(__file__), "fuzz_parity.py"))
    fz = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(fz)
    rng = np.random.default_rng(40275)
    cap = int(rng.integers(80, 200))
    n0 = int(cap * rng.uniform(0.4, 0.9))
    o = orc.World(cap)
    for pos, tris in scenes.load_static("stacked"):
        o.add_mesh(pos, tris)
    area = 2.5 * max(1.0, (cap / 150.0) ** 0.5)
    live = []
    for _ in range(n0):
        live.append(o.create(orc.body_desc(**fz.random_desc(rng, area))))
    for tick in range(1, 221):
        for _ in range(int(rng.integers(0, 3))):
            op = rng.integers(0, 5)
            if op == 0 and len(live) < cap:
                live.append(o.create(orc.body_desc(**fz.random_desc(rng, area))))
            elif op == 1 and len(live) > 2:
                o.destroy(live.pop(int(rng.integers(0, len(live)))))
            elif op == 2 and live:
                b = live[int(rng.integers(0, len(live)))]
                o.set_velocity(b, tuple(float(x) for x in rng.uniform(-3, 3, 3)), tuple(float(x) for x in rng.uniform(-3, 3, 3)))
            elif op == 3 and live:
                b = live[int(rng.integers(0, len(live)))]
                o.set_position(b, (float(rng.uniform(-area, area)), float(rng.uniform(0.0, 2.5)), float(rng.uniform(-area, area) - 1.5)))
                o.wake(b)
        assert o.step() == 0
        x, v = o.state(cap)
        assert np.isfinite(x[live]).all() and np.isfinite(v[live]).all(), f"tick {tick}"
        assert np.abs(v[live][:, :3]).max() < 60.0, f"tick {tick}: {np.abs(v[live][:, :3]).max()} m/s"
