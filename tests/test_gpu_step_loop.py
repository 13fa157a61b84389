"""GPU: the tick driven by the fixed-tick step loop (PhysicsThread.c:59-112 -> MapFixedUpdate, MapPhysics.c:58-119),
with getters served to another thread while it runs."""
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_stack8_under_the_step_loop_matches_the_oracle(gpx, orc, scenes):
    g = gpx.World(worlds=1, max_bodies=8)
    o = orc.World(8)
    for pos, tris in scenes.load_static("stacked"):
        g.add_mesh(pos, tris)
        o.add_mesh(pos, tris)
    g.commit()
    for p in scenes.stack_positions(8):
        d = gpx.body_desc(position=tuple(p))
        assert g.create(d) == o.create(d)
    L = gpx.lib()
    errors, done, seen = [], threading.Event(), []
    TICKS = 300

    def map_fixed_update(state, delta):
        # MapFixedUpdate: deltaTime = (float)delta / PHYSICS_TARGET_TPS, then Update(system, deltaTime, 2, jobs)
        if len(seen) >= TICKS:
            return
        rc = g.step(dt=float(np.float32(delta) / np.float32(60.0))) or g.sync()
        if rc:
            errors.append(rc)
        seen.append(delta)
        if len(seen) == TICKS:
            done.set()

    fn = gpx.FIXED_UPDATE_FN(map_fixed_update)
    assert L.gpx_thread_init(None) == 0
    try:
        L.gpx_thread_set_pinned_delta(1)
        L.gpx_thread_set_function(fn)
        # the render / LOD threads read transforms while the tick thread works (LodThread.c:63, RenderingHelpers.c:110)
        reads = 0
        while not done.wait(0.001):
            x = g.get_transform(7)
            assert np.isfinite(x).all()
            reads += 1
        L.gpx_thread_set_function(None)
        L.gpx_thread_lock_tick_mutex()
        L.gpx_thread_unlock_tick_mutex()
    finally:
        L.gpx_thread_terminate()
    assert not errors and len(seen) == TICKS and set(seen) == {1.0} and reads > 0
    for _ in range(TICKS):
        assert o.step() == 0
    assert np.array_equal(g.transforms()[0].view(np.uint32), o.state(8)[0].view(np.uint32))


def test_getters_never_see_half_a_tick(gpx, scenes):
    """The transform mirror is double-buffered: position and rotation read by another thread belong to one tick.
    A kinematic body moves 1 cm and turns 0.01 rad per tick, so both fields encode the tick they were written in."""
    worlds = 64                                        # a mirror of a few hundred KB: the two copies are not instantaneous
    g = gpx.World(worlds=worlds, max_bodies=64)
    g.commit()
    d = gpx.body_desc(half_extents=(0.5, 0.1, 0.5), motion_type=gpx.MOTION_KINEMATIC, layer=0, position=(0.0, 0.0, 0.0),
                      linear_velocity=(0.6, 0.0, 0.0), angular_velocity=(0.0, 0.6, 0.0), linear_damping=0.0, angular_damping=0.0)
    ids = g.create_all([d] * 64)
    last = (worlds - 1, int(ids[-1]))                  # the last slot: its position arrives long before its rotation
    stop = threading.Event()
    bad, reads = [], [0]

    def reader():
        while not stop.is_set():
            x = g.get_transform(last[1], world=last[0])
            tick_p = x[0] / 0.01
            tick_q = 2.0 * np.arctan2(x[4], x[6]) / 0.01
            reads[0] += 1
            if abs(tick_p - tick_q) > 0.25:
                bad.append((float(tick_p), float(tick_q)))

    t = threading.Thread(target=reader)
    t.start()
    try:
        for _ in range(300):
            assert g.step() == 0
            assert g.sync() == 0
    finally:
        stop.set()
        t.join()
    assert reads[0] > 100
    assert not bad, f"{len(bad)} torn reads, e.g. {bad[:3]}"
    x = g.get_transform(last[1], world=last[0])
    assert abs(x[0] - 3.0) < 1e-3 and abs(2.0 * np.arctan2(x[4], x[6]) - 3.0) < 1e-3
