"""The fixed-tick step loop (engine/src/subsystem/threads/PhysicsThread.c:59-159): host logic, no GPU.

What the engine's loop guarantees and these tests pin: 60 Hz pacing with idle time counted into `delta`, `delta` in units
of one target tick and clamped to 6 (the 10 ticks/s floor), the frame counter (reset by SetFunction, advanced even
without a function), queued input events delivered at the start of the next tick, the tick mutex excluding the fixed
update, and — for headless runs — a pinned delta of exactly 1 with no sleeping."""
import ctypes as C
import threading
import time

import pytest


def _stop(L):
    """Clear the function and wait for a tick that may still be running it (the callback object must outlive it)."""
    L.gpx_thread_set_function(None)
    L.gpx_thread_lock_tick_mutex()
    L.gpx_thread_unlock_tick_mutex()


@pytest.fixture()
def loop(gpx):
    L = gpx.lib()
    assert L.gpx_thread_init(None) == 0
    yield L
    L.gpx_thread_terminate()


def test_sixty_hertz_pacing_and_delta_normalisation(gpx, loop):
    deltas = []
    fn = gpx.FIXED_UPDATE_FN(lambda state, delta: deltas.append(delta))
    loop.gpx_thread_set_function(fn)
    time.sleep(0.6)
    _stop(loop)
    n = len(deltas)
    assert 28 <= n <= 40, f"{n} ticks in 0.6 s"                      # 36 at exactly 60 Hz
    assert deltas[0] == 1.0                                        # lastTickTime starts at the target (PhysicsThread.c:61)
    steady = deltas[5:]
    assert 0.9 < sum(steady) / len(steady) < 1.25                  # work + idle add up to one tick
    assert abs(loop.gpx_thread_last_tick_ns() / 1e9 - 1 / 60) < 4e-3


def test_slow_ticks_stretch_delta_up_to_the_clamp(gpx, loop):
    deltas = []

    def update(state, delta):
        deltas.append(delta)
        k = len(deltas)
        if k == 3:
            time.sleep(0.05)       # a 50 ms tick -> the next delta is about 3
        if k == 6:
            time.sleep(0.25)       # a 250 ms tick -> clamped to 100 ms = 6 ticks

    fn = gpx.FIXED_UPDATE_FN(update)
    loop.gpx_thread_set_function(fn)
    time.sleep(0.6)
    _stop(loop)
    assert len(deltas) >= 8
    assert 2.7 < deltas[3] < 3.6
    assert deltas[6] == pytest.approx(6.0, abs=1e-9)
    assert all(d <= 6.0 for d in deltas)


def test_frame_counter_and_idle_loop(gpx, loop):
    time.sleep(0.12)
    idle_frames = loop.gpx_thread_frame()
    assert 4 <= idle_frames <= 10                                    # no function: frames still advance at 60 Hz
    fn = gpx.FIXED_UPDATE_FN(lambda s, d: None)
    loop.gpx_thread_set_function(fn)                               # resets physicsFrame (PhysicsThread.c:133)
    assert loop.gpx_thread_frame() <= 1
    time.sleep(0.1)
    assert 4 <= loop.gpx_thread_frame() <= 9
    _stop(loop)


def test_input_events_are_copied_and_delivered_before_the_tick(gpx, loop):
    log = []
    handler = gpx.INPUT_EVENT_FN(lambda state, ev, size: log.append(("event", C.string_at(ev, size))))
    fn = gpx.FIXED_UPDATE_FN(lambda state, delta: log.append(("tick", None)))
    loop.gpx_thread_set_input_handler(handler)
    loop.gpx_thread_lock_tick_mutex()                              # hold the loop at the top of an iteration
    loop.gpx_thread_set_function(fn)
    buf = C.create_string_buffer(b"key-W-down")
    loop.gpx_thread_queue_input_event(buf, 10)
    buf.value = b"overwritten"                                     # the queue holds its own copy
    loop.gpx_thread_unlock_tick_mutex()
    time.sleep(0.1)
    _stop(loop)
    loop.gpx_thread_set_input_handler(None)
    kinds = [k for k, _ in log]
    assert ("event", b"key-W-down") in log
    assert kinds.index("event") < kinds.index("tick") or kinds[0] == "tick" and kinds[1] == "event"
    assert kinds.count("event") == 1


def test_tick_mutex_excludes_the_fixed_update(gpx, loop):
    ticks = []
    fn = gpx.FIXED_UPDATE_FN(lambda s, d: ticks.append(time.perf_counter()))
    loop.gpx_thread_set_function(fn)
    time.sleep(0.08)
    loop.gpx_thread_lock_tick_mutex()                              # what ChangeMap does (GlobalState.c:179-192)
    t0 = time.perf_counter()
    n0 = len(ticks)
    time.sleep(0.15)
    assert len(ticks) == n0                                        # nothing ran while the mutex was ours
    loop.gpx_thread_unlock_tick_mutex()
    time.sleep(0.08)
    _stop(loop)
    assert len(ticks) > n0 and min(t for t in ticks[n0:]) >= t0 + 0.15 - 1e-3


def test_pinned_delta_runs_flat_out_with_delta_one(gpx, loop):
    deltas = []
    done = threading.Event()

    def update(state, delta):
        deltas.append(delta)
        if len(deltas) == 600:
            done.set()

    fn = gpx.FIXED_UPDATE_FN(update)
    loop.gpx_thread_set_pinned_delta(1)
    t0 = time.perf_counter()
    loop.gpx_thread_set_function(fn)
    assert done.wait(5.0)                                          # 600 ticks = 10 s of simulated time in well under 5 s
    _stop(loop)
    assert time.perf_counter() - t0 < 5.0
    assert set(deltas[:600]) == {1.0}


def test_terminate_joins_and_allows_restart(gpx):
    L = gpx.lib()
    assert L.gpx_thread_init(None) == 0
    assert L.gpx_thread_init(None) != 0                            # one loop per process, as in the engine
    L.gpx_thread_terminate()
    assert L.gpx_thread_init(None) == 0
    L.gpx_thread_terminate()
    L.gpx_thread_terminate()                                       # idempotent
