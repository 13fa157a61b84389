#!/usr/bin/env python
"""Differential fuzzing of the CUDA path against the CPU oracle (needs a GPU; test infrastructure).

  python tests/fuzz_parity.py [first_seed] [seeds] [ticks]

Modes (environment):
  default              one world per seed: random boxes and spheres (size, mass, friction, restitution, orientation, spin),
                       kinematic movers, sensors, restricted degrees of freedom, bodies that may or may not sleep, and random
                       host edits between ticks (create, destroy, set velocity, set position).  Even seeds: up to 64 body
                       slots (k_tick); odd seeds: 80..200 (the wide kernels).  FUZZ_CAP=<n> fixes the slot count.
  FUZZ_WORLDS=<w>      an ensemble of w such worlds in one gpx world against one oracle world each (FUZZ_CAP slots, default 8);
                       with GPX_TILE=8|16|32 the narrow tiles and their routing of toppled worlds are forced.
  FUZZ_MODE=queries    20 000 rays and 4000 sphere casts against a shipped map with random bodies.
  FUZZ_MODE=character  the player capsule on a random walk (speed changes, jumps, ExtendedUpdate settings) among random bodies.
  FUZZ_NO=a,b,...      switch features off: dofs, rest, spheres, kin, sensor, sleep, edits, spin, events.
  FUZZ_EVERY_TICK=1    compare after every tick (default: every fourth).   FUZZ_LOG=1  print the last edits on a mismatch.

Transforms, velocities, sleep flags and contact events (ids, faces, fractions, normals for the queries; position, velocity,
ground state and contact list for the character) must be bit-identical.  A scene that overflows a capacity of the device
path (reported through the tick's error code) ends its seed: the oracle has no such limits.
tests/test_gpu_fuzz.py runs a slice of every mode; GPX_NO_BLOCK_TILE, GPX_TILE and a -DGPX_JITTER build of the library
(README) widen what a run exercises."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
gpx = importlib.import_module("c-game-engine_b200")
scenes = importlib.import_module("c-game-engine_b200.scenes")
import orc

OFF = set(filter(None, os.environ.get("FUZZ_NO", "").split(",")))   # dofs, rest, spheres, kin, sensor, sleep, edits, spin
ticks = 240


class CapacityError(Exception):
    pass


def random_desc(rng, area):
    kind = rng.integers(0, 10)
    pos = (float(rng.uniform(-area, area)), float(rng.uniform(-1.0, 2.5)), float(rng.uniform(-area, area) - 1.5))
    d = dict(position=pos, friction=float(rng.uniform(0.0, 1.0)), restitution=float(rng.choice([0.0, 0.0, 0.3, 0.8])),
             mass=float(rng.uniform(0.5, 30.0)), allow_sleeping=int(rng.integers(0, 2)),
             linear_velocity=tuple(float(v) for v in rng.uniform(-2, 2, 3)),
             angular_velocity=tuple(float(v) for v in rng.uniform(-4, 4, 3)))
    if "spheres" in OFF and 6 <= kind < 8:
        kind = 0
    if "kin" in OFF and kind == 8:
        kind = 0
    if "sensor" in OFF and kind == 9:
        kind = 0
    if "rest" in OFF:
        d["restitution"] = 0.0
    if "sleep" in OFF:
        d["allow_sleeping"] = 0
    if "spin" in OFF:
        d["angular_velocity"] = (0.0, 0.0, 0.0)
    if kind < 6:
        d.update(shape=1, half_extents=tuple(float(v) for v in rng.uniform(0.08, 0.4, 3)))
        a = rng.normal(size=4)
        d.update(rotation=tuple(float(v) for v in (a / np.linalg.norm(a)).astype(np.float32)))
    elif kind < 8:
        d.update(shape=2, half_extents=(float(rng.uniform(0.08, 0.35)), 0.0, 0.0))
    elif kind == 8:
        d.update(shape=1, half_extents=(0.5, 0.05, 0.5), motion_type=1, angular_velocity=(0.0, 0.0, 0.0),
                 linear_velocity=(float(rng.uniform(-0.4, 0.4)), 0.0, float(rng.uniform(-0.4, 0.4))))
    else:
        d.update(shape=1, half_extents=(0.4, 0.3, 0.4), layer=3, motion_type=0, is_sensor=1, linear_velocity=(0, 0, 0),
                 angular_velocity=(0, 0, 0))
    if rng.integers(0, 6) == 0 and "dofs" not in OFF:
        d["allowed_dofs"] = int(rng.choice([0b010111, 0b111000 | 0b111, 0b000111]))
    return d


def compare(g, o, cap, live, what):
    """state of the bodies that exist (a destroyed slot reads as zeros on one side and as its last state on the other)"""
    rc = g.sync()
    if rc != 0:
        raise CapacityError(f"{what}: the tick reported error code {rc} (a capacity of the device path, e.g. more than 32 "
                            "manifolds on one body of a wide world); nothing to compare from here on")
    live = np.array(sorted(live), np.int64)
    xg, vg = g.transforms()[0][live], g.velocities()[0][live]
    xo, vo = o.state(cap)
    xo, vo = xo[live], vo[live]
    if not np.array_equal(xg.view(np.uint32), xo.view(np.uint32)):
        bad = np.nonzero(np.any(xg.view(np.uint32) != xo.view(np.uint32), axis=1))[0]
        raise AssertionError(f"{what}: transforms differ at bodies {live[bad[:8]]}: {xg[bad[0]]} vs {xo[bad[0]]}")
    if not np.array_equal(vg.view(np.uint32), vo.view(np.uint32)):
        bad = np.nonzero(np.any(vg.view(np.uint32) != vo.view(np.uint32), axis=1))[0]
        sg, so = g.sleeping()[0][live], o.asleep(cap)[live]
        raise AssertionError(f"{what}: velocities differ at bodies {live[bad[:8]]}: {vg[bad[0]]} vs {vo[bad[0]]}; asleep {sg[bad[0]]} vs {so[bad[0]]}; "
                             f"transform {xg[bad[0]]}")
    sg, so = g.sleeping()[0][live], o.asleep(cap)[live]
    if not np.array_equal(sg, so):
        raise AssertionError(f"{what}: sleep flags differ at {np.nonzero(sg != so)[0][:8]}")


def run(seed):
    rng = np.random.default_rng(seed)
    wide = bool(seed & 1) if "FUZZ_CAP" not in os.environ else int(os.environ["FUZZ_CAP"]) > 64
    cap = int(rng.integers(80, 200)) if wide else int(rng.integers(6, 65))
    cap = int(os.environ.get("FUZZ_CAP", cap))
    n0 = int(cap * rng.uniform(0.4, 0.9))
    meshes = scenes.load_static("stacked")
    g = gpx.World(worlds=1, max_bodies=cap)
    o = orc.World(cap)
    for pos, tris in meshes:
        g.add_mesh(pos, tris)
        o.add_mesh(pos, tris)
    g.commit()
    g.enable_events()
    area = (2.5 * max(1.0, (cap / 150.0) ** 0.5)) if wide else 1.5   # the same crowding whatever the size
    live = []
    for _ in range(n0):
        d = random_desc(rng, area)
        a, b = g.create(gpx.body_desc(**d)), o.create(orc.body_desc(**d))
        assert a == b, (a, b)
        live.append(a)
    errors = 0
    log = []
    for tick in range(1, ticks + 1):
        # host edits
        for _ in range(int(rng.integers(0, 3)) if "edits" not in OFF else 0):
            op = rng.integers(0, 5)
            if op == 0 and len(live) < cap:
                d = random_desc(rng, area)
                a, b = g.create(gpx.body_desc(**d)), o.create(orc.body_desc(**d))
                assert a == b, (a, b)
                live.append(a)
                log.append((tick, 'create', a, d))
            elif op == 1 and len(live) > 2:
                b = live.pop(int(rng.integers(0, len(live))))
                g.destroy(b)
                o.destroy(b)
                log.append((tick, 'destroy', b))
            elif op == 2 and live:
                b = live[int(rng.integers(0, len(live)))]
                v = tuple(float(x) for x in rng.uniform(-3, 3, 3))
                w = tuple(float(x) for x in rng.uniform(-3, 3, 3))
                g.set_velocity(b, v, w)
                o.set_velocity(b, v, w)
                log.append((tick, 'velocity', b, v, w))
            elif op == 3 and live:
                b = live[int(rng.integers(0, len(live)))]
                p = (float(rng.uniform(-area, area)), float(rng.uniform(0.0, 2.5)), float(rng.uniform(-area, area) - 1.5))
                g.set_position(b, p)      # JPH_Activation_Activate
                o.set_position(b, p)
                o.wake(b)
                log.append((tick, 'position', b, p))
        rg, ro = g.step(), o.step()
        if rg != ro:
            raise AssertionError(f"seed {seed} tick {tick}: step returned {rg} (gpu) vs {ro} (oracle)")
        errors += rg != 0
        rc = g.sync()
        if rc != 0:
            raise CapacityError(f"seed {seed} tick {tick}: the tick reported error code {rc} on the GPU")
        eg, eo = g.poll_events(), o.events()
        got = np.stack([eg["body_a"], eg["body_b"], eg["kind"]], axis=1) if len(eg) else np.zeros((0, 3), np.uint32)
        if rg == 0 and "events" not in OFF and not np.array_equal(got, eo):
            sg_, so_ = {tuple(int(v) for v in r) for r in got}, {tuple(int(v) for v in r) for r in eo}
            raise AssertionError(f"seed {seed} tick {tick}: events differ: {len(got)} vs {len(eo)} records; only on the GPU "
                                 f"{sorted(sg_ - so_)[:6]}, only in the oracle {sorted(so_ - sg_)[:6]}; same set, other order: {sg_ == so_}")
        if tick % 4 == 0 or tick < 8 or os.environ.get("FUZZ_EVERY_TICK"):
            try:
                compare(g, o, cap, live, f"seed {seed} ({'wide' if wide else 'ensemble'}, cap {cap}) tick {tick}")
            except AssertionError:
                if os.environ.get("FUZZ_LOG"):
                    for e in log[-12:]:
                        print("edit", e)
                raise
    return wide, cap, len(live), errors


def run_ensemble(seed, worlds, cap):
    """Many small worlds in one gpx world (the benchmarked shape: narrow tiles, toppled worlds routed to the 32-lane
    launch) against one oracle world each; random scenes per world, random edits in random worlds; state of every world
    compared after every tick, contact events too."""
    rng = np.random.default_rng(seed)
    meshes = scenes.load_static("stacked")
    g = gpx.World(worlds=worlds, max_bodies=cap)
    os_ = [orc.World(cap) for _ in range(worlds)]
    for pos, tris in meshes:
        g.add_mesh(pos, tris)
        for o in os_:
            o.add_mesh(pos, tris)
    g.commit()
    events = "events" not in OFF
    try:
        if events:
            g.enable_events()
    except gpx.GpxError:
        events = False          # an ensemble of worlds of more than 64 bodies each reports no events
    live = [[] for _ in range(worlds)]
    for wi in range(worlds):
        for _ in range(int(rng.integers(1, cap + 1))):
            d = random_desc(rng, 1.2)
            a, b = g.create(gpx.body_desc(**d), world=wi), os_[wi].create(orc.body_desc(**d))
            assert a == b, (a, b)
            live[wi].append(a)
    for tick in range(1, ticks + 1):
        for _ in range(int(rng.integers(0, 4)) if "edits" not in OFF else 0):
            wi = int(rng.integers(0, worlds))
            o = os_[wi]
            op = rng.integers(0, 4)
            if op == 0 and len(live[wi]) < cap:
                d = random_desc(rng, 1.2)
                a, b = g.create(gpx.body_desc(**d), world=wi), o.create(orc.body_desc(**d))
                assert a == b, (a, b)
                live[wi].append(a)
            elif op == 1 and len(live[wi]) > 1:
                b = live[wi].pop(int(rng.integers(0, len(live[wi]))))
                g.destroy(b, world=wi)
                o.destroy(b)
            elif op == 2 and live[wi]:
                b = live[wi][int(rng.integers(0, len(live[wi])))]
                v = tuple(float(x) for x in rng.uniform(-3, 3, 3))
                w = tuple(float(x) for x in rng.uniform(-3, 3, 3))
                g.set_velocity(b, v, w, world=wi)
                o.set_velocity(b, v, w)
            elif op == 3 and live[wi]:
                b = live[wi][int(rng.integers(0, len(live[wi])))]
                p = (float(rng.uniform(-1.2, 1.2)), float(rng.uniform(0.0, 2.5)), float(rng.uniform(-1.2, 1.2) - 1.5))
                g.set_position(b, p, world=wi)
                o.set_position(b, p)
                o.wake(b)
        rg = g.step()
        ro = 0
        for o in os_:
            ro |= o.step()
        assert rg == ro, f"seed {seed} tick {tick}: step returned {rg} (gpu) vs {ro} (oracles)"
        assert g.sync() == 0
        xg, vg, sg = g.transforms(), g.velocities(), g.sleeping()
        eg = g.poll_events() if events else None
        for wi, o in enumerate(os_):
            if not live[wi]:
                continue
            idx = np.array(sorted(live[wi]), np.int64)
            xo, vo = o.state(cap)
            what = f"seed {seed} ({worlds} worlds x {cap}) tick {tick} world {wi}"
            if not np.array_equal(xg[wi][idx].view(np.uint32), xo[idx].view(np.uint32)):
                bad = idx[np.nonzero(np.any(xg[wi][idx].view(np.uint32) != xo[idx].view(np.uint32), axis=1))[0]]
                raise AssertionError(f"{what}: transforms differ at bodies {bad[:6]} of {len(idx)}: {xg[wi][bad[0]]} vs {xo[bad[0]]}; "
                                     f"manifolds {o.manifolds()}, stats {g.stats()[wi]}")
            assert np.array_equal(vg[wi][idx].view(np.uint32), vo[idx].view(np.uint32)), f"{what}: velocities differ"
            assert np.array_equal(sg[wi][idx], o.asleep(cap)[idx]), f"{what}: sleep flags differ"
            if rg == 0 and events:
                e = eg[eg["world"] == wi]
                got = np.stack([e["body_a"], e["body_b"], e["kind"]], axis=1) if len(e) else np.zeros((0, 3), np.uint32)
                assert np.array_equal(got, o.events()), f"{what}: events differ"
    return sum(len(l) for l in live)


def run_queries(seed):
    """Rays and sphere casts against a random scene (a shipped map plus random bodies, some of which block lasers): random
    origins inside the scene's bounds, random directions (a share of them axis-aligned), random lengths and layer masks;
    ids, faces, fractions and cast normals must be the oracle's, bit for bit."""
    rng = np.random.default_rng(seed)
    name = ("stacked", "shapes", "test")[seed % 3]
    meshes = scenes.load_static(name)
    cap = 48
    g = gpx.World(worlds=1, max_bodies=cap)
    o = orc.World(cap)
    lo, hi = np.full(3, 1e9), np.full(3, -1e9)
    for pos, tris in meshes:
        g.add_mesh(pos, tris)
        o.add_mesh(pos, tris)
        w = np.asarray(tris, np.float64).reshape(-1, 3) + np.asarray(pos, np.float64)
        lo, hi = np.minimum(lo, w.min(axis=0)), np.maximum(hi, w.max(axis=0))
    g.commit()
    for _ in range(int(rng.integers(4, cap))):
        d = random_desc(rng, 1.0)
        d["position"] = tuple(float(v) for v in rng.uniform(lo, hi))
        d["ray_flags"] = int(rng.integers(0, 2))
        a, b = g.create(gpx.body_desc(**d)), o.create(orc.body_desc(**d))
        assert a == b
    for _ in range(int(rng.integers(0, 4))):     # let things move a little so that bodies are not where they were created
        assert g.step() == 0 and o.step() == 0
    rc = g.sync()
    if rc != 0:
        raise CapacityError(f"seed {seed} ({name}): a tick reported error code {rc} (bodies spawned inside each other)")
    m = 4000
    n = 20000
    rays = np.zeros(n, gpx.RAY_DTYPE)
    rays["origin"] = rng.uniform(lo - 1.0, hi + 1.0, (n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3))
    axis = rng.integers(0, 8, n) == 0
    d[axis] = np.eye(3)[rng.integers(0, 3, axis.sum())] * rng.choice([-1.0, 1.0], (axis.sum(), 1))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays["dir"] = d.astype(np.float32)
    rays["tmax"] = rng.choice([0.5, 3.0, 20.0, 200.0], n).astype(np.float32)
    rays["mask"] = rng.choice([1, 2, 3, 0xB, 0xF, 3 | (1 << 8), 0xF | (1 << 8)], n).astype(np.uint32)
    casts = np.zeros(m, gpx.CAST_DTYPE)
    for f in ("origin", "dir", "tmax", "mask"):
        casts[f] = rays[f][:m]
    casts["mask"] &= 0xF
    casts["radius"] = rng.choice([0.0, 0.05, 0.25, 0.6], m).astype(np.float32)
    if seed & 1:
        cg = g.spherecast(casts)      # straight after the (asynchronous) ticks: the batch has to wait for them itself
    hg, ho = g.raycast(rays), o.raycast(rays, mt=True)
    for f in ("body", "face"):
        assert np.array_equal(hg[f], ho[f]), f"seed {seed} ({name}): ray {f} differs at {np.nonzero(hg[f] != ho[f])[0][:5]}"
    assert np.array_equal(hg["fraction"].view(np.uint32), ho["fraction"].view(np.uint32)), f"seed {seed} ({name}): ray fractions differ"
    if not seed & 1:
        cg = g.spherecast(casts)
    co = o.spherecast(casts)
    xo, vo = o.state(cap)
    rc = g.sync()
    if rc != 0:
        raise CapacityError(f"seed {seed} ({name}): a tick reported error code {rc} (bodies spawned inside each other)")
    assert np.array_equal(g.transforms()[0].view(np.uint32), xo.view(np.uint32)), f"seed {seed} ({name}): body state differs"

    for f in ("body", "face"):
        assert np.array_equal(cg[f], co[f]), f"seed {seed} ({name}): cast {f} differs at {np.nonzero(cg[f] != co[f])[0][:5]}"
    assert np.array_equal(cg["fraction"].view(np.uint32), co["fraction"].view(np.uint32)), f"seed {seed} ({name}): cast fractions differ"
    assert np.array_equal(cg["normal"].view(np.uint32), co["normal"].view(np.uint32)), f"seed {seed} ({name}): cast normals differ"
    return name, int((hg["body"] != 0xFFFFFFFF).sum()), int((cg["body"] != 0xFFFFFFFF).sum())


def run_character(seed):
    """The player capsule on a shipped map among random bodies: a random walk (speed changes, jumps, the engine's
    ExtendedUpdate settings or the plain update), the physics tick after every move as in the engine; position, velocity,
    ground state, contact list and the bodies it pushes must be the oracle's after every tick."""
    rng = np.random.default_rng(seed)
    name = ("stacked", "test")[seed % 2]
    meshes = scenes.load_static(name)
    cap = 24
    g = gpx.World(worlds=1, max_bodies=cap)
    o = orc.World(cap)
    for pos, tris in meshes:
        g.add_mesh(pos, tris)
        o.add_mesh(pos, tris)
    g.commit()
    start = (0.0, 0.5, -1.5) if name == "stacked" else (0.0, 1.0, 0.0)
    for _ in range(int(rng.integers(2, cap))):
        d = random_desc(rng, 1.5)
        d["position"] = (start[0] + float(rng.uniform(-2, 2)), start[1] + float(rng.uniform(-0.5, 1.5)), start[2] + float(rng.uniform(-2, 2)))
        a, b = g.create(gpx.body_desc(**d)), o.create(orc.body_desc(**d))
        assert a == b
    settings = (0.25, 0.25, 0.02, 0.15, float(np.cos(np.radians(75.0)))) if seed % 3 else None
    for side in (g, o):
        side.character_create(start)
    vx = vz = 0.0
    for tick in range(1, ticks + 1):
        if rng.integers(0, 12) == 0:
            speed = float(rng.choice([0.0, 1.5, 4.0, 9.0]))
            ang = float(rng.uniform(0, 2 * np.pi))
            vx, vz = speed * float(np.cos(ang)), speed * float(np.sin(ang))
        jump = rng.integers(0, 40) == 0
        for side in (g, o):
            p, v, ground, _ = side.character_get()
            vy = 0.0
            if ground != 0:
                vy = float(np.float32(v[1]) + np.float32(-9.81 / 60.0))
            elif jump:
                vy = 4.5
            side.character_set_velocity((vx, vy, vz))
            side.character_update(settings=settings)
        assert g.step() == 0 and o.step() == 0
        pg, vg, gg, bg = g.character_get()
        po, vo, go, bo = o.character_get()
        what = f"seed {seed} ({name}) tick {tick}"
        assert np.array_equal(pg.view(np.uint32), po.view(np.uint32)), f"{what}: character position {pg} vs {po}"
        assert np.array_equal(vg.view(np.uint32), vo.view(np.uint32)), f"{what}: character velocity {vg} vs {vo}"
        assert (gg, bg) == (go, bo), f"{what}: ground {gg}/{bg:#x} vs {go}/{bo:#x}"
        assert list(g.character_contacts()) == list(o.character_contacts()), f"{what}: contact lists differ"
        assert g.sync() == 0
        xo, vo_ = o.state(cap)
        assert np.array_equal(g.transforms()[0].view(np.uint32), xo.view(np.uint32)), f"{what}: bodies differ"
        assert np.array_equal(g.velocities()[0].view(np.uint32), vo_.view(np.uint32)), f"{what}: body velocities differ"
    return name


if __name__ == "__main__":
    first = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    seeds = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    ticks = int(sys.argv[3]) if len(sys.argv) > 3 else 240
    failed = 0
    for seed in range(first, first + seeds):
        if os.environ.get("FUZZ_MODE") in ("queries", "character"):
            try:
                r = run_queries(seed) if os.environ["FUZZ_MODE"] == "queries" else run_character(seed)
                print(f"seed {seed}: {os.environ['FUZZ_MODE']} {r}: identical", flush=True)
            except CapacityError as e:
                print("SKIPPED", str(e)[:200], flush=True)
            except AssertionError as e:
                failed += 1
                print("FAILED", str(e)[:400], flush=True)
            continue
        if os.environ.get("FUZZ_WORLDS"):
            try:
                n = run_ensemble(seed, int(os.environ["FUZZ_WORLDS"]), int(os.environ.get("FUZZ_CAP", 8)))
                print(f"seed {seed}: {os.environ['FUZZ_WORLDS']} worlds, {n} bodies at the end, {ticks} ticks: identical", flush=True)
            except AssertionError as e:
                failed += 1
                print("FAILED", str(e)[:400], flush=True)
            continue
        try:
            wide, cap, n, errors = run(seed)
        except CapacityError as e:
            print("SKIPPED", str(e)[:200], flush=True)
            continue
        except AssertionError as e:
            failed += 1
            print("FAILED", str(e)[:400], flush=True)
            continue
        print(f"seed {seed}: {'wide' if wide else 'ensemble'} cap {cap}, {n} bodies at the end, {ticks} ticks, "
              f"{errors} ticks with a reported error code (same on both sides): identical", flush=True)
    print(f"{seeds - failed} of {seeds} seeds identical")
    sys.exit(1 if failed else 0)
