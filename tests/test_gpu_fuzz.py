"""GPU parity under random scenes and random host edits (tests/fuzz_parity.py): boxes and spheres of random size, mass,
friction, restitution and spin, kinematic movers, sensors, restricted degrees of freedom, bodies that may sleep; create /
destroy / set velocity / set position between ticks.  Transforms, velocities, sleep flags and contact events are compared
with the oracle after EVERY tick.  The seeds cover every launch shape of the tick: a warp tile per world (up to 32
bodies), a block per world (33..64 bodies — where this test found a missing barrier between the warm-start match and the
colouring), the same worlds forced onto the warp tile with parked rows, and the wide-world kernels."""
import importlib.util
import os
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _fuzz():
    spec = importlib.util.spec_from_file_location("fuzz_parity", os.path.join(ROOT, "tests", "fuzz_parity.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("cap,no_block,seeds", [(24, False, (200, 201, 202)), (48, False, (200, 216, 217)),
                                                (64, False, (238, 239)), (64, True, (204, 205)), (150, False, (300, 301))])
def test_random_scenes_with_host_edits_match_the_oracle_every_tick(monkeypatch, cap, no_block, seeds):
    monkeypatch.setenv("FUZZ_CAP", str(cap))
    monkeypatch.setenv("FUZZ_EVERY_TICK", "1")
    if no_block:
        monkeypatch.setenv("GPX_NO_BLOCK_TILE", "1")
    else:
        monkeypatch.delenv("GPX_NO_BLOCK_TILE", raising=False)
    fz = _fuzz()
    fz.ticks = 240
    for seed in seeds:
        fz.run(seed)


@pytest.mark.parametrize("worlds,cap,tile", [(48, 8, "8"), (48, 8, "16"), (32, 16, "16"), (24, 8, "32")])
def test_random_ensembles_match_one_oracle_world_each(monkeypatch, worlds, cap, tile):
    """The benchmarked launch shape with random content: narrow tiles (four or two worlds per warp), worlds that outgrow
    their lanes routed to the 32-lane launch, edits in random worlds; every world, every tick, events included."""
    monkeypatch.setenv("GPX_TILE", tile)
    monkeypatch.delenv("GPX_NO_BLOCK_TILE", raising=False)
    fz = _fuzz()
    fz.ticks = 160
    for seed in (500, 501):
        fz.run_ensemble(seed, worlds, cap)


def test_restitution_bias_is_taken_before_any_warm_start(monkeypatch):
    """Seed 6028, 64 worlds x 8 slots on 8-lane tiles: in world 27 a bouncy body's manifold read the velocity of a body
    another lane had already warm-started (the row set-up of all manifolds precedes the first impulse; a barrier was
    missing) — visible only when all four worlds of that warp were busy, at tick 107."""
    monkeypatch.setenv("GPX_TILE", "8")
    monkeypatch.delenv("GPX_NO_BLOCK_TILE", raising=False)
    fz = _fuzz()
    fz.ticks = 120
    fz.run_ensemble(6028, 64, 8)


def test_random_rays_and_sphere_casts_match_the_oracle():
    """20 000 random rays and 4000 sphere casts per scene (random origins, a share of axis-aligned directions, random
    lengths, layer masks and radii) against a shipped map with random bodies: ids, faces, fractions, normals."""
    fz = _fuzz()
    for seed in (0, 1, 2):
        fz.run_queries(seed)


def test_random_character_walks_match_the_oracle(monkeypatch):
    """The player capsule among random bodies on two shipped maps: random walk with speed changes and jumps, the tick after
    every move; the character, its contact list and the bodies it pushes, every tick."""
    monkeypatch.delenv("GPX_TILE", raising=False)
    fz = _fuzz()
    fz.ticks = 240
    for seed in (1, 2, 4):
        fz.run_character(seed)


def test_gmap_loaders_survive_corrupted_input(gpx, scenes):
    """Mutation fuzzing of the .gmap readers (decoded body and gzip container): flipped bytes, overwritten counts,
    truncations, insertions.  A map is loaded or refused — and a world that refused a map still loads the intact one and
    answers rays."""
    import ctypes as C
    import gasset
    rng = np.random.default_rng(5)
    blob = open(f"{scenes.GOLDEN}/stacked_min.gmap", "rb").read()
    _, _, body = gasset.read_container(blob)
    loaded = refused = 0
    for data, container in ((body, False), (blob, True)):
        for _ in range(120):
            b = bytearray(data)
            kind = rng.integers(0, 4)
            if kind == 0:
                for _ in range(int(rng.integers(1, 6))):
                    b[int(rng.integers(0, len(b)))] = int(rng.integers(0, 256))
            elif kind == 1:
                at = int(rng.integers(0, max(1, len(b) - 8)))
                b[at:at + 8] = (0, 1, 0xFFFFFFFF, 1 << 40, (1 << 64) - 1)[int(rng.integers(0, 5))].to_bytes(8, "little")
            elif kind == 2:
                b = b[:int(rng.integers(0, len(b)))]
            else:
                at = int(rng.integers(0, len(b)))
                b[at:at] = bytes(int(rng.integers(1, 64)))
            g = gpx.World(worlds=1, max_bodies=8)
            buf = (C.c_uint8 * max(1, len(b))).from_buffer_copy(bytes(b) or b"\0")
            fn = g.L.gpx_static_load_gmap_container if container else g.L.gpx_static_load_gmap
            rc = fn(g.h, C.addressof(buf), len(b))
            if rc >= 0:
                loaded += 1
                g.commit()
            else:
                refused += 1
                assert g.load_gmap(body) == 10        # the intact map still loads into the same world
                g.commit()
                rays = np.zeros(1, gpx.RAY_DTYPE)
                rays["origin"], rays["dir"], rays["tmax"], rays["mask"] = (0.0, 0.5, -1.5), (0.0, -1.0, 0.0), 10.0, 1
                assert g.raycast(rays)["body"][0] != 0xFFFFFFFF
            g.close()
    assert loaded + refused == 240 and refused > 60


def test_plate_wedged_between_floor_and_wall_stays_identical_and_finite(monkeypatch):
    """Fuzz seed 40275 (a wide world of 161 slots): the scene in which a box's floor and wall manifolds used to feed each
    other's warm start until the box left with a NaN orientation — bit-identical to the oracle through that tick, and the
    oracle's own test (tests/test_oracle.py) says it stays finite."""
    monkeypatch.delenv("FUZZ_CAP", raising=False)
    monkeypatch.delenv("GPX_TILE", raising=False)
    monkeypatch.setenv("FUZZ_EVERY_TICK", "1")
    fz = _fuzz()
    fz.ticks = 220
    fz.run(40275)
