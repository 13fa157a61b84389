"""GPU parity under random scenes and random host edits (tests/fuzz_parity.py): boxes and spheres of random size, mass,
friction, restitution and spin, kinematic movers, sensors, restricted degrees of freedom, bodies that may sleep; create /
destroy / set velocity / set position between ticks.  Transforms, velocities, sleep flags and contact events are compared
with the oracle after EVERY tick.  The seeds cover every launch shape of the tick: a warp tile per world (up to 32
bodies), a block per world (33..64 bodies — where this test found a missing barrier between the warm-start match and the
colouring), the same worlds forced onto the warp tile with parked rows, and the wide-world kernels."""
import importlib.util
import os
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
