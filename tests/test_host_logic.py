"""CPU suite: harness-side logic — asset fixtures, the Philox workload generator, world sharding across ranks and
the N>1 stats gather (gloo, world_size 2)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ------------------------------------------------------------------------------------------------ fixtures

def test_static_fixtures_have_the_surveyed_counts(scenes):
    """Triangle counts per collision mesh as decoded from the shipped .gmap files (MapLoader.c:200-273)."""
    expect = {"test": [6, 22, 16, 508, 12, 10, 20, 8, 20, 12, 10],
              "stacked": [9, 12, 10, 12, 12, 41, 10, 46, 65, 179],
              "shapes": None, "orb": None}
    totals = {"test": 644, "stacked": 396, "shapes": 512}
    for name, per_mesh in expect.items():
        meshes = scenes.load_static(name)
        counts = [len(t) for _, t in meshes]
        if per_mesh:
            assert counts == per_mesh
        if name in totals:
            assert sum(counts) == totals[name]
    assert len(scenes.load_static("shapes")) == 24


def test_min_gmap_fixture_round_trips_through_the_container(scenes):
    import gasset
    for name in ("stacked", "test", "shapes", "orb"):
        blob = open(f"{scenes.GOLDEN}/{name}_min.gmap", "rb").read()
        typ, tver, body = gasset.read_container(blob)
        assert typ == gasset.MAP_ASSET_TYPE
        m = gasset.parse_gmap(body)
        assert m.leftover == 0
        ref = scenes.load_static(name)
        assert len(m.meshes) == len(ref)
        for cm, (pos, tris) in zip(m.meshes, ref):
            assert np.array_equal(np.asarray(cm.pos, np.float32), pos)
            got = np.concatenate(cm.subshapes) if len(cm.subshapes) else np.zeros((0, 3, 3), np.float32)
            assert np.array_equal(got, tris)
        # container integrity checks of AssetReader.c:150-257
        with pytest.raises(ValueError):
            gasset.read_container(b"XXXX" + blob[4:])
        with pytest.raises(ValueError):
            gasset.read_container(blob[:-3])


def test_model_fixture_facts(scenes):
    """cube.gmdl is a 0.4 m bevelled cube hull, orb.gmdl a radius-0.4 hull: the shapes the body store
    represents as BOX / SPHERE."""
    m = np.load(scenes.GOLDEN + "/models.npz")
    assert int(m["cube_type"]) == 2 and list(m["cube_hull_counts"]) == [120]
    assert np.allclose(m["cube_hull_aabb"], [-0.2] * 3 + [0.2] * 3, atol=1e-6)
    assert list(m["orb_hull_counts"]) == [32514]
    assert abs(float(m["orb_hull_maxr"]) - 0.4) < 2e-3 and abs(float(m["orb_hull_minr"]) - 0.4) < 2e-3
    assert m["laseremitter_tris"].shape[0] == 310


# ------------------------------------------------------------------------------------------------ workload generator

def test_philox_known_answers(scenes):
    """Random123 known-answer vectors for Philox4x32-10."""
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, out in kat:
        r = scenes.philox4x32(np.array([ctr], np.uint32), key[0], key[1])
        assert tuple(int(x) for x in r[0]) == out


def test_workloads_are_counter_based_and_shardable(scenes):
    """Any rank can regenerate its own slice: worlds [a,b) of the ensemble and rays [a,b) of the batch."""
    full = scenes.ensemble_velocities(64, 8)
    part = scenes.ensemble_velocities(16, 8, first_world=32)
    assert np.array_equal(full[32:48], part)
    assert full.min() >= -0.5 and full.max() < 0.5 and abs(full.mean()) < 0.02
    pos = np.array([p for p, _ in scenes.load_static("shapes")])
    r = scenes.shapes_rays(1000, pos)
    r2 = scenes.shapes_rays(300, pos, first=500)
    assert np.array_equal(r[500:800].view(np.uint8), r2.view(np.uint8))
    assert np.allclose(np.linalg.norm(r["dir"], axis=1), 1.0, atol=1e-6)
    assert abs(r["dir"].mean()) < 0.05
    assert (r["tmax"] == 50).all() and (r["mask"] == 1).all()


def test_scene_generators(scenes):
    p = scenes.stack_positions(8)
    assert np.allclose(np.diff(p[:, 1]), 0.401) and abs(p[0, 1] - (-1.5 + 0.25)) < 1e-6
    assert scenes.block_positions().shape == (64, 3)
    lat = scenes.lattice_positions(10, 4, 10)
    assert lat.shape == (400, 3) and abs(lat[:, 1].min() - (-511.75)) < 1e-6
    (pos, tris), = scenes.box_map()
    n = np.cross(tris[:, 1] - tris[:, 0], tris[:, 2] - tris[:, 0])
    assert tris.shape == (12, 3, 3) and (np.einsum("ij,ij->i", n, -tris.mean(axis=1)) > 0).all()   # normals face inward


# ------------------------------------------------------------------------------------------------ bench contract

def test_bench_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "5", "--warmup", "3",
                        "--ref-worlds", "16"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in line, k
    assert line["impl"] == "reference" and line["metric"] == "body_steps_per_s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["vs_baseline"] is None


def test_bench_gpu_arm_refuses_to_run_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "3"], capture_output=True, text=True,
                       timeout=300)
    assert r.returncode != 0 and "no CPU path" in (r.stderr + r.stdout)


# ------------------------------------------------------------------------------------------------ N > 1 (gloo)

_GLOO_WORKER = r"""
import os, sys, json
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
import importlib
dist.init_process_group("gloo")
rank, ws = dist.get_rank(), dist.get_world_size()
ens = importlib.import_module("c-game-engine_b200.ensemble")
import orc
scenes = importlib.import_module("c-game-engine_b200.scenes")
W = 6
first, count = ens.shard(W * ws, rank, ws)
assert (first, count) == (rank * W, W)
vel = scenes.ensemble_velocities(count, 8, first_world=first)
pos = scenes.stack_positions(8)
stats = np.zeros(count, ens.STATS_DTYPE)
for wi in range(count):
    o = orc.World(8)
    for p, t in scenes.load_static("stacked"):
        o.add_mesh(p, t)
    for k in range(8):
        o.create(orc.body_desc(position=tuple(pos[k]), linear_velocity=tuple(vel[wi, k])))
    for _ in range(5):
        assert o.step() == 0
    xf, v = o.state(8)
    stats[wi] = ens.host_stats(xf, v, mass=10.0, ticks=5)
allstats = ens.gather_stats(stats, dist)
t = ens.max_over_ranks(float(rank + 1), dist, device="cpu")
if rank == 0:
    print(json.dumps({"n": int(len(allstats)), "tmax": t, "ticks": int(allstats["ticks"].min()),
                      "distinct": int(len(set(allstats["position_checksum"].tolist()))),
                      "first_ke": float(allstats["kinetic_energy"][0]), "last_ke": float(allstats["kinetic_energy"][-1])}))
dist.destroy_process_group()
"""


def test_two_rank_sharding_and_stats_gather_over_gloo(tmp_path):
    """world_size 2 on CPU: contiguous world blocks per rank, no data-path exchange, one final gather."""
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29577", str(script), ROOT],
                       capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr[-3000:]
    out = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
    assert out["n"] == 12 and out["tmax"] == 2.0 and out["ticks"] == 5 and out["distinct"] == 12
    assert out["first_ke"] > 0 and out["last_ke"] > 0


# ------------------------------------------------------------------------------------------------ hull -> primitive

def test_hull_classifier_on_the_shipped_models(gpx, scenes):
    """gpx_shape_from_hull (host-side, no device): what CreateDynamicModelShape's convex hulls become in the body store
    (engine/src/assets/ModelLoader.c:324-341).  Hull points decoded from assets/game/model/*.gmdl by tests/golden/make_golden.py."""
    m = np.load(scenes.GOLDEN + "/models.npz")
    shape, he, c, exact = gpx.shape_from_hull(m["cube_hull_points"])
    assert shape == gpx.SHAPE_BOX and exact and np.allclose(he, 0.2, atol=1e-6) and np.allclose(c, 0, atol=1e-6)
    shape, he, c, exact = gpx.shape_from_hull(m["orb_hull_points"])
    assert shape == gpx.SHAPE_SPHERE and exact and abs(he[0] - 0.4) < 2e-3 and np.abs(c).max() < 2e-3
    shape, he, c, exact = gpx.shape_from_hull(m["leafy_hull_points"])
    assert shape == gpx.SHAPE_BOX and not exact                       # two tall hulls: only the bounding box is used
    assert np.allclose(he, (m["leafy_hull_aabb"][3:] - m["leafy_hull_aabb"][:3]) / 2, atol=1e-6)
    shape, he, c, exact = gpx.shape_from_hull(m["eraser_w_hull_points"])
    assert shape == gpx.SHAPE_BOX
    # synthetic: an exact box, a cylinder (NpcJohn.c:29 — neither a box nor a sphere), too few points
    box = np.array([[sx * 0.3, sy * 0.1, sz * 0.5] for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)], np.float32) + 1.0
    shape, he, c, exact = gpx.shape_from_hull(box)
    assert shape == gpx.SHAPE_BOX and exact and np.allclose(he, (0.3, 0.1, 0.5), atol=1e-6) and np.allclose(c, 1.0, atol=1e-6)
    ang = np.linspace(0, 2 * np.pi, 24, endpoint=False)
    cyl = np.array([[0.25 * np.cos(a), y, 0.25 * np.sin(a)] for a in ang for y in (-0.5, 0.5)], np.float32)
    shape, he, c, exact = gpx.shape_from_hull(cyl, tolerance=0.01)
    assert shape == gpx.SHAPE_BOX and not exact and np.allclose(he, (0.25, 0.5, 0.25), atol=1e-6)
    with pytest.raises(gpx.GpxError):
        gpx.shape_from_hull(cyl[:3])


def test_gmdl_collision_section_is_parsed_by_the_library(gpx, scenes):
    """gpx_model_load_gmdl(_container) (host-side, no device): the collision section of a .gmdl as ModelLoader.c:145-211
    reads it.  The files are rebuilt from the hulls / triangles decoded from the shipped models (tests/golden/make_golden.py),
    with render data that the parser has to skip (two materials, one skin, one LOD with vertices and indices)."""
    import struct
    import gasset
    m = np.load(scenes.GOLDEN + "/models.npz")

    def with_render_data(collision):
        # 2 materials, 2 slots, 1 skin, 1 lod; then what build_gmdl_body appends after its empty header
        b = struct.pack("<4IB", 2, 2, 1, 1, collision[16])
        for name in (b"texture/a\0", b"texture/bb\0"):
            b += struct.pack("<Q", len(name)) + name + struct.pack("<4fI", 1, 1, 1, 1, 0)
        b += struct.pack("<2I", 0, 1)                                 # the skin's material per slot
        b += struct.pack("<2fQ", 10.0, 100.0, 3) + bytes(3 * 48)      # lod distances, 3 vertices
        b += struct.pack("<I2I", 6, 3, 3) + bytes(4 * 3) + bytes(4 * 3)
        return b + collision[17:]

    cube = gasset.build_gmdl_body(2, m["cube_bb"][:3], m["cube_bb"][3:], hulls=[((0, 0, 0), m["cube_hull_points"])])
    mc = gpx.model_collision(with_render_data(cube))
    assert mc.collision_type == 2 and mc.n_hulls == 1 and mc.hull_points[0] == 120 and mc.exact == 1
    assert mc.hull[0].shape == gpx.SHAPE_BOX and np.allclose(list(mc.hull[0].half_extents), 0.2, atol=1e-6)
    assert np.allclose(list(mc.bb_extents), 0.2)
    # leafy: two hulls with offsets (none of them a primitive); the offsets land in `center`
    n0, n1 = (int(x) for x in m["leafy_hull_counts"])
    pts = m["leafy_hull_points"]
    leafy = gasset.build_gmdl_body(2, m["leafy_bb"][:3], m["leafy_bb"][3:],
                                   hulls=[((0, 0.5, 0), pts[:n0] - np.float32([0, 0.5, 0])), ((0, 0, 0), pts[n0:])])
    mc = gpx.model_collision(gasset.write_container(gasset.MODEL_ASSET_TYPE if hasattr(gasset, "MODEL_ASSET_TYPE") else 2, 1, with_render_data(leafy)),
                             container=True)
    assert mc.n_hulls == 2 and mc.exact == 0 and (mc.hull_points[0], mc.hull_points[1]) == (n0, n1)
    lo, hi = pts[:n0].min(0), pts[:n0].max(0)
    assert np.allclose(list(mc.hull[0].center), (lo + hi) / 2, atol=1e-5) and np.allclose(list(mc.hull[0].half_extents), (hi - lo) / 2, atol=1e-5)
    # a static model: its triangle count; hostile counts are refused, not allocated
    emitter = gasset.build_gmdl_body(1, m["laseremitter_bb"][:3], m["laseremitter_bb"][3:], tris=m["laseremitter_tris"])
    mc = gpx.model_collision(with_render_data(emitter))
    assert mc.collision_type == 1 and mc.n_triangles == 310
    bad = bytearray(emitter)
    bad[-310 * 36 - 8:-310 * 36] = struct.pack("<Q", 1 << 62)
    with pytest.raises(gpx.GpxError):
        gpx.model_collision(bytes(bad))
    with pytest.raises(gpx.GpxError):
        gpx.model_collision(cube[:40])


def test_gmdl_parser_survives_corrupted_input(gpx, scenes):
    """Mutation fuzzing of the host-side .gmdl reader (plain body and gzip container): flipped bytes, overwritten counts,
    truncations.  Every outcome is either a parsed model or GPX_ERR_INVALID_ARG — never a crash, a hang or an allocation
    sized by the file's say-so."""
    import gasset
    m = np.load(scenes.GOLDEN + "/models.npz")
    n0 = int(m["leafy_hull_counts"][0])
    pts = m["leafy_hull_points"]
    bodies = [gasset.build_gmdl_body(2, m["cube_bb"][:3], m["cube_bb"][3:], hulls=[((0, 0, 0), m["cube_hull_points"])]),
              gasset.build_gmdl_body(2, m["leafy_bb"][:3], m["leafy_bb"][3:], hulls=[((0, 0.5, 0), pts[:n0]), ((0, 0, 0), pts[n0:])]),
              gasset.build_gmdl_body(1, m["laseremitter_bb"][:3], m["laseremitter_bb"][3:], tris=m["laseremitter_tris"])]
    rng = np.random.default_rng(3)
    parsed = refused = 0
    for body in bodies:
        blobs = [(body, False), (gasset.write_container(2, 1, body), True)]
        for blob, container in blobs:
            for _ in range(150):
                b = bytearray(blob)
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
                try:
                    gpx.model_collision(bytes(b), container=container)
                    parsed += 1
                except gpx.GpxError:
                    refused += 1
    assert parsed + refused == 900 and refused > 300
