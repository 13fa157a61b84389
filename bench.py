#!/usr/bin/env python
"""bench.py — throughput of the physics tick (and the batched ray query) on B200, next to the CPU restatement.

Workload (BASELINE.json configs[4], the one `metric` is quoted on): an ensemble of independent worlds, each the
8-box column over sector 0 of stacked.gmap with Philox-randomised initial velocities.  One STEP is
one fixed tick (dt = 1/60 s, two collision sub-steps; engine/src/physics/MapPhysics.c:72,105-108) of every world a
rank owns.  Worlds are independent, so ranks shard the ensemble with no data-path collective.  The default split is
the one BASELINE configs[4] / SURVEY §8d C5 name: 4096 worlds IN TOTAL, rank g owning worlds [g*4096/G, (g+1)*4096/G)
("strong"); `--scaling weak` gives every GPU its own 4096 worlds, and the strong run reports that figure as well
(`extra.weak`).  NCCL is used only for the timing reduction and the end-of-run stats gather.

  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm (CUDA through the C ABI, libgpx.so)
  python bench.py --impl reference ...                            the CPU arm: the oracle restatement on all host
                                                                  threads (Jolt/joltc is not buildable here)

JSON keys: see the contract in the task statement; `value` = body-steps/s with state resident in HBM, `e2e` = the
same metric through the engine-facing per-tick sequence with HOST buffers (ray batch H2D -> rays -> hits D2H -> step
-> transform mirror D2H), `rays` = the secondary metric (2^20 hitscan rays against shapes.gmap).
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import shutil
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

BYTES_PER_BODY_STEP = 136   # state R+W 104 B + read-only properties 32 B
BYTES_PER_RAY = 48          # ray in 32 B + hit out 16 B
BOXES = 8
RAYS_PER_WORLD_TICK = 5     # crosshair ray + 4 lasers, as in test.gmap (PlayerPhysics.c:305, Laser.c:142)
L2_FLUSH_BYTES = 256 << 20


def measured_traffic(kernel):
    """dram bytes per launch of `kernel` from the committed ncu capture (profiles/traffic.json), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            v = json.load(f).get(kernel)
        return float(v) if v is not None else None
    except Exception:
        return None


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device = device
        self.proc = None
        self.lines: list[str] = []
        self.th = None

    def start(self):
        exe = shutil.which("nvidia-smi")
        if not exe:
            return
        try:
            self.proc = subprocess.Popen([exe, "-i", str(self.device), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        self.th = threading.Thread(target=self._pump, daemon=True)
        self.th.start()

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if f[5 + k].lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


# --------------------------------------------------------------------------------------------- scene set-up

def make_gpu_ensemble(gpx, scenes, worlds: int, first_world: int, device: int):
    g = gpx.World(worlds=worlds, max_bodies=BOXES, device=device)
    for pos, tris in scenes.load_static("stacked"):
        g.add_mesh(pos, tris)
    g.commit()
    vel = scenes.ensemble_velocities(worlds, BOXES, first_world=first_world)
    descs = [gpx.body_desc(position=tuple(p)) for p in scenes.stack_positions(BOXES)]
    g.create_all(descs, linvel=vel)
    return g


def make_cpu_ensemble(orc, scenes, worlds: int, first_world: int = 0):
    import ctypes as C
    meshes = scenes.load_static("stacked")
    vel = scenes.ensemble_velocities(worlds, BOXES, first_world=first_world)
    pos = scenes.stack_positions(BOXES)
    ws = []
    for wi in range(worlds):
        o = orc.World(BOXES)
        for p, t in meshes:
            o.add_mesh(p, t)
        for k in range(BOXES):
            o.create(orc.body_desc(position=tuple(pos[k]), linear_velocity=tuple(vel[wi, k])))
        ws.append(o)
    arr = (C.c_void_p * worlds)(*[o.h for o in ws])
    return ws, arr


def tick_rays(gpx, worlds: int, tick: int, out: np.ndarray):
    """Per-tick engine-style rays for every world: 5 rays from around the column, STATIC|DYNAMIC layers."""
    n = worlds * RAYS_PER_WORLD_TICK
    rng = np.random.default_rng(1000 + tick)
    out["origin"][:n] = np.array([0.0, -0.5, -1.5], np.float32) + rng.uniform(-1.5, 1.5, (n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    out["dir"][:n] = d / np.linalg.norm(d, axis=1, keepdims=True)
    out["tmax"][:n] = 50.0
    out["mask"][:n] = gpx.RAYMASK_STATIC_DYNAMIC | (np.repeat(np.arange(worlds, dtype=np.uint32), RAYS_PER_WORLD_TICK) << 16)
    return n


# --------------------------------------------------------------------------------------------- CPU arm

def cpu_sample(orc, scenes, worlds: int, warm_ticks: int, ticks: int):
    """The CPU arm's one code path: `worlds` sample worlds of the ensemble (world indices 0 .. worlds-1) stepped by ONE
    orc_step_many call (pthreads over worlds, all host threads) for `ticks` ticks after `warm_ticks` untimed ones.
    Returns body-steps/s, threads, seconds."""
    ws, arr = make_cpu_ensemble(orc, scenes, worlds)
    L = orc.lib()
    threads = min(L.orc_max_threads(), worlds)
    if warm_ticks:
        assert L.orc_step_many(arr, worlds, 1.0 / 60.0, 2, warm_ticks) == 0
    t0 = time.perf_counter()
    err = L.orc_step_many(arr, worlds, 1.0 / 60.0, 2, ticks)
    dt = time.perf_counter() - t0
    assert err == 0
    return worlds * BOXES * ticks / dt, threads, dt


def cpu_tick_rate(orc, scenes, target_s: float, worlds: int, warm_ticks: int):
    """cpu_baseline leg of our arm: the same sample, its tick count scaled to ~target_s of CPU work."""
    _, _, dt = cpu_sample(orc, scenes, worlds, warm_ticks, 10)
    ticks = int(max(10, min(600, target_s / max(dt / 10, 1e-6))))
    v, threads, _ = cpu_sample(orc, scenes, worlds, warm_ticks, ticks)
    return v, threads, ticks


def run_reference(args):
    """--impl reference: the CPU restatement of the same tick on all host threads, same metric/config.  Each of the K
    steps is one tick of a bounded sample of the workload (--ref-worlds of its worlds)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import orc
    scenes = importlib.import_module("c-game-engine_b200.scenes")
    worlds = args.ref_worlds
    value, threads, dt = cpu_sample(orc, scenes, worlds, args.warmup, args.steps)
    total = args.worlds if args.scaling == "strong" else args.worlds * args.gpus
    sample = (f"{worlds} of the {total} worlds (world indices 0..{worlds - 1}), {args.steps} ticks after {args.warmup} "
              f"warm-up ticks, one orc_step_many call on {threads} host threads")
    line = {
        "impl": "reference", "metric": "body_steps_per_s", "value": value, "unit": "body-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": value, "unit": "body-steps/s", "cores": threads, "kind": "port", "sample": sample,
                         "note": "CPU restatement (oracle/orc.c): brute-force triangle loop, all-pairs broadphase (sweep above 256 bodies); "
                                 "Jolt/joltc is not vendored and cannot be built here"},
        "e2e": {"value": value, "unit": "body-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    return line


def worlds_of_rank(args, rank: int, world_size: int):
    """(first world index, worlds) of a rank: strong = BASELINE configs[4]'s split of args.worlds in total."""
    if args.scaling == "strong":
        first = args.worlds * rank // world_size
        return first, args.worlds * (rank + 1) // world_size - first
    return rank * args.worlds, args.worlds


def workload_config(args):
    total = args.worlds if args.scaling == "strong" else args.worlds * args.gpus
    return {"workload": f"C5 ensemble: {total} independent stacked.gmap worlds x {BOXES}-box column over {args.gpus} GPU(s) "
                        f"({args.scaling} scaling), Philox initial velocities (key 0x5EED0005), dt=1/60, 2 sub-steps, "
                        "10 velocity + 2 position iterations, bodies may not sleep",
            "worlds_total": total, "worlds_per_gpu": total // max(args.gpus, 1), "bodies_per_world": BOXES, "static_triangles": 396,
            "l2": "flushed between timed steps (256 MiB memset outside the event brackets)",
            "parallelism": f"worlds sharded, {args.gpus} rank(s), no data-path collective"}


def timed_ticks(g, steps, flush, torch):
    """`steps` ticks of the ensemble, CUDA events around each gpx_step on the library's stream, L2 flushed before each."""
    ms = 0.0
    for _ in range(steps):
        flush.zero_()
        torch.cuda.synchronize()
        g.timer_begin()
        rc = g.step()
        ms += g.timer_end()
        assert rc == 0, f"gpx_step error {rc}"
    return ms


# --------------------------------------------------------------------------------------------- GPU arm

def run_gpu(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world_size > 1:
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    gpx = importlib.import_module("c-game-engine_b200")
    scenes = importlib.import_module("c-game-engine_b200.scenes")
    L = gpx.lib()
    hbm_peak, peak_src = peaks()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    first, W = worlds_of_rank(args, rank, world_size)
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device="cuda")

    # ---------------- value: ticks with state resident in HBM
    g = make_gpu_ensemble(gpx, scenes, W, first, local_rank)
    for _ in range(args.warmup):
        assert g.step() == 0
    assert g.sync() == 0
    barrier()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    launches0 = L.gpx_launch_count()
    wall0 = time.perf_counter()
    dev_ms = timed_ticks(g, args.steps, flush, torch)
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - wall0)
    launches = L.gpx_launch_count() - launches0
    clk = clocks.stop() if rank == 0 else None
    assert g.sync() == 0
    dev_ms = max_over_ranks(dev_ms)
    ms_per_step = dev_ms / args.steps
    bodies_per_rank = W * BOXES
    total_worlds = args.worlds if args.scaling == "strong" else args.worlds * world_size
    value = total_worlds * BOXES / (ms_per_step * 1e-3)
    achieved = bodies_per_rank * BYTES_PER_BODY_STEP / (ms_per_step * 1e-3) / 1e9
    ticks_done = args.warmup + args.steps

    # ---------------- e2e: the engine-facing per-tick sequence with host buffers
    n_rays = W * RAYS_PER_WORLD_TICK
    h_rays = gpx.pinned_array(n_rays, gpx.RAY_DTYPE)
    h_hits = gpx.pinned_array(n_rays, gpx.HIT_DTYPE)
    tick_rays(gpx, W, 0, h_rays)
    for _ in range(3):
        g.raycast_into_async(h_rays, h_hits)
        g.step()
        g.sync()
    ticks_done += 3
    check_hits = np.array(h_hits[:2048], copy=True)
    g.raycast_into(h_rays, h_hits)
    # (the state moved on by one tick between the two batches, so only static hits are comparable)
    assert check_hits.shape == (min(2048, n_rays),)
    barrier()
    e2e_steps = args.steps
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        g.raycast_into_async(h_rays, h_hits)  # H2D rays, k_raycast, D2H hits (PlayerPhysics.c:305, Laser.c:142)
        rc = g.step()                         # JPH_PhysicsSystem_Update (MapPhysics.c:105)
        rc |= g.sync()                        # transform mirror D2H + the tick's one synchronisation (rows a9)
        assert rc == 0
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    ticks_done += e2e_steps
    e2e_value = total_worlds * BOXES * e2e_steps / e2e_s
    h2d = n_rays * 32
    d2h = n_rays * 16 + 2 * 16 * bodies_per_rank + 4
    stats = g.stats()
    assert (stats["error"] == 0).all()

    # the state the sample check below compares with the oracle (rank 0's first worlds, after `ticks_done` ticks)
    n_check = min(16, W)
    xf_check = g.transforms()[:n_check].copy() if rank == 0 else None

    only = args.headline_only  # profiling runs: the launch list then holds the headline step's kernels and nothing else
    # ---------------- the other split of the same ensemble (N > 1): 4096 worlds per GPU when the headline is the strong
    # split, so both curves come from one run
    weak_res = None
    if world_size > 1 and args.scaling == "strong" and not only:
        g2 = make_gpu_ensemble(gpx, scenes, args.worlds, rank * args.worlds, local_rank)
        for _ in range(args.warmup):
            assert g2.step() == 0
        assert g2.sync() == 0
        barrier()
        k = min(args.steps, 200)
        ms2 = max_over_ranks(timed_ticks(g2, k, flush, torch)) / k
        weak_res = {"scaling": "weak", "worlds_per_gpu": args.worlds, "steps": k, "ms_per_step": ms2,
                    "value": world_size * args.worlds * BOXES / (ms2 * 1e-3), "unit": "body-steps/s"}
        del g2

    # ---------------- secondary metric: 2^20 hitscan rays against shapes.gmap (C3)
    rays_res = None if only else bench_rays(gpx, scenes, args, local_rank, rank, world_size, barrier, max_over_ranks, flush, hbm_peak)

    # ---------------- C4: one wide world of 100k boxes (replicated per rank)
    wide_res = None if args.no_wide or only else bench_wide(gpx, scenes, args, local_rank, rank, world_size, barrier, max_over_ranks, flush, hbm_peak)

    # ---------------- end-of-run stats gather over NCCL (the only collective).  Everything after it runs on
    # rank 0 alone, with the process group gone, so the other GPUs are released instead of spinning in a barrier.
    gathered_worlds = W
    if dist is not None:
        t = torch.from_numpy(stats.view(np.uint8).reshape(-1).copy()).cuda()
        if args.scaling == "strong" and total_worlds % world_size:
            raise SystemExit("bench.py: --worlds must be a multiple of --gpus for the gather")
        out = [torch.empty_like(t) for _ in range(world_size)]
        dist.all_gather(out, t)
        gathered_worlds = sum(o.numel() for o in out) // 32
        dist.barrier()
        dist.destroy_process_group()

    line = None
    if rank == 0:
        # ---------------- latency lines (C2, C1) and the saturation sweep: one GPU
        single_res = bench_single_world(gpx, scenes, args, local_rank, rank) if not only else None
        test_map_res = bench_test_map(gpx, scenes, args, local_rank, rank) if not only else None
        sweep = bench_saturation(gpx, scenes, args, local_rank, flush, torch) if not only else None
        block = bench_block_variant(gpx, scenes, args, local_rank, flush, torch) if not only else None
        cpu = None
        sample_ok = None
        if not args.no_cpu:
            import orc
            v, threads, ticks = cpu_tick_rate(orc, scenes, args.cpu_seconds, args.ref_worlds, 5)
            cpu = {"value": v, "unit": "body-steps/s", "cores": threads, "kind": "port",
                   "sample": f"{args.ref_worlds} worlds of the same ensemble (world indices 0..{args.ref_worlds - 1}) x {ticks} ticks, one "
                             f"orc_step_many call on {threads} host threads (oracle/orc.c: brute-force triangle loop, all-pairs "
                             "broadphase; Jolt unavailable in this environment)"}
            # the timed ensemble against the oracle: the first worlds of rank 0, every tick this run made, bit for bit
            ws, arr = make_cpu_ensemble(orc, scenes, n_check, first_world=first)
            assert orc.lib().orc_step_many(arr, n_check, 1.0 / 60.0, 2, ticks_done) == 0
            xo = np.stack([o.state(BOXES)[0] for o in ws])
            sample_ok = bool(np.array_equal(xo.view(np.uint32), xf_check.view(np.uint32)))
            if rays_res is not None:
                cpu_rays_leg(orc, scenes, args, rays_res)
            if wide_res is not None:
                cpu_wide_leg(orc, scenes, args, wide_res)
        if rays_res is not None:
            rays_res.pop("_check", None)
        if wide_res is not None:
            wide_res.pop("_x30", None)
        issue = measured_traffic("k_tick_inst_executed")
        line = {
            "metric": "body_steps_per_s", "value": value, "unit": "body-steps/s", "n_gpus": world_size,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args),
            "clocks": clk,
            "e2e": {"value": e2e_value, "unit": "body-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "what": f"per tick: gpx_raycast_batch_async({n_rays} host rays, pinned) + gpx_step + gpx_sync_transforms; wall clock, max over ranks"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": "k_tick", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved / hbm_peak, "traffic": args.traffic if args.traffic is not None else measured_traffic("k_tick"), "peak_source": peak_src,
                         "bytes_per_unit": BYTES_PER_BODY_STEP, "units_per_launch": bodies_per_rank},
            # what actually bounds k_tick: warp instructions issued.  inst = smsp__inst_executed.sum per launch from the
            # committed ncu capture (profiles/traffic.json, 4096 worlds); peak = 148 SMs x 4 schedulers x 1 instruction / cycle
            "roofline_issue": None if issue is None or W != 4096 else {
                "bound": "fp32_issue", "kernel": "k_tick", "achieved": issue / (ms_per_step * 1e-3) / 1e9,
                "peak": 148 * 4 * (clk["sm_mhz"] or 1965.0) * 1e6 / 1e9 if clk else 148 * 4 * 1.965, "unit": "G warp-inst/s",
                "frac": issue / (ms_per_step * 1e-3) / (148 * 4 * ((clk["sm_mhz"] if clk and clk["sm_mhz"] else 1965.0) * 1e6)),
                "inst_per_launch": issue},
            "cpu_baseline": cpu,
            "sample_matches_oracle": sample_ok,
            "sample_check": f"transforms of worlds {first}..{first + n_check - 1} after all {ticks_done} ticks of this run vs independent oracle worlds, bit for bit",
            "rays": rays_res,
            "wide": wide_res,
            "single_world": single_res,
            "test_map": test_map_res,
            "extra": {"weak": weak_res, "saturation": sweep, "block_variant": block},
            "wall_ms_timed_region": wall_ms,
            "stats_gathered_worlds": gathered_worlds,
            "kinetic_energy_mean": float(stats["kinetic_energy"].mean()),
        }
    return line


def bench_block_variant(gpx, scenes, args, device, flush, torch):
    """SURVEY §8d C5's optional variant: 512 worlds x the 4 x 4 x 4 block of 64 boxes (about 150 manifolds per world; a block
    of 256 threads per world)."""
    W = 512
    pos = scenes.block_positions()
    g = gpx.World(worlds=W, max_bodies=64, max_manifolds=256, device=device)
    for p, t in scenes.load_static("stacked"):
        g.add_mesh(p, t)
    g.commit()
    g.create_all([gpx.body_desc(position=tuple(p)) for p in pos], linvel=scenes.ensemble_velocities(W, 64))
    for _ in range(30):
        assert g.step() == 0
    assert g.sync() == 0
    k = 60
    ms = timed_ticks(g, k, flush, torch) / k
    st = g.stats()
    assert (st["error"] == 0).all()
    return {"workload": "C5 block variant: 512 worlds x 64 boxes (4 x 4 x 4, pitch 0.45) on stacked.gmap, same kicks, ticks 30..89",
            "ms_per_tick": ms, "body_steps_per_s": W * 64 / (ms * 1e-3), "manifolds_per_world": float(st["manifolds"].mean())}


def bench_saturation(gpx, scenes, args, device, flush, torch):
    """One GPU, the same columns, 512 ... 32768 worlds: where the tick stops being one wave of latency-bound warps."""
    out = []
    for W in (512, 1024, 2048, 4096, 8192, 16384, 32768):
        g = make_gpu_ensemble(gpx, scenes, W, 0, device)
        for _ in range(20):
            assert g.step() == 0
        assert g.sync() == 0
        k = 60
        ms = timed_ticks(g, k, flush, torch) / k
        out.append({"worlds": W, "ms_per_tick": ms, "body_steps_per_s": W * BOXES / (ms * 1e-3)})
        del g
    return out


def make_wide_world(gpx, scenes, pos, device):
    n = len(pos)
    g = gpx.World(worlds=1, max_bodies=n, device=device)
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
    return g


def bench_single_world(gpx, scenes, args, device, rank):
    """BASELINE configs[1]: ONE stacked.gmap world, the 8-box column, 600 ticks.  A latency figure, not a throughput
    one — a single small world occupies one warp of one SM (the GPU path exists for thousands of them); reported so the
    configuration has a measured line next to the CPU restatement."""
    g = gpx.World(worlds=1, max_bodies=8, device=device)
    for pos, tris in scenes.load_static("stacked"):
        g.add_mesh(pos, tris)
    g.commit()
    for p in scenes.stack_positions(BOXES):
        g.create(gpx.body_desc(position=tuple(p)))
    for _ in range(10):
        assert g.step() == 0
    assert g.sync() == 0
    ticks = 600
    g.timer_begin()
    for _ in range(ticks):
        rc = g.step()
    ms = g.timer_end() / ticks
    assert rc == 0 and g.sync() == 0
    res = {"workload": "C2: one stacked.gmap world, 8-box column at rest height, 600 ticks (a single warp's latency)",
           "ms_per_tick": ms, "body_steps_per_s": BOXES / (ms * 1e-3)}
    if rank == 0 and not args.no_cpu:
        import orc
        o = orc.World(BOXES)
        for pos, tris in scenes.load_static("stacked"):
            o.add_mesh(pos, tris)
        for p in scenes.stack_positions(BOXES):
            o.create(orc.body_desc(position=tuple(p)))
        for _ in range(10):
            o.step()
        t0 = time.perf_counter()
        for _ in range(ticks):
            o.step()
        cpu_ms = 1e3 * (time.perf_counter() - t0) / ticks
        res["cpu_port"] = {"ms_per_tick": cpu_ms, "body_steps_per_s": BOXES / (cpu_ms * 1e-3), "cores": 1}
    return res


def bench_test_map(gpx, scenes, args, device, rank):
    """BASELINE configs[0]: the physics content of mapSources/test.json (test.gmap: 644 map triangles + 4 laser-emitter
    meshes, 13 bodies of which 2 dynamic), 600 fixed ticks; per tick the four lasers cast their rays, then the update and
    the transform readback — the engine-facing sequence with host buffers.  Latency again: one world."""
    sc = scenes.test_map_scene()
    g = gpx.World(worlds=1, max_bodies=16, device=device)
    for pos, rot, tris, fr in sc["meshes"]:
        g.add_mesh(pos, tris, friction=fr, rot=rot)
    g.commit()
    for d in sc["bodies"]:
        g.create(gpx.body_desc(**d))
    rays = scenes.laser_rays(sc["lasers"])
    h_rays = gpx.pinned_array(len(rays), gpx.RAY_DTYPE)
    h_hits = gpx.pinned_array(len(rays), gpx.HIT_DTYPE)
    h_rays[:] = rays
    for _ in range(10):
        g.raycast_into_async(h_rays, h_hits)
        assert (g.step() | g.sync()) == 0
    ticks = 600
    t0 = time.perf_counter()
    for _ in range(ticks):
        g.raycast_into_async(h_rays, h_hits)
        rc = g.step() | g.sync()
    ms = 1e3 * (time.perf_counter() - t0) / ticks
    assert rc == 0
    dyn = sum(1 for d in sc["bodies"] if d.get("motion_type", 2) == 2)
    res = {"workload": "C1: test.gmap, its 13 bodies (2 dynamic) and 4 lasers, 600 ticks; per tick rays + gpx_step + "
                       "gpx_sync_transforms with host buffers (wall clock)",
           "ms_per_tick": ms, "dynamic_bodies": dyn, "body_steps_per_s": dyn / (ms * 1e-3)}
    if rank == 0 and not args.no_cpu:
        import orc
        o = orc.World(16)
        for pos, rot, tris, fr in sc["meshes"]:
            o.add_mesh(pos, tris, friction=fr, rot=rot)
        for d in sc["bodies"]:
            o.create(orc.body_desc(**d))
        for _ in range(10):
            o.raycast(rays)
            o.step()
        t0 = time.perf_counter()
        for _ in range(ticks):
            o.raycast(rays)
            o.step()
        cpu_ms = 1e3 * (time.perf_counter() - t0) / ticks
        res["cpu_port"] = {"ms_per_tick": cpu_ms, "body_steps_per_s": dyn / (cpu_ms * 1e-3), "cores": 1}
    return res


def bench_wide(gpx, scenes, args, device, rank, world_size, barrier, max_over_ranks, flush, hbm_peak):
    """BASELINE configs[3]: mapSources/max_box.json scaled to 100k dynamic boxes in ONE world (broadphase + solver
    stress); every rank runs its own replica (a single world does not shard)."""
    import torch
    pos = scenes.lattice_positions()
    n = len(pos)
    g = make_wide_world(gpx, scenes, pos, device)
    for _ in range(30):           # let the lattice drop the 5-10 cm onto the floor / each other: contacts everywhere
        assert g.step() == 0
    assert g.sync() == 0
    x30 = g.transforms()[0].copy() if rank == 0 else None   # checked against the CPU port's state after its 30 ticks
    L = gpx.lib()
    barrier()
    l0 = L.gpx_launch_count()
    ms = 0.0
    for _ in range(args.wide_steps):
        flush.zero_()
        torch.cuda.synchronize()
        g.timer_begin()
        rc = g.step()
        ms += g.timer_end()
        assert rc == 0
    launches = (L.gpx_launch_count() - l0) / args.wide_steps
    assert g.sync() == 0
    ms = max_over_ranks(ms) / args.wide_steps
    y = g.transforms()[0, :, 1]
    res = {"metric": "body_steps_per_s", "value": world_size * n / (ms * 1e-3), "unit": "body-steps/s", "bodies": n,
           "ms_per_tick": ms, "launches_per_tick": launches,
           "workload": "C4: 100 x 10 x 100 lattice of 0.4 m boxes (pitch 0.5, xz jitter +-0.02, Philox key 0x5EED0004) in the "
                       "12-triangle 1024 m room of mapSources/max_box.json, one world per GPU, after 30 settling ticks",
           "roofline": {"bound": "hbm", "kernel": "wide tick (all kw_* kernels)", "achieved": n * BYTES_PER_BODY_STEP / (ms * 1e-3) / 1e9,
                        "peak": hbm_peak, "unit": "GB/s", "frac": n * BYTES_PER_BODY_STEP / (ms * 1e-3) / 1e9 / hbm_peak, "traffic": measured_traffic("wide_tick")},
           "min_y": float(y.min())}
    if x30 is not None:
        res["_x30"] = x30
    return res


def cpu_wide_leg(orc, scenes, args, res):
    """CPU port on C4 itself: the 100 x 10 x 100 lattice, settled for the same 30 ticks, stepped by orc_step_mt (the
    oracle's tick over all host threads, bit-identical to its serial tick).  With fewer than four threads the sample
    shrinks to a 32 x 10 x 32 corner of the lattice so that the leg stays within its time budget."""
    threads = orc.lib().orc_max_threads()
    dims = (100, 10, 100) if threads >= 4 else (32, 10, 32)
    sp = scenes.lattice_positions(*dims)
    o = orc.World(len(sp))
    for p, t in scenes.box_map():
        o.add_mesh(p, t)
    desc = orc.body_desc()
    for q in sp:
        desc.position[0], desc.position[1], desc.position[2] = float(q[0]), float(q[1]), float(q[2])
        o.create(desc)
    for _ in range(30):
        o.step_mt()
    x30 = res.pop("_x30", None)
    if x30 is not None and len(x30) == len(sp):
        res["sample_matches_oracle"] = bool(np.array_equal(o.state(len(sp))[0].view(np.uint32), x30.view(np.uint32)))
        res["sample_check"] = "transforms of all 100 000 boxes after the 30 settling ticks vs the CPU port's, bit for bit"
    t0 = time.perf_counter()
    k = 0
    while (time.perf_counter() - t0 < args.cpu_seconds / 2 and k < 60) or k < 3:
        o.step_mt()
        k += 1
    dt = time.perf_counter() - t0
    res["cpu_baseline"] = {"value": len(sp) * k / dt, "unit": "body-steps/s", "cores": threads, "kind": "port",
                           "sample": f"{dims[0]} x {dims[1]} x {dims[2]} = {len(sp)} boxes of the same lattice after the same 30 "
                                     f"settling ticks, {k} ticks, orc_step_mt on {threads} host threads (oracle/orc.c wide "
                                     "mode: sort-and-sweep candidates, colours solved in parallel)"}


def bench_rays(gpx, scenes, args, device, rank, world_size, barrier, max_over_ranks, flush, hbm_peak):
    import torch
    meshes = scenes.load_static("shapes")
    g = gpx.World(worlds=1, max_bodies=8, device=device)
    for pos, tris in meshes:
        g.add_mesh(pos, tris)
    g.commit()
    n = args.rays
    rays = scenes.shapes_rays(n, np.array([p for p, _ in meshes]), first=rank * n)
    L = gpx.lib()
    d_rays = L.gpx_device_alloc(n * 32)
    d_hits = L.gpx_device_alloc(n * 16)
    assert d_rays and d_hits
    L.gpx_memcpy_h2d(d_rays, rays.ctypes.data, n * 32)
    for _ in range(3):
        g.raycast_device(d_rays, n, d_hits)
    g.device_sync()
    barrier()
    reps = args.ray_reps
    ms = 0.0
    for _ in range(reps):
        flush.zero_()
        torch.cuda.synchronize()
        g.timer_begin()
        g.raycast_device(d_rays, n, d_hits)
        ms += g.timer_end()
    ms = max_over_ranks(ms) / reps
    # e2e with pinned host buffers
    h_rays = gpx.pinned_array(n, gpx.RAY_DTYPE)
    h_hits = gpx.pinned_array(n, gpx.HIT_DTYPE)
    h_rays[:] = rays
    g.raycast_into(h_rays, h_hits)
    barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        g.raycast_into(h_rays, h_hits)
    e2e_s = max_over_ranks(time.perf_counter() - t0) / reps
    hit_frac = float((h_hits["body"] != gpx.INVALID_BODY).mean())
    L.gpx_device_free(d_rays)
    L.gpx_device_free(d_hits)
    res = {"metric": "rays_per_s", "value": world_size * n / (ms * 1e-3), "unit": "rays/s", "rays_per_gpu": n,
           "ms_per_batch": ms, "workload": "C3: 2^20 closest-hit rays vs shapes.gmap (512 tris), STATIC mask, Philox key 0x5EED0003",
           "e2e": {"value": world_size * n / e2e_s, "unit": "rays/s", "h2d_bytes_per_step": n * 32, "d2h_bytes_per_step": n * 16},
           "roofline": {"bound": "hbm", "kernel": "k_raycast", "achieved": n * BYTES_PER_RAY / (ms * 1e-3) / 1e9,
                        "peak": hbm_peak, "unit": "GB/s", "frac": n * BYTES_PER_RAY / (ms * 1e-3) / 1e9 / hbm_peak,
                        "traffic": args.ray_traffic if args.ray_traffic is not None else measured_traffic("k_raycast")},
           "hit_fraction": hit_frac}
    if rank == 0:
        res["_check"] = (rays, np.array(h_hits[:min(n, args.cpu_rays)], copy=True))
    return res


def cpu_rays_leg(orc, scenes, args, res):
    rays, h_hits = res.pop("_check")
    n = len(rays)
    if True:
        meshes = scenes.load_static("shapes")
        o = orc.World(8)
        for pos, tris in meshes:
            o.add_mesh(pos, tris)
        m = min(n, args.cpu_rays)
        t0 = time.perf_counter()
        ho = o.raycast(rays[:m], mt=True)
        dt = time.perf_counter() - t0
        res["cpu_baseline"] = {"value": m / dt, "unit": "rays/s", "cores": orc.lib().orc_max_threads(), "kind": "port",
                               "sample": f"first {m} rays of the batch, brute force over 512 triangles (oracle/orc.c)"}
        res["sample_matches_oracle"] = bool(np.array_equal(ho.view(np.uint8), np.asarray(h_hits[:m]).view(np.uint8)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=600)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--worlds", type=int, default=4096, help="worlds in total (--scaling strong) or per GPU (weak)")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="strong: BASELINE configs[4], --worlds in total split over the GPUs; weak: --worlds per GPU")
    ap.add_argument("--rays", type=int, default=1 << 20, help="rays per GPU for the secondary metric")
    ap.add_argument("--ray-reps", type=int, default=20)
    ap.add_argument("--ref-worlds", type=int, default=256, help="worlds per step in the CPU sample")
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--cpu-rays", type=int, default=1 << 17)
    ap.add_argument("--no-wide", action="store_true", help="skip the C4 wide-world section")
    ap.add_argument("--wide-steps", type=int, default=30)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline legs (profiling runs)")
    ap.add_argument("--headline-only", action="store_true", help="skip the secondary sections (rays, C1, C2, C4): profiling runs")
    ap.add_argument("--traffic", type=float, default=None, help="ncu dram bytes per k_tick launch (profiles/), if known")
    ap.add_argument("--ray-traffic", type=float, default=None)
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    # stdout carries the ONE JSON line and nothing else: libraries that print there (NCCL's version banner at the first
    # communicator) are sent to stderr by pointing fd 1 at fd 2 while the run lasts
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    try:
        line = run_reference(args) if args.impl == "reference" else run_gpu(args)
    finally:
        sys.stdout.flush()
        os.dup2(json_fd, 1)
        os.close(json_fd)
    if line is not None:
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
