#!/usr/bin/env python
"""Per-phase SM-cycle split of the tick kernel on the bench workload (uses gpx_debug_phase_cycles)."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
gpx = importlib.import_module("c-game-engine_b200")
scenes = importlib.import_module("c-game-engine_b200.scenes")
W = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
g = bench.make_gpu_ensemble(gpx, scenes, W, 0, 0)
WARM = int(sys.argv[2]) if len(sys.argv) > 2 else 20
for _ in range(WARM):
    g.step()
g.phase_cycles(True)
N = 50
if os.environ.get("FLUSH"):  # FLUSH=1: cold L2 for every tick, as bench.py times them
    import torch
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ms = 0.0
    for _ in range(N):
        flush.zero_()
        torch.cuda.synchronize()
        g.timer_begin()
        g.step()
        ms += g.timer_end()
else:
    g.timer_begin()
    for _ in range(N):
        g.step()
    ms = g.timer_end()
ph = g.phase_cycles(False)
tot = sum(ph.values())
st = g.stats()
print("manifolds per world: mean", st["manifolds"].mean(), "max", st["manifolds"].max(), "errors", (st["error"] != 0).sum())
print(f"{W} worlds after {WARM} ticks: {ms / N * 1e3:.1f} us/tick (with counters armed); cycles per world-tick (lane 0): {tot / (N * W):.0f}")
for k, v in ph.items():
    print(f"  {k:14s} {v / (N * W):10.0f} cyc/world-tick  {100 * v / tot:5.1f}%")
