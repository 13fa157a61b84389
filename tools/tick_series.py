#!/usr/bin/env python
"""Time series of the headline workload: ms per tick (device time, no L2 flush) and the manifold-count distribution
every `block` ticks — shows what the ensemble looks like while the 600-tick average is taken."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
gpx = importlib.import_module("c-game-engine_b200")
scenes = importlib.import_module("c-game-engine_b200.scenes")
W = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
TICKS = int(sys.argv[2]) if len(sys.argv) > 2 else 600
BLOCK = int(sys.argv[3]) if len(sys.argv) > 3 else 50
g = bench.make_gpu_ensemble(gpx, scenes, W, 0, 0)
FLUSH = os.environ.get("FLUSH")  # FLUSH=1: write 256 MiB between ticks, as bench.py does (cold L2 for every tick)
if FLUSH:
    import torch
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for t0 in range(0, TICKS, BLOCK):
    if FLUSH:
        ms = 0.0
        for _ in range(BLOCK):
            flush.zero_()
            torch.cuda.synchronize()
            g.timer_begin()
            g.step()
            ms += g.timer_end()
    else:
        g.timer_begin()
        for _ in range(BLOCK):
            g.step()
        ms = g.timer_end()
    st = g.stats()
    m = st["manifolds"]
    print(f"ticks {t0:4d}..{t0 + BLOCK - 1:4d}: {ms / BLOCK * 1e3:7.1f} us/tick  manifolds mean {m.mean():5.2f} max {m.max():3d} "
          f"worlds>8: {(m > 8).sum():4d}  >16: {(m > 16).sum():3d}  awake bodies {st['awake_bodies'].mean():.2f}  "
          f"max speed p50 {np.median(st['max_speed']):.3f} p99 {np.percentile(st['max_speed'], 99):.3f}")
