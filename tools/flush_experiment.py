#!/usr/bin/env python
"""How much of the flushed-L2 tick time is instruction / static-map refetch?  (diagnostic, not a bench number)"""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, bench
gpx = importlib.import_module("c-game-engine_b200")
scenes = importlib.import_module("c-game-engine_b200.scenes")
W = 4096
g = bench.make_gpu_ensemble(gpx, scenes, W, 0, 0)
tiny = bench.make_gpu_ensemble(gpx, scenes, 4, 0, 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(20):
    g.step(); tiny.step()
g.sync(); tiny.sync()
def run(mode, n=100):
    ms = 0.0
    for _ in range(n):
        if mode != "warm":
            flush.zero_(); torch.cuda.synchronize()
        if mode == "flush+code":
            tiny.step(); tiny.sync()
        g.timer_begin(); g.step(); ms += g.timer_end()
    return ms / n * 1e3
for mode in ("warm", "flush", "flush+code", "warm"):
    print(f"{mode:12s} {run(mode):8.1f} us/tick")
