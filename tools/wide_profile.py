#!/usr/bin/env python
"""Timing of the wide-world tick (BASELINE config 4: 100 x 10 x 100 boxes in the big room) on one GPU."""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
gpx = importlib.import_module("c-game-engine_b200")
scenes = importlib.import_module("c-game-engine_b200.scenes")
nx, ny, nz = (int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (100, 10, 100)))
ticks = int(sys.argv[4]) if len(sys.argv) > 4 else 60


def make_world(pos):
    n = len(pos)
    g = gpx.World(worlds=1, max_bodies=n)
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


pos = scenes.lattice_positions(nx, ny, nz)
g = make_world(pos)
for _ in range(5):
    assert g.step() == 0
g.sync()
l0 = g.L.gpx_launch_count()
g.timer_begin()
for _ in range(ticks):
    g.step()
ms = g.timer_end() / ticks
launches = (g.L.gpx_launch_count() - l0) / ticks
assert g.sync() == 0
c = np.zeros(8, np.uint32)
g.L.gpx_debug_wide_counters(g.h, c.ctypes.data)
print("manifold slots", c[0], "small islands", c[1], "medium islands", c[4], "large-island manifolds", c[5], "colours", c[2], "error", c[3], "sub-steps that needed the radix passes", c[6])
print(f"{len(pos)} boxes: {ms:.3f} ms/tick, {len(pos) / ms * 1e3 / 1e6:.1f} M body-steps/s, {launches:.0f} launches/tick")
