#!/usr/bin/env python
"""The C5 "4 x 4 x 4 block" variant: W worlds x 64 boxes, through the ensemble kernel and through the
wide-world kernels (GPX_WORLD_WIDE).  Usage: python tools/block_ensemble.py [worlds] [ticks]"""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
gpx = importlib.import_module("c-game-engine_b200")
scenes = importlib.import_module("c-game-engine_b200.scenes")
W = int(sys.argv[1]) if len(sys.argv) > 1 else 512
T = int(sys.argv[2]) if len(sys.argv) > 2 else 60
pos = scenes.block_positions()
for wide in (False, True):
    g = gpx.World(worlds=W, max_bodies=64, max_manifolds=256, wide=wide) if wide else gpx.World(worlds=W, max_bodies=64, max_manifolds=256)
    for p, t in scenes.load_static("stacked"):
        g.add_mesh(p, t)
    g.commit()
    vel = scenes.ensemble_velocities(W, 64)
    g.create_all([gpx.body_desc(position=tuple(p)) for p in pos], linvel=vel)
    for _ in range(30):
        assert g.step() == 0
    g.sync()
    g.timer_begin()
    for _ in range(T):
        g.step()
    ms = g.timer_end() / T
    st = g.stats()
    print(f"{'wide kernels ' if wide else 'ensemble kernel'}: {W} worlds x 64 boxes: {ms:.3f} ms/tick, {W * 64 / ms / 1e3:.1f} M body-steps/s, "
          f"manifolds mean {st['manifolds'].mean():.0f} max {st['manifolds'].max()}, errors {(st['error'] != 0).sum()}")
    if not wide and os.environ.get("PHASES"):
        g.phase_cycles(True)
        for _ in range(10):
            g.step()
        ph = g.phase_cycles(False)
        tot = sum(ph.values())
        for k, v in ph.items():
            print(f"  {k:14s} {v / (10 * W):10.0f} cyc/world-tick  {100 * v / tot:5.1f}%")
    del g
