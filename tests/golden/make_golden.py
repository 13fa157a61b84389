#!/usr/bin/env python
"""Generate tests/golden/* from the reference's shipped assets and from the CPU oracle.

Run in the build container (needs /root/reference); the outputs are committed because the GPU box has no
reference tree.  Three kinds of fixture:

  static_<map>.npz   collision meshes decoded from assets/game/map/<map>.gmap (MapLoader.c:200-273 layout)
  <map>_min.gmap     the same collision section re-serialised as a minimal asset container (no render data),
                     so the C loader can be tested against the real container + gzip + map layout
  models.npz         collision hull summaries of assets/game/model/*.gmdl (ModelLoader.c:145-211)
  oracle_*.npz       oracle outputs pinned as regression vectors (rays on shapes.gmap, the 8-box stack)
"""
from __future__ import annotations

import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
import gasset  # noqa: E402

REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")


def main():
    os.makedirs(OUT, exist_ok=True)
    if "--only-oracle" not in sys.argv:
        assets()
    oracle_vectors()


def assets():
    for name in ("test", "stacked", "shapes", "orb"):
        m = gasset.load_gmap(f"{REF}/assets/game/map/{name}.gmap")
        assert m.leftover == 0
        pos = np.array([cm.pos for cm in m.meshes], dtype=np.float32)
        tris = [np.concatenate(cm.subshapes) if len(cm.subshapes) else np.zeros((0, 3, 3), np.float32) for cm in m.meshes]
        start = np.cumsum([0] + [len(t) for t in tris]).astype(np.int64)
        actors = np.array([a.cls for a in m.actors])
        actor_xf = np.array([list(a.pos) + list(a.euler) for a in m.actors], dtype=np.float32).reshape(-1, 6)
        np.savez_compressed(os.path.join(OUT, f"static_{name}.npz"), mesh_pos=pos, mesh_start=start,
                            tris=np.concatenate(tris).astype(np.float32), actors=actors, actor_xf=actor_xf)
        body = gasset.build_gmap_body(m.meshes, actors=m.actors)
        with open(os.path.join(OUT, f"{name}_min.gmap"), "wb") as f:
            f.write(gasset.write_container(gasset.MAP_ASSET_TYPE, 1, body))
        print(name, len(pos), "meshes", int(start[-1]), "tris", len(m.actors), "actors")

    models = {}
    for name in ("cube", "orb", "leafy", "eraser_w", "laseremitter"):
        g = gasset.load_gmdl(f"{REF}/assets/game/model/{name}.gmdl")
        assert g.leftover == 0
        models[f"{name}_type"] = np.int32(g.collision_type)
        models[f"{name}_bb"] = np.array(list(g.bb_origin) + list(g.bb_extents), np.float32)
        if g.hulls:
            models[f"{name}_hull_counts"] = np.array([len(p) for _, p in g.hulls], np.int32)
            pts = np.concatenate([p + o for o, p in g.hulls])
            models[f"{name}_hull_aabb"] = np.concatenate([pts.min(0), pts.max(0)]).astype(np.float32)
            models[f"{name}_hull_maxr"] = np.float32(np.linalg.norm(pts, axis=1).max())
            models[f"{name}_hull_minr"] = np.float32(np.linalg.norm(pts, axis=1).min())
            # the points themselves (orb: every 16th of its 32 514) for the hull -> primitive classifier
            keep = pts if len(pts) <= 4096 else pts[::16]
            models[f"{name}_hull_points"] = keep.astype(np.float32)
        if g.tris is not None:
            models[f"{name}_tris"] = g.tris.astype(np.float32)
    np.savez_compressed(os.path.join(OUT, "models.npz"), **models)



def oracle_vectors():
    # ---- oracle regression vectors (re-run with --only-oracle after a deliberate change of the restated solver)
    import orc  # noqa: E402
    scenes = importlib.import_module("c-game-engine_b200.scenes")

    meshes = scenes.load_static("shapes")
    w = orc.World(8)
    for pos, tris in meshes:
        w.add_mesh(pos, tris)
    rays = scenes.shapes_rays(8192, np.array([p for p, _ in meshes]))
    hits = w.raycast(rays)
    np.savez_compressed(os.path.join(OUT, "oracle_rays_shapes.npz"), hits=hits)
    print("rays: hit fraction", float((hits["body"] != orc.INVALID).mean()))

    meshes = scenes.load_static("stacked")
    w = orc.World(8)
    for pos, tris in meshes:
        w.add_mesh(pos, tris)
    for p in scenes.stack_positions(8):
        w.create(orc.body_desc(position=tuple(p)))
    snaps = {}
    for tick in range(1, 601):
        assert w.step() == 0
        if tick in (1, 10, 60, 600):
            xf, vel = w.state(8)
            snaps[f"xf_{tick}"] = xf
            snaps[f"vel_{tick}"] = vel
    np.savez_compressed(os.path.join(OUT, "oracle_stack8.npz"), **snaps)
    print("stack8 y after 600 ticks:", snaps["xf_600"][:, 1])


if __name__ == "__main__":
    main()
