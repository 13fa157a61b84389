import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench, orc
gpx = importlib.import_module("c-game-engine_b200")
scenes = importlib.import_module("c-game-engine_b200.scenes")
W = 4096
g = bench.make_gpu_ensemble(gpx, scenes, W, 0, 0)
hist = []
for t in range(600):
    g.step()
    if t % 100 == 99:
        st = g.stats()
        hist.append((t + 1, int((st["manifolds"] > 8).sum()), int(st["manifolds"].max())))
print(hist)
st = g.stats()
bad = np.nonzero(st["manifolds"] > 8)[0]
print("worlds with > 8 manifolds:", bad[:20], "count", len(bad))
x = g.transforms()
for wi in bad[:3]:
    print("world", wi, "manifolds", st["manifolds"][wi])
    print(np.round(x[wi, :, :3], 3))
    print("quat", np.round(x[wi, :, 3:], 3))
