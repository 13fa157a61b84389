import importlib, os, sys
sys.path.insert(0,'/root/repo')
import bench, torch
gpx = importlib.import_module("c-game-engine_b200"); scenes = importlib.import_module("c-game-engine_b200.scenes")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for W in (512, 1024, 2048, 4096):
    g = bench.make_gpu_ensemble(gpx, scenes, W, 0, 0)
    for _ in range(20): g.step()
    g.sync()
    k = 200
    ms = bench.timed_ticks(g, k, flush, torch) / k
    print(os.environ.get("GPX_TILE","auto"), W, f"{ms*1e3:.1f} us/tick  {W*8/ms/1e3:.1f} M body-steps/s")
    del g
