"""Timing of the ray batch on one GPU, device-resident and host-to-host, with a parity check against the oracle (test
infrastructure: it lives under tests/ because it loads the oracle)."""
import importlib, os, sys, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
gpx=importlib.import_module('c-game-engine_b200'); scenes=importlib.import_module('c-game-engine_b200.scenes')
import orc
meshes=scenes.load_static('shapes')
g=gpx.World(worlds=1,max_bodies=8)
for p,t in meshes: g.add_mesh(p,t)
g.commit()
n=1<<20
rays=scenes.shapes_rays(n, np.array([p for p,_ in meshes]))
d_r=g.L.gpx_device_alloc(n*32); d_h=g.L.gpx_device_alloc(n*16)
g.L.gpx_memcpy_h2d(d_r, rays.ctypes.data, n*32)
for _ in range(3): g.raycast_device(d_r,n,d_h)
g.sync(); g.timer_begin()
for _ in range(20): g.raycast_device(d_r,n,d_h)
ms=g.timer_end()/20
hits=np.zeros(n,gpx.HIT_DTYPE); g.L.gpx_memcpy_d2h(hits.ctypes.data,d_h,n*16)
o=orc.World(8)
for p,t in meshes: o.add_mesh(p,t)
m=1<<15
ho=o.raycast(rays[:m], mt=True)
ok=np.array_equal(hits['body'][:m],ho['body']) and np.array_equal(hits['face'][:m],ho['face']) and np.array_equal(hits['fraction'][:m].view(np.uint32),ho['fraction'].view(np.uint32))
print(os.environ.get('GPX_RAY_LEAVES','default'), f"{ms:.4f} ms/batch {n/ms/1e6:.2f} G rays/s identical to oracle on {m}: {ok}")
# host-to-host (pinned): the chunked H2D | kernel | D2H pipeline of gpx_raycast_batch
import time
h_r = gpx.pinned_array(n, gpx.RAY_DTYPE); h_h = gpx.pinned_array(n, gpx.HIT_DTYPE)
h_r[:] = rays
g.raycast_into(h_r, h_h)
t0 = time.perf_counter()
for _ in range(10): g.raycast_into(h_r, h_h)
e2e = (time.perf_counter() - t0) / 10
same = np.array_equal(np.asarray(h_h).view(np.uint8), hits.view(np.uint8))
print(f"host-to-host {e2e*1e3:.3f} ms/batch {n/e2e/1e9:.2f} G rays/s, identical to the device-resident result: {same}")
