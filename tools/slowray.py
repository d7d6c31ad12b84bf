import sys, os, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "raytracer-utah_b200", "python")); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np
import rtu_b200 as R, make_synthetic
make_synthetic.ensure(("grid1M",))
hs = R.HostScene(os.path.join(R.SCENES, "synthetic/grid1M.xml"))
ctx = R.Context(0); sc = R.Scene(ctx, hs.desc)
W, H = 1920, 1080
p = R.default_params(width=W, height=H, spp=1, pattern=R.PATTERN_CENTER, mode=R.MODE_WHITTED)
rays = sc.camera_rays(p).reshape(H, W)
def t(r):
    r = np.ascontiguousarray(r.reshape(-1))
    sc.trace(r)
    t0 = time.perf_counter(); sc.trace(r); return (time.perf_counter() - t0) * 1e3
print("all", t(rays))
bands = [(t(rays[y:y + 8]), y) for y in range(0, H, 8)]
bands.sort(reverse=True); print(bands[:5], bands[len(bands) // 2])
y = bands[0][1]
cols = [(t(rays[y:y + 8, x:x + 32]), x) for x in range(0, W, 32)]
cols.sort(reverse=True); print(cols[:5], cols[len(cols) // 2])
x = cols[0][1]
px = [(t(rays[yy:yy + 1, xx:xx + 1]), yy, xx) for yy in range(y, y + 8) for xx in range(x, x + 32)]
px.sort(reverse=True); print(px[:5], px[len(px) // 2])
_, yy, xx = px[0]
print("ray", rays[yy, xx], sc.trace(rays[yy:yy+1, xx:xx+1].reshape(-1)))
