#!/usr/bin/env python3
"""Per-frame fixed cost: small frames rendered back to back with the scene resident.
Prints device time (CUDA events around the frame) and host wall time per frame for 1-spp frames."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "raytracer-utah_b200", "python"))
import rtu_b200 as R
CASES = [("Teapot/scene2.xml", 1920, 1080, R.MODE_WHITTED), ("Project1Example.xml", 800, 600, R.MODE_WHITTED), ("Project4.xml", 800, 600, R.MODE_WHITTED),
         ("Project11/scene.xml", 800, 600, R.MODE_PATH)]
ctx = R.Context(0)
for scene, W, H, mode in CASES:
    hs = R.HostScene(os.path.join(R.SCENES, scene))
    sc = R.Scene(ctx, hs.desc)
    for spp in (1, 4):
        p = R.default_params(width=W, height=H, spp=spp, pattern=R.PATTERN_REFERENCE if spp > 1 else R.PATTERN_CENTER, mode=mode)
        for _ in range(20):
            sc.render_device(p)
        n = 200
        t0 = time.perf_counter()
        dev = []
        for _ in range(n):
            sc.render_device(p)
            dev.append(sc.stats()["device_ms"]) if _ % 20 == 0 else None
        wall = (time.perf_counter() - t0) / n * 1e3
        st = sc.stats()
        rays = st["trace_rays"] + st["shadow_rays"]
        print(json.dumps(dict(scene=scene, size=[W, H], spp=spp, device_ms=round(min(dev), 4), wall_ms_per_frame=round(wall, 4), launches=st["kernel_launches"],
                              rays=rays, mrays_device=round(rays / min(dev) * 1e-3, 1))))
    # host-buffer frame (rtu_render: + resolve + D2H of RGB8), the e2e shape
    p = R.default_params(width=W, height=H, spp=1, mode=mode)
    for _ in range(5):
        sc.render(p, want=("rgb8",))
    t0 = time.perf_counter()
    for _ in range(50):
        sc.render(p, want=("rgb8",))
    print(json.dumps(dict(scene=scene, what="rtu_render rgb8 to pageable host memory, 1 spp", wall_ms_per_frame=round((time.perf_counter() - t0) / 50 * 1e3, 4))))
    sc.close(); hs.close()
