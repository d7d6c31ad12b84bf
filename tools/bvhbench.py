#!/usr/bin/env python3
"""Host hierarchy builds vs the device LBVH: load + upload times and frame times of the same scene."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "raytracer-utah_b200", "python")); sys.path.insert(0, os.path.join(ROOT, "tools"))
import rtu_b200 as R
import make_synthetic
make_synthetic.ensure(("grid1M",))
ctx = R.Context(0)
for scene, W, H, spp in (("synthetic/grid1M.xml", 3840, 2160, 4), ("Teapot/scene2.xml", 1920, 1080, 64)):
    for flags, name in ((0, "host cyBVH + SAH (reference-exact)"), (R.LOAD_DEVICE_BVH, "device LBVH")):
        t0 = time.perf_counter(); hs = R.HostScene(os.path.join(R.SCENES, scene), flags=flags); t_load = time.perf_counter() - t0
        t0 = time.perf_counter(); sc = R.Scene(ctx, hs.desc); ctx.synchronize(); t_up = time.perf_counter() - t0
        st0 = sc.stats()
        p = R.default_params(width=W, height=H, spp=spp, pattern=R.PATTERN_REFERENCE, mode=R.MODE_WHITTED, flags=R.FLAG_TIME_KERNELS)
        best = None
        for _ in range(3):
            sc.render_device(p); st = sc.stats()
            if best is None or st["device_ms"] < best["device_ms"]: best = st
        print(json.dumps(dict(scene=scene, hierarchy=name, load_s=round(t_load, 3), upload_s=round(t_up, 3), device_build_ms=round(st0["bvh_build_ms"], 3),
                              frame_ms=round(best["device_ms"], 3), kernels_ms={k: round(v["ms"], 3) for k, v in best.items() if isinstance(v, dict)})))
        sc.close(); hs.close()
