#!/usr/bin/env python3
"""Device-time of whole frames, for iterating on kernels:  quickbench.py [scene.xml [W H [whitted|path|primary|photon|gather [spp,spp,..]]]]"""
import sys, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "raytracer-utah_b200", "python"))
import rtu_b200 as R
scene = sys.argv[1] if len(sys.argv) > 1 else "Teapot/scene2.xml"
W, H = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (1920, 1080)
mode = {"path": R.MODE_PATH, "primary": R.MODE_PRIMARY, "photon": R.MODE_PHOTON, "gather": R.MODE_PHOTON_GATHER}.get(sys.argv[4] if len(sys.argv) > 4 else "", R.MODE_WHITTED)
spps = [int(x) for x in sys.argv[5].split(",")] if len(sys.argv) > 5 else [1, 16, 64]
if scene.startswith("synthetic/"):
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import make_synthetic
    make_synthetic.ensure((os.path.basename(scene)[:-4],))
hs = R.HostScene(os.path.join(R.SCENES, scene))
ctx = R.Context(0); sc = R.Scene(ctx, hs.desc)
if mode in (R.MODE_PHOTON, R.MODE_PHOTON_GATHER): sc.photon_map_generate(seed=0)
for spp in spps:
    p = R.default_params(width=W, height=H, spp=spp, pattern=R.PATTERN_REFERENCE, mode=mode, flags=R.FLAG_TIME_KERNELS)
    best = None
    for it in range(3):
        sc.render_device(p); st = sc.stats()
        if best is None or st["device_ms"] < best["device_ms"]: best = st
    rays = best["trace_rays"] + best["shadow_rays"]
    ks = {k: round(v["ms"], 3) for k, v in best.items() if isinstance(v, dict)}
    print(json.dumps(dict(scene=scene, spp=spp, rays=rays, ms=round(best["device_ms"], 3), mrays=round(rays / best["device_ms"] * 1e-3, 1),
                          launches=best["kernel_launches"], box=best["box_tests"], tri=best["tri_tests"], kernels_ms=ks)))
