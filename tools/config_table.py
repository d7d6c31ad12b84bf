#!/usr/bin/env python3
"""Every BASELINE.json config (and the SURVEY 8d synthetic shapes) on one GPU: rays, ms per frame, Mrays/s, with the
CPU oracle (C restatement, all host threads) on a 1-spp frame of the same workload beside it.  Prints a markdown table."""
import sys, os, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "raytracer-utah_b200", "python")); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np
import rtu_b200 as R
import make_synthetic
from oracle import oracle_py as O

make_synthetic.ensure(("grid1M", "spheres_100", "spheres_1000", "spheres_10000"))
ctx = R.Context(0)
MODES = {"whitted": R.MODE_WHITTED, "path": R.MODE_PATH, "photon": R.MODE_PHOTON, "gather": R.MODE_PHOTON_GATHER}
ONLY = set(sys.argv[1:])
ROWS = [
    ("1", "Project1Example.xml", 800, 600, "whitted", [1]),
    ("2", "Project4.xml", 800, 600, "whitted", [1, 16, 64]),
    ("3", "Teapot/scene2.xml", 1920, 1080, "whitted", [1, 64, 1024]),
    ("3", "Teapot/scene.xml", 1920, 1080, "whitted", [64]),
    ("3", "synthetic/grid1M.xml", 3840, 2160, "whitted", [1, 16]),
    ("3", "synthetic/spheres_100.xml", 1920, 1080, "whitted", [16]),
    ("3", "synthetic/spheres_1000.xml", 1920, 1080, "whitted", [16]),
    ("3", "synthetic/spheres_10000.xml", 1920, 1080, "whitted", [4]),
    ("4", "Project10/scene.xml", 800, 600, "path", [64, 256]),
    ("4", "Project11/scene.xml", 800, 600, "path", [64, 256, 1024]),
    ("4", "Project11/scene_glossy_soft.xml", 800, 600, "path", [64]),
    ("4", "Project11/scene_86.xml", 800, 600, "path", [64]),
    ("4", "Project11/scene.xml", 1920, 1080, "path", [64]),
    ("5", "Project13/scene.xml", 800, 600, "photon", [1]),
    ("5", "Project13/scene.xml", 1920, 1080, "photon", [1]),
    ("5", "Project13/scene.xml", 800, 600, "gather", [16]),
]
print("| config | scene | size | mode | spp | rays / frame | ms / frame | Mrays/s | CPU oracle Mrays/s (threads) |")
print("|---|---|---|---|---|---|---|---|---|")
cache = {}
for cfg, scene, W, H, mode, spps in ROWS:
    if ONLY and cfg not in ONLY:
        continue
    key = scene
    if key not in cache:
        hs = R.HostScene(os.path.join(R.SCENES, scene))
        sc = R.Scene(ctx, hs.desc)
        cache[key] = (hs, sc)
    hs, sc = cache[key]
    if mode in ("photon", "gather") and not getattr(sc, "_has_map", False):
        sc.photon_map_generate(seed=1)       # (the first call also allocates the page-locked staging and starts the host threads)
        st = sc.photon_map_generate(seed=1)  # the same map again: steady state
        sc._has_map = True
        ph = sc.photon_map_get()
        bal = np.zeros(len(ph) + 1, R.PHOTON_DTYPE); bal[1:] = ph
        O.set_photon_map(bal, 1.0, 0.5)
        print("| 5 | %s | - | photon map: 10^6 photons | - | %d (emission) | %.1f emit + %.1f kd-tree | %.0f | - |" %
              (scene, st["trace_rays"], st["emit_ms"], st["build_ms"], st["trace_rays"] / st["emit_ms"] * 1e-3))
    # CPU oracle on one 1-spp frame (bounded: small sizes only take seconds)
    cpu = "-"
    try:
        t0 = time.perf_counter()
        po = R.default_params(width=W if W * H <= 2100000 else W // 2, height=H if W * H <= 2100000 else H // 2, spp=1, pattern=R.PATTERN_CENTER,
                              mode=MODES[mode], shade_bounces=5, gi_bounces=4)
        o = O.render(hs.desc, params=po, want=("rgb",))
        dt = time.perf_counter() - t0
        cpu = "%.1f (%d)" % ((o["stats"]["trace_rays"] + o["stats"]["shadow_rays"]) / dt * 1e-6, o["stats"]["threads"])
    except Exception as e:  # noqa
        cpu = "n/a"
    for spp in spps:
        p = R.default_params(width=W, height=H, spp=spp, pattern=R.PATTERN_REFERENCE if spp > 1 else R.PATTERN_CENTER, mode=MODES[mode],
                             shade_bounces=5, gi_bounces=4, flags=R.FLAG_TIME_KERNELS)
        best = None
        for it in range(3 if spp <= 256 else 2):
            sc.render_device(p); st = sc.stats()
            if best is None or st["device_ms"] < best["device_ms"]: best = st
        rays = best["trace_rays"] + best["shadow_rays"]
        print("| %s | %s | %dx%d | %s | %d | %d | %.3f | %.0f | %s |" % (cfg, scene, W, H, mode, spp, rays, best["device_ms"], rays / best["device_ms"] * 1e-3, cpu))
        sys.stdout.flush()
