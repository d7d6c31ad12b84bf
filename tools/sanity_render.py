#!/usr/bin/env python3
"""Small frames through every kernel family (for compute-sanitizer runs)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "raytracer-utah_b200", "python")); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np
import rtu_b200 as R
import make_synthetic
make_synthetic.ensure(("spheres_1000",))
ctx = R.Context(0)
for scene, mode, spp in (("Teapot/scene2.xml", R.MODE_WHITTED, 2), ("Project5/scene.xml", R.MODE_WHITTED, 1), ("Project11/scene.xml", R.MODE_PATH, 2),
                         ("synthetic/spheres_1000.xml", R.MODE_WHITTED, 1), ("Project13/scene.xml", R.MODE_PHOTON_GATHER, 1)):
    hs = R.HostScene(os.path.join(R.SCENES, scene)); sc = R.Scene(ctx, hs.desc)
    if mode == R.MODE_PHOTON_GATHER:
        sc.photon_map_generate(map_size=20000, seed=1)
    p = R.default_params(width=96, height=64, spp=spp, pattern=R.PATTERN_REFERENCE if spp > 1 else R.PATTERN_CENTER, mode=mode, shade_bounces=5, gi_bounces=2)
    out = sc.render(p, want=("rgb", "z", "node_id"))
    st = sc.stats()
    print(scene, mode, st["trace_rays"], st["shadow_rays"], float(np.nanmean(out["rgb"])))
    sc.close(); hs.close()
print("done")
