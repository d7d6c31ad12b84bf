#!/usr/bin/env python3
"""Adaptive sampling against fixed sample counts: samples spent, time and RMSE vs a converged frame."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "raytracer-utah_b200", "python"))
import rtu_b200 as R
ctx = R.Context(0)
for scene, W, H in (("Project11/scene.xml", 800, 600), ("Project10/scene.xml", 800, 600)):
    hs = R.HostScene(os.path.join(R.SCENES, scene)); sc = R.Scene(ctx, hs.desc)
    kw = dict(width=W, height=H, pattern=R.PATTERN_REFERENCE, mode=R.MODE_PATH, shade_bounces=5, gi_bounces=4, seed=1)
    ref = sc.render(R.default_params(spp=4096, seed=99, **{k: v for k, v in kw.items() if k != "seed"}), want=("rgb",))["rgb"].astype(np.float64)
    def report(name, p):
        out = sc.render(p, want=("rgb",))["rgb"].astype(np.float64); st = sc.stats()
        print(json.dumps(dict(scene=scene, run=name, mean_spp=round(st["pixel_samples"] / (W * H), 1), device_ms=round(st["device_ms"], 2),
                              rays=st["trace_rays"] + st["shadow_rays"], rmse=float(np.sqrt(np.mean((out - ref) ** 2))))))
    for spp in (64, 256, 1024):
        report("fixed %d spp" % spp, R.default_params(spp=spp, **kw))
    for tgt in (1e-4, 2.5e-5, 6e-6):
        report("adaptive, target variance %g (min 8, step 8, max 1024)" % tgt, R.default_params(spp=1024, adaptive_min_spp=8, adaptive_step=8, adaptive_target=tgt, **kw))
    sc.close(); hs.close()
