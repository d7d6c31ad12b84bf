#!/usr/bin/env python3
"""Every shipped scene x {whitted, head} at low resolution in a fresh context: no errors, ray counts equal to the oracle's
(whitted) and finite images."""
import sys, os, glob, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "raytracer-utah_b200", "python")); sys.path.insert(0, ROOT)
import numpy as np
import rtu_b200 as R
from oracle import oracle_py as O
scenes = sorted(glob.glob(os.path.join(R.SCENES, "*.xml")) + glob.glob(os.path.join(R.SCENES, "Project*/*.xml")) + glob.glob(os.path.join(R.SCENES, "Teapot/*.xml")))
bad = 0
for path in scenes:
    name = os.path.relpath(path, R.SCENES)
    try:
        hs = R.HostScene(path)
    except Exception as e:
        print(name, "LOAD FAILED", str(e)[:100]); bad += 1; continue
    if hs.desc.n_materials == 0:
        print(name, "skipped: no materials (the reference's Shade would dereference NULL)"); continue
    d = hs.desc
    stochastic = d.camera.dof > 0 or any(d.lights[i].size > 0 for i in range(d.n_lights)) or \
        any(d.materials[i].reflection_glossiness > 0 or d.materials[i].refraction_glossiness > 0 for i in range(d.n_materials))
    ctx = R.Context(0); sc = R.Scene(ctx, hs.desc)
    for mode, spp in ((R.MODE_WHITTED, 1), (R.MODE_PATH, 4)):
        p = R.default_params(width=160, height=120, spp=spp, pattern=R.PATTERN_CENTER if spp == 1 else R.PATTERN_REFERENCE, mode=mode, shade_bounces=5, gi_bounces=4)
        try:
            t0 = time.time()
            out = sc.render(p, want=("rgb",))["rgb"]
            st = sc.stats()
            msg = "rays %d+%d  %.1f ms  finite %.3f" % (st["trace_rays"], st["shadow_rays"], st["device_ms"], np.isfinite(out).mean())
            if mode == R.MODE_WHITTED:
                o = O.render(hs.desc, params=p, want=("rgb",))
                same = o["stats"]["trace_rays"] == st["trace_rays"] and o["stats"]["shadow_rays"] == st["shadow_rays"]
                fin = np.isfinite(o["rgb"]) & np.isfinite(out)
                err = np.abs(out - o["rgb"])[fin]
                tol = (1e-4 * np.maximum(np.abs(out), np.abs(o["rgb"])) + 1e-6)[fin]
                nbad = int((err > tol).sum())
                if stochastic:  # soft lights / glossy lobes / depth of field draw random numbers: only the means are comparable
                    msg += "  stochastic scene: mean %.4f vs oracle %.4f" % (out[fin].mean(), o["rgb"][fin].mean())
                else:
                    msg += "  counts==oracle %s  pixels out of tol %d  nan mask equal %s" % (same, nbad, np.array_equal(np.isfinite(o["rgb"]), np.isfinite(out)))
                    if not same or nbad: bad += 1
            print(name, "whitted" if mode == R.MODE_WHITTED else "head", msg)
        except Exception as e:
            print(name, mode, "FAILED", str(e)[:160]); bad += 1
    sc.close(); ctx.close(); hs.close()
print("problems:", bad)
