#!/usr/bin/env python3
"""GPU vs C oracle on one scene at low resolution: ids, z bit-exact; Whitted image within 1e-4; ray counts equal.
usage: check_scene.py scene.xml [W H]"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "raytracer-utah_b200", "python")); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np
import rtu_b200 as R
from oracle import oracle_py as O
scene = sys.argv[1]
W, H = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (240, 135)
if scene.startswith("synthetic/"):
    import make_synthetic
    make_synthetic.ensure((os.path.basename(scene)[:-4],))
hs = R.HostScene(os.path.join(R.SCENES, scene))
ctx = R.Context(0); sc = R.Scene(ctx, hs.desc)
g = sc.render(R.default_params(width=W, height=H, mode=R.MODE_PRIMARY), want=("z", "node_id", "face_id"))
o = O.render(hs.desc, width=W, height=H, mode=R.MODE_PRIMARY, want=("z", "node_id", "face_id"))
print("ids equal", np.array_equal(g["node_id"], o["node_id"]), "faces equal", np.array_equal(g["face_id"], o["face_id"]),
      "z bit-exact", np.array_equal(g["z"].view("u4"), o["z"].view("u4")), "hit fraction %.3f" % (o["node_id"] >= 0).mean())
p = R.default_params(width=W, height=H, mode=R.MODE_WHITTED, shade_bounces=5)
gw = sc.render(p, want=("rgb",))["rgb"]; st = sc.stats()
ow = O.render(hs.desc, params=p, want=("rgb",))
a, b = gw.astype("f8"), ow["rgb"].astype("f8")
bad = int((np.abs(a - b) > 1e-4 * np.maximum(np.abs(a), np.abs(b)) + 1e-6).any(axis=2).sum())
print("whitted: pixels out of tolerance", bad, "ray counts equal", st["trace_rays"] == ow["stats"]["trace_rays"] and st["shadow_rays"] == ow["stats"]["shadow_rays"])
