import sys, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "raytracer-utah_b200", "python")); sys.path.insert(0, os.path.join(ROOT, "tools")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import rtu_b200 as R
from conftest import load_golden, synthetic_scene
name = sys.argv[1]
g, meta = load_golden("synthetic_" + name)
hs = R.HostScene(synthetic_scene(name, meta))
ctx = R.Context(0); sc = R.Scene(ctx, hs.desc)
p = R.default_params(width=meta["width"], height=meta["height"], mode=R.MODE_WHITTED, shade_bounces=5)
out = sc.render(p, want=("rgb",)); st = sc.stats()
a, b = out["rgb"].astype("f8"), g["rgb"].astype("f8")
bad = ~(np.abs(a - b) <= 1e-4 * np.maximum(np.abs(a), np.abs(b)) + 1e-6).all(axis=2)
for y, x in zip(*np.nonzero(bad)):
    print("pixel", x, y, "gpu", out["rgb"][y, x], "ref", g["rgb"][y, x], "node", g["node"][y, x])
print("rays gpu", st["trace_rays"], st["shadow_rays"], "ref", meta["whitted"]["trace_rays"], meta["whitted"]["shadow_rays"])
