#!/usr/bin/env python3
"""BASELINE config 5 (Project13 photon map): emission Mrays/s, kd-tree build ms, gather Mqueries/s at 800x600 and 1080p."""
import sys, os, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "raytracer-utah_b200", "python"))
import numpy as np
import rtu_b200 as R
hs = R.HostScene(os.path.join(R.SCENES, "Project13/scene.xml"))
ctx = R.Context(0); sc = R.Scene(ctx, hs.desc)
best = None
for it in range(3):
    t0 = time.perf_counter()
    st = sc.photon_map_generate(seed=it)
    st["wall_ms"] = (time.perf_counter() - t0) * 1e3
    if best is None or st["emit_ms"] < best["emit_ms"]: best = st
best["emit_mrays_per_s"] = best["trace_rays"] / best["emit_ms"] * 1e-3
print(json.dumps({"photon_map": best}))
# the device kd-tree build alone, on a map without tied coordinates (this scene's own map has ties: axis-aligned walls)
rng = np.random.default_rng(7)
big = np.zeros(1000000, R.PHOTON_DTYPE)
for k in range(3):
    big["position"][:, k] = (rng.permutation(len(big)).astype("f4") - 500000.0) * np.float32(0.001 * (k + 1))
big["power"] = 1
tb = None
for it in range(3):
    t0 = time.perf_counter(); sc.photon_map_set(big); dt = (time.perf_counter() - t0) * 1e3
    tb = dt if tb is None else min(tb, dt)
print(json.dumps({"device_kd_build_1M_distinct": {"wall_ms_incl_24MB_upload": round(tb, 2), "on_device": sc.photon_map_info()[1]}}))
sc.photon_map_generate(seed=0)
for W, H in ((800, 600), (1920, 1080)):
    p = R.default_params(width=W, height=H, spp=1, pattern=R.PATTERN_CENTER, mode=R.MODE_PHOTON, flags=R.FLAG_TIME_KERNELS)
    out = sc.render(p, want=("node_id",))
    hits = int((out["node_id"] >= 0).sum())
    b = None
    for it in range(3):
        sc.render_device(p); s = sc.stats()
        if b is None or s["device_ms"] < b["device_ms"]: b = s
    g = b["shade_kernels"]["ms"]
    print(json.dumps({"size": [W, H], "queries": hits, "gather_ms": round(g, 3), "gather_mqueries_per_s": round(hits / g * 1e-3, 2),
                      "primary_ms": round(b["primary_wave"]["ms"], 3), "frame_ms": round(b["device_ms"], 3)}))
