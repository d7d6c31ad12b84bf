import sys, os, json, subprocess
# runs tools/quickbench-style timing for every tuning variant in variants/
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import sys, os, json
sys.path.insert(0, os.path.join(%r, "raytracer-utah_b200", "python"))
import rtu_b200 as R
hs = R.HostScene(os.path.join(R.SCENES, "Teapot/scene2.xml"))
ctx = R.Context(0); sc = R.Scene(ctx, hs.desc)
p = R.default_params(width=1920, height=1080, spp=16, pattern=R.PATTERN_REFERENCE, mode=R.MODE_WHITTED, flags=R.FLAG_TIME_KERNELS)
best = None
for it in range(4):
    sc.render_device(p); st = sc.stats()
    if best is None or st["device_ms"] < best["device_ms"]: best = st
print(json.dumps({k: (round(v["ms"],3) if isinstance(v, dict) else v) for k, v in best.items() if k in ("device_ms","primary_wave","secondary_waves","shadow_waves","shade_kernels")}))
''' % ROOT
for f in sorted(os.listdir(os.path.join(ROOT, "variants"))):
    env = dict(os.environ, RTU_B200_LIB=os.path.join(ROOT, "variants", f))
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    print(f, r.stdout.strip(), r.stderr.strip()[-300:])
