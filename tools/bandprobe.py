import sys, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "raytracer-utah_b200", "python")); sys.path.insert(0, os.path.join(ROOT, "tools"))
import rtu_b200 as R, make_synthetic
make_synthetic.ensure(("grid1M",))
hs = R.HostScene(os.path.join(R.SCENES, "synthetic/grid1M.xml"))
ctx = R.Context(0); sc = R.Scene(ctx, hs.desc)
W, H = int(sys.argv[1]), int(sys.argv[2]); nb = int(sys.argv[3])
for b in range(nb):
    p = R.default_params(width=W, height=H, spp=1, pattern=(R.PATTERN_CENTER if len(sys.argv)>4 and sys.argv[4]=='center' else R.PATTERN_REFERENCE), mode=R.MODE_WHITTED, flags=R.FLAG_TIME_KERNELS,
                         row_begin=b * H // nb, row_end=(b + 1) * H // nb)
    for it in range(2):
        sc.render_device(p); st = sc.stats()
    print(b, round(st["device_ms"], 3), {k: round(v["ms"], 3) for k, v in st.items() if isinstance(v, dict)}, st["box_tests"], st["tri_tests"])
