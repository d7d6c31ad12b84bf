import json, os, sys, time
ROOT = "/root/repo"
sys.path.insert(0, os.path.join(ROOT, "raytracer-utah_b200", "python"))
import rtu_b200 as R
CASES = [("Teapot/scene2.xml", 1920, 1080, R.MODE_WHITTED), ("Project4.xml", 800, 600, R.MODE_WHITTED), ("Project11/scene.xml", 800, 600, R.MODE_PATH)]
ctx = R.Context(0)
for scene, W, H, mode in CASES:
    hs = R.HostScene(os.path.join(R.SCENES, scene))
    sc = R.Scene(ctx, hs.desc)
    for spp in (1, 4):
        p = R.default_params(width=W, height=H, spp=spp, pattern=R.PATTERN_REFERENCE if spp > 1 else R.PATTERN_CENTER, mode=mode)
        dev = []
        for _ in range(60):
            sc.render_device(p)
            dev.append(sc.stats()["device_ms"])
        st = sc.stats()
        print(os.environ.get("RTU_TAIL_RAYS"), scene, spp, round(min(dev[5:]), 4), st["kernel_launches"])
    sc.close(); hs.close()
