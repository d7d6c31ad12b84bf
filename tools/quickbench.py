import sys, os, time, json
sys.path.insert(0, 'raytracer-utah_b200/python')
import numpy as np, rtu_b200 as R
hs = R.HostScene(os.path.join(R.SCENES, 'Teapot/scene2.xml'))
ctx = R.Context(0); sc = R.Scene(ctx, hs.desc)
for spp in (1, 16, 64):
    p = R.default_params(width=1920, height=1080, spp=spp, pattern=R.PATTERN_REFERENCE, mode=R.MODE_WHITTED)
    for it in range(3):
        sc.render_device(p); st = sc.stats()
    rays = st['trace_rays'] + st['shadow_rays']
    print(json.dumps(dict(spp=spp, rays=rays, ms=st['device_ms'], mrays=rays/st['device_ms']*1e-3, launches=st['kernel_launches'], box=st['box_tests'], tri=st['tri_tests'], nodes=st['node_visits'])))
