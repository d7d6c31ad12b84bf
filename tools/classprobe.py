import sys, os, json
sys.path.insert(0, '/root/repo/raytracer-utah_b200/python')
import rtu_b200 as R
hs = R.HostScene(os.path.join(R.SCENES, 'Teapot/scene2.xml'))
ctx = R.Context(0); sc = R.Scene(ctx, hs.desc)
p = R.default_params(width=1920, height=1080, spp=16, pattern=R.PATTERN_REFERENCE, mode=R.MODE_WHITTED, flags=R.FLAG_TIME_KERNELS)
for it in range(2):
    sc.render_device(p); st = sc.stats()
for k in ('primary_wave','secondary_waves','shadow_waves','shade_kernels'):
    v = st[k]; print(k, {a: v[a] for a in v}, 'box/ray %.1f tri/ray %.1f node/ray %.1f Mrays/s %.0f' % (v['box_tests']/max(1,v['rays']), v['tri_tests']/max(1,v['rays']), v['node_visits']/max(1,v['rays']), v['rays']/max(1e-9,v['ms'])*1e-3))
