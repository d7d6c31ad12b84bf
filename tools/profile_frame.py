#!/usr/bin/env python3
"""One small Whitted frame for ncu captures: python tools/profile_frame.py [spp] [scene] [W H]

Kept short on purpose (ncu replays every kernel ~40 times).  Prints the frame's counters."""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "raytracer-utah_b200", "python"))
import rtu_b200 as R

spp = int(sys.argv[1]) if len(sys.argv) > 1 else 4
scene = sys.argv[2] if len(sys.argv) > 2 else "Teapot/scene2.xml"
W = int(sys.argv[3]) if len(sys.argv) > 3 else 1920
H = int(sys.argv[4]) if len(sys.argv) > 4 else 1080
hs = R.HostScene(os.path.join(R.SCENES, scene))
ctx = R.Context(0)
sc = R.Scene(ctx, hs.desc)
p = R.default_params(width=W, height=H, spp=spp, pattern=R.PATTERN_REFERENCE, mode=R.MODE_WHITTED, flags=R.FLAG_TIME_KERNELS)
sc.render_device(p)
st = sc.stats()
print(json.dumps(st))
sc.close()
ctx.close()
