#!/usr/bin/env python3
"""Runs tools/quickbench.py <args> once per tuning variant in vbuild/ (tools/build_variants.sh)."""
import sys, os, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for f in sorted(os.listdir(os.path.join(ROOT, "vbuild"))):
    env = dict(os.environ, RTU_B200_LIB=os.path.join(ROOT, "vbuild", f))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "quickbench.py")] + sys.argv[1:], env=env, capture_output=True, text=True)
    for line in r.stdout.strip().splitlines():
        print(f, line)
    if r.returncode:
        print(f, "FAILED", r.stderr.strip()[-300:])
