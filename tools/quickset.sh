#!/bin/bash
# device times of the four iteration workloads (one JSON line each)
python tools/quickbench.py Teapot/scene2.xml 1920 1080 whitted 64 2>&1 | tail -1
python tools/quickbench.py Project11/scene.xml 800 600 path 16 2>&1 | tail -1
python tools/quickbench.py Project10/scene.xml 800 600 path 16 2>&1 | tail -1
python tools/quickbench.py synthetic/grid1M.xml 3840 2160 whitted 4 2>&1 | tail -1
