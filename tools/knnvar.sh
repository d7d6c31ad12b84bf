#!/bin/bash
# photon gather time of every vbuild/knn_*.so variant
for f in vbuild/knn_*.so; do echo "== $f"; RTU_B200_LIB=$PWD/$f python tools/photonbench.py 2>&1 | tail -2; done
