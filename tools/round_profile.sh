#!/bin/bash
# everything the round's profiles/ directory is made from, in two calls (gpurun copies back at most 64 MiB per call):
#   tools/round_profile.sh <tag> bench     tests, bench (both arms), launch list, config table, photon, fixed cost
#   tools/round_profile.sh <tag> ncu       ncu --set full captures of the frame kernels (Teapot, Project11 GI) and the photon estimate
T=${1:-rXX}
O=gpurun_out
if [ "$2" != "ncu" ]; then
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3 > $O/${T}_tests.log
python bench.py > $O/${T}_bench.json 2> $O/${T}_bench.err
python bench.py --impl reference > $O/${T}_ref.json 2> $O/${T}_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/${T}_launches.csv python bench.py --steps 1 --warmup 1 > $O/${T}_ncu.log 2>&1
python tools/config_table.py > $O/${T}_configs.md 2> $O/${T}_configs.err
python tools/photonbench.py > $O/${T}_photon.txt 2>&1
python tools/fixedcost.py > $O/${T}_fixed.txt 2>&1
else
ncu --set full --import-source on --clock-control none -k regex:"k_extend_pool|k_shade|k_shadow_wave" -c 4 -o $O/${T}_teapot python tools/quickbench.py Teapot/scene2.xml 1920 1080 whitted 64 > $O/${T}_ncu_teapot.log 2>&1
ncu --set full --clock-control none -k regex:"k_extend_pool|k_shade|k_shadow_wave" -c 3 -o $O/${T}_gi python tools/quickbench.py Project11/scene.xml 800 600 path 16 > $O/${T}_ncu_gi.log 2>&1
ncu --set full --clock-control none -k regex:"knn_cand|knn_replay" -c 2 -o $O/${T}_photon python tools/photonbench.py > $O/${T}_ncu_photon.log 2>&1
fi
du -sh $O; ls -la $O | tail -12
