#!/bin/bash
# photon gather: tests, bench, and per-kernel ncu metrics of the estimate kernels
timeout 600 python -m pytest tests -m gpu -q -x -k "photon or gather" 2>&1 | tail -3
python tools/photonbench.py 2>&1 | tail -2
ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__inst_executed.avg.per_cycle_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum --clock-control none -k regex:"knn_cand|knn_replay" -c 4 --csv --log-file gpurun_out/knn_ncu.csv python tools/photonbench.py > /dev/null 2>&1
python - <<'PY'
import csv
rows=list(csv.DictReader(l for l in open('gpurun_out/knn_ncu.csv') if l.startswith('"')))
d={}
for r in rows: d.setdefault(r['ID'],{'k':r['Kernel Name'][:18]})[r['Metric Name'].split('__')[1][:28]]=r['Metric Value']
for i,v in d.items(): print(i,v)
PY
