#!/bin/bash
# tests + smoke + bench on the GPU box; summaries into gpurun_out/
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_final.json"))
print("ours", round(d["value"], 1), round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), d["clocks"], d["gpu_launches"], round(d["roofline"]["frac"], 3), d["cpu_baseline"]["value"])
PY
