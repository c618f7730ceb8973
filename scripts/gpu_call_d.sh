#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_quadtree.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python scripts/devtests/qt_phases.py 2>&1 | tee gpurun_out/d_phases.txt
ORBX_QT_FAST=1 timeout 400 python bench.py --steps 5 --warmup 3 --no-matchers --no-cpu-baseline > gpurun_out/d_bench.json 2> gpurun_out/d_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/d_bench.json"))
print("value %.0f e2e %.0f" % (d["value"], d["e2e"]["value"]), {k: round(v,3) for k,v in d["roofline"]["stage_ms"].items()})
print("   latency", d["latency"])
print("   hd", d["configs"]["hd_1080p"]["value"], d["configs"]["hd_1080p"]["stage_ms"], d["configs"]["hd_1080p"]["p50_ms_device_single_pair"])
print("   sweep", {k: v["p50_ms_device_graph"] for k, v in d["configs"]["latency_sweep"]["n_features"].items()})
PY
