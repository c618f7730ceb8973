#!/bin/bash
mkdir -p gpurun_out
ORBX_QT_FAST=1 timeout 200 python scripts/devtests/qt_debug.py 2>&1 | tail -8
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/h_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/h_pytest.log
tail -5 gpurun_out/h_pytest.log
timeout 300 python scripts/devtests/qt_phases.py 2>&1 | tee gpurun_out/h_phases.txt
timeout 400 python bench.py --steps 5 --warmup 3 --no-matchers --no-cpu-baseline > gpurun_out/h_bench.json 2> gpurun_out/h_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/h_bench.json"))
print("value %.0f e2e %.0f" % (d["value"], d["e2e"]["value"]), {k: round(v,3) for k,v in d["roofline"]["stage_ms"].items()})
print("   latency", d["latency"])
print("   hd", d["configs"]["hd_1080p"]["value"], d["configs"]["hd_1080p"]["stage_ms"], d["configs"]["hd_1080p"]["p50_ms_device_single_pair"])
print("   sweep", {k: v["p50_ms_device_graph"] for k, v in d["configs"]["latency_sweep"]["n_features"].items()})
print("   check", d["check"]["gathered_checksum"])
PY
