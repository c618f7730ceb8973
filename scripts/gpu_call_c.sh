#!/bin/bash
# quadtree fast path: correctness first, then timing with and without it
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_quadtree.py -m gpu -x -q > gpurun_out/c_pytest_qt.log 2>&1; echo "pytest qt rc=$?" >> gpurun_out/c_pytest_qt.log
tail -25 gpurun_out/c_pytest_qt.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c_pytest.log
tail -8 gpurun_out/c_pytest.log
for fast in 1 0; do
  ORBX_QT_FAST=$fast timeout 400 python bench.py --steps 5 --warmup 3 --no-matchers --no-cpu-baseline > gpurun_out/c_bench_fast$fast.json 2> gpurun_out/c_bench_fast$fast.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/c_bench_fast$fast.json"))
    print("QT_FAST=$fast: value %.0f e2e %.0f" % (d["value"], d["e2e"]["value"]), {k: round(v,3) for k,v in d["roofline"]["stage_ms"].items()})
    print("   latency", d["latency"])
    print("   hd", d["configs"]["hd_1080p"]["value"], d["configs"]["hd_1080p"]["stage_ms"], d["configs"]["hd_1080p"]["p50_ms_device_single_pair"])
    print("   tum", d["configs"]["tum_rgbd"]["value"])
    print("   sweep", {k: v["p50_ms_device_graph"] for k, v in d["configs"]["latency_sweep"]["n_features"].items()})
    print("   check", d["check"]["gathered_checksum"], d["check"]["oracle_frame0"])
except Exception as e:
    print("fast=$fast failed", e)
PY
  tail -3 gpurun_out/c_bench_fast$fast.err
done
