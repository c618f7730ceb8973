#!/bin/bash
# first GPU call of round 2: tests, smoke, bench, pipeline-shape sweep, PCIe ceiling
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/a_gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/a_pytest.log
tail -5 gpurun_out/a_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/a_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/a_smoke.log
tail -3 gpurun_out/a_smoke.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err; echo "bench rc=$?"
tail -c 600 gpurun_out/a_bench.err
timeout 120 python scripts/pcie_ceiling.py > gpurun_out/a_ceiling.json 2>&1
cat gpurun_out/a_ceiling.json
for cfg in "8 8" "4 16" "2 32" "8 4" "2 8" "4 8" "8 16"; do
  set -- $cfg
  ORBX_PIPE=$1 ORBX_CHUNK=$2 timeout 300 python bench.py --steps 3 --warmup 3 --no-configs --no-matchers --no-cpu-baseline > gpurun_out/a_sweep_$1_$2.json 2> gpurun_out/a_sweep_$1_$2.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/a_sweep_$1_$2.json"))
    print("pipe $1 chunk $2: value %.0f e2e %.0f stage_ms %s" % (d["value"], d["e2e"]["value"], {k: round(v,3) for k,v in d["roofline"]["stage_ms"].items()}))
except Exception as e:
    print("pipe $1 chunk $2: failed", e)
PY
done
