#!/bin/bash
# headline bench with the new pyramid / FAST kernels + pipeline shape sweep + whole-step DRAM traffic with natural cache state
mkdir -p gpurun_out
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/n_bench.json 2> gpurun_out/n_bench.err; echo "bench rc=$?"; tail -c 300 gpurun_out/n_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/n_bench.json"))
print("value %.0f e2e %.0f ceiling %.0f" % (d["value"], d["e2e"]["value"], d["e2e"]["platform_ceiling"]["value"]), {k: round(v,3) for k,v in d["roofline"]["stage_ms"].items()})
print("   latency", d["latency"])
print("   tum", d["configs"]["tum_rgbd"]["value"], "hd", d["configs"]["hd_1080p"]["value"], d["configs"]["hd_1080p"]["stage_ms"])
print("   sweep", {k: v["p50_ms_device_graph"] for k, v in d["configs"]["latency_sweep"]["n_features"].items()})
print("   check", d["check"]["gathered_checksum"], d["check"]["oracle_frame0"])
PY
for cfg in "8 8" "4 8" "2 8" "3 8" "2 16" "4 16" "1 64"; do
  set -- $cfg
  ORBX_PIPE=$1 ORBX_CHUNK=$2 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-matchers --no-configs 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('pipe $1 chunk $2: value %.0f  ms/step %.3f  e2e %.0f' % (d['value'], d['ms_per_step'], d['e2e']['value']))"
done 2>&1 | tee gpurun_out/n_tune_pipe.txt
# DRAM bytes of every kernel of one 64-frame batch call, caches NOT flushed between kernels (one pass per kernel: only two metrics)
for cfg in "8 8" "1 64"; do
  set -- $cfg
  ORBX_PIPE=$1 ORBX_CHUNK=$2 timeout 600 ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --launch-skip 21 -c 70 --csv --log-file gpurun_out/n_traffic_p$1c$2.csv python scripts/profile_driver.py 4 stereo > /dev/null 2>&1
done
ls -la gpurun_out | grep n_
