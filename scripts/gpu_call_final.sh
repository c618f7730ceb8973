#!/bin/bash
# evidence call: full GPU tests, smoke, bench (N=1, full line), reference arm, ncu launch list + full captures, stage times,
# sanitizer, whole-step DRAM traffic with natural cache state.   usage: bash scripts/gpu_call_final.sh <tag>
TAG=${1:-f}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -4 gpurun_out/${TAG}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/${TAG}_smoke.txt
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; tail -c 400 gpurun_out/${TAG}_bench.err
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; echo "ref rc=$?"
timeout 900 bash scripts/profile_round.sh ${TAG} 2>&1 | tail -4
timeout 300 python scripts/stage_times.py 1 8 64 128 256 2>&1 | tee gpurun_out/${TAG}_stage_times.txt
timeout 1200 bash scripts/sanitize.sh 2>&1 | tee gpurun_out/${TAG}_sanitize.txt | grep -E "==|exit|SUMMARY"
# DRAM bytes of every kernel of one 64-frame batch call, caches NOT flushed between kernels (two metrics: one pass per kernel)
ORBX_PIPE=8 ORBX_CHUNK=8 timeout 600 ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --launch-skip 168 -c 56 --csv --log-file gpurun_out/${TAG}_traffic_p8c8.csv python scripts/profile_driver.py 4 stereo > /dev/null 2>&1
ORBX_PIPE=1 ORBX_CHUNK=64 timeout 600 ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --launch-skip 21 -c 7 --csv --log-file gpurun_out/${TAG}_traffic_p1c64.csv python scripts/profile_driver.py 4 stereo > /dev/null 2>&1
ls -la gpurun_out | grep ${TAG}_
