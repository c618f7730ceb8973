#!/bin/bash
# pyramid split (level 0 / resized levels): parity suite, stage times, ncu of the two pyramid kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/j_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j_pytest.log
tail -15 gpurun_out/j_pytest.log
timeout 300 python scripts/stage_times.py 1 8 64 2>&1 | tee gpurun_out/j_stage_times.txt
export ORBX_PIPE=1 ORBX_CHUNK=64
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pyramid --launch-skip 6 -c 2 -o gpurun_out/prof_pyr_j -f python scripts/profile_driver.py 4 stereo > /dev/null 2>&1
ls -la gpurun_out | grep prof_pyr_j
