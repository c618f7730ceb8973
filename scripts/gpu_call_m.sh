#!/bin/bash
# warp-per-cell FAST: parity suite, stage times, ncu of the FAST kernel, memcheck/racecheck of a small driver
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=10 > gpurun_out/m_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/m_pytest.log
tail -25 gpurun_out/m_pytest.log
timeout 300 python scripts/stage_times.py 1 8 64 2>&1 | tee gpurun_out/m_stage_times.txt
export ORBX_PIPE=1 ORBX_CHUNK=64
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fast_cells --launch-skip 3 -c 1 -o gpurun_out/prof_fast_m -f python scripts/profile_driver.py 4 stereo > /dev/null 2>&1
ls -la gpurun_out | grep prof_fast_m
timeout 900 bash scripts/sanitize.sh 2>&1 | tee gpurun_out/m_sanitize.txt | grep -E "==|exit|SUMMARY"
