#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/e_pytest.log
tail -6 gpurun_out/e_pytest.log
timeout 300 python scripts/stage_times.py 1 8 64 128 256 2>&1 | tee gpurun_out/e_stage_times.txt
timeout 900 bash scripts/profile_round.sh r2a 2>&1 | tail -5
timeout 1500 bash scripts/sanitize.sh 2>&1 | tee gpurun_out/e_sanitize.txt | tail -20
