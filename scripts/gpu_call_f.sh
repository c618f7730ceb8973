#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/f_pytest.log
tail -5 gpurun_out/f_pytest.log
for i in 1 2 3; do timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "random_configurations or extract_matches" 2>&1 | tail -2; done
timeout 1200 bash scripts/sanitize.sh 2>&1 | tee gpurun_out/f_sanitize.txt | grep -E "==|exit|SUMMARY"
