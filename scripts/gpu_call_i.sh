#!/bin/bash
# evidence call: full GPU tests, smoke, bench (N=1, full line), ncu launch list + full captures, sanitizer, reference arm
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/i_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/i_pytest.log
tail -4 gpurun_out/i_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/i_bench.json 2> gpurun_out/i_bench.err; echo "bench rc=$?"; tail -c 400 gpurun_out/i_bench.err
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/i_bench_reference.json 2> gpurun_out/i_bench_reference.err; echo "ref rc=$?"
timeout 900 bash scripts/profile_round.sh r2 2>&1 | tail -4
timeout 300 python scripts/stage_times.py 1 8 64 128 256 2>&1 | tee gpurun_out/i_stage_times.txt
timeout 1200 bash scripts/sanitize.sh 2>&1 | tee gpurun_out/i_sanitize.txt | grep -E "==|exit|SUMMARY"
