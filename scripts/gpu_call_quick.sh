#!/bin/bash
# quick validation: parity suite + per-stage times (+ optional ncu of one kernel: NCU_KERNEL=regex)
mkdir -p gpurun_out
TAG=${TAG:-q}
timeout 900 python -m pytest tests -m gpu -q --maxfail=10 > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -12 gpurun_out/${TAG}_pytest.log
timeout 300 python scripts/stage_times.py 1 8 64 2>&1 | tee gpurun_out/${TAG}_stage_times.txt
if [ -n "${NCU_KERNEL:-}" ]; then
  export ORBX_PIPE=1 ORBX_CHUNK=64
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$NCU_KERNEL --launch-skip ${NCU_SKIP:-3} -c ${NCU_COUNT:-1} -o gpurun_out/prof_${TAG} -f python scripts/profile_driver.py 4 stereo > /dev/null 2>&1
  ls -la gpurun_out | grep prof_${TAG}
fi
if [ -n "${EXTRA_ENV:-}" ]; then
  echo "--- with $EXTRA_ENV"
  env $EXTRA_ENV timeout 300 python scripts/stage_times.py 1 64 2>&1 | tee gpurun_out/${TAG}_stage_times_extra.txt
fi
