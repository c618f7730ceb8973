#!/bin/bash
# ncu evidence for profiles/: launch list (durations only) + one --set full capture of every kernel.
# usage: bash scripts/profile_round.sh <tag>      (outputs under gpurun_out/)
set -u
TAG=${1:-rX}
export ORBX_PIPE=1 ORBX_CHUNK=64   # one 64-frame launch per kernel: the launch shape bench.py's per-stage times use
B="python bench.py --steps 2 --warmup 3 --pool 64 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 21 -c 28 --csv --log-file gpurun_out/launches_$TAG.csv $B --no-matchers > /dev/null 2>&1
ncu --set full --clock-control none --import-source on --launch-skip 21 -c 7 -o gpurun_out/prof_all_$TAG -f $B --no-matchers > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:area_match -c 1 -o gpurun_out/prof_area_$TAG -f $B > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:serialize -c 1 -o gpurun_out/prof_ser_$TAG -f $B > /dev/null 2>&1
ls -la gpurun_out | grep $TAG
