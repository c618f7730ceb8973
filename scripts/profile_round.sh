#!/bin/bash
# ncu evidence for profiles/: launch list (durations only) + one --set full capture of every kernel of the stereo step.
# usage: bash scripts/profile_round.sh <tag>      (outputs under gpurun_out/)
set -u
TAG=${1:-rX}
export ORBX_PIPE=1 ORBX_CHUNK=64   # one 64-frame launch per kernel: the launch shape bench.py's per-stage times use
# 4 steps x 7 kernels (pyramid level 0, pyramid levels, FAST, quadtree, orientation+BRIEF, frame index, stereo) after 3 warm steps
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 21 -c 28 --csv --log-file gpurun_out/launches_$TAG.csv python scripts/profile_driver.py 7 stereo > /dev/null 2>&1
ncu --set full --clock-control none --import-source on --launch-skip 21 -c 7 -o gpurun_out/prof_all_$TAG -f python scripts/profile_driver.py 4 stereo > /dev/null 2>&1
ncu --set full --clock-control none --import-source on --launch-skip 28 -c 20 -o gpurun_out/prof_other_$TAG -f python scripts/profile_driver.py 4 all > /dev/null 2>&1
ls -la gpurun_out | grep $TAG
