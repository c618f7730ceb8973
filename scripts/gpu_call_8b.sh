#!/bin/bash
# 8-GPU call on the final kernels: the sequence bench at N = 8 (both transports), 4 and 2
mkdir -p gpurun_out
R="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for cfg in "8 peer" "8 nccl" "4 peer" "2 peer"; do
  set -- $cfg
  timeout 400 $R --nproc-per-node $1 --master-port 2962$1 bench.py --gpus $1 --steps 20 --warmup 5 --transport $2 --no-matchers --no-configs > gpurun_out/m_bench$1_$2.json 2> gpurun_out/m_bench$1_$2.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/m_bench$1_$2.json"))
    print("N=$1 $2: value %.0f e2e %.0f (ceiling %.0f) ms/step %.2f" % (d["value"], d["e2e"]["value"], d["e2e"]["platform_ceiling"]["value"], d["ms_per_step"]), d["check"]["gathered_checksum"], d["check"]["gathered_checksum_equal_on_all_ranks"], d["check"]["sampled_frames_vs_single_frame_call"]["identical"], d["clocks"])
except Exception as e:
    print("N=$1 $2 failed", e)
PY
  tail -2 gpurun_out/m_bench$1_$2.err
done
