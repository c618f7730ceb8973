#!/bin/bash
# 2-GPU call: NCCL / peer-memory sequence tests under torchrun, bench at N=2 with both transports
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sequence.py -m gpu -x -q > gpurun_out/b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/b_pytest.log
tail -8 gpurun_out/b_pytest.log
for tr in nccl peer; do
  NCCL_DEBUG=WARN timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 --transport $tr > gpurun_out/b_bench2_$tr.json 2> gpurun_out/b_bench2_$tr.err
  echo "bench2 $tr rc=$?"
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/b_bench2_$tr.json"))
    print("$tr: value %.0f e2e %.0f ms/step %.2f transport %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["transport"]), d["check"])
except Exception as e:
    print("$tr failed", e)
PY
  tail -3 gpurun_out/b_bench2_$tr.err
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 scripts/pcie_ceiling.py > gpurun_out/b_ceiling2.json 2>gpurun_out/b_ceiling2.err
cat gpurun_out/b_ceiling2.json
timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 --no-configs --no-matchers --no-cpu-baseline > gpurun_out/b_bench1.json 2> gpurun_out/b_bench1.err
python -c "
import json; d=json.load(open('gpurun_out/b_bench1.json')); print('N=1: value %.0f e2e %.0f' % (d['value'], d['e2e']['value']), d['latency'], d['check']['gathered_checksum'])"
