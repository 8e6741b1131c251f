#!/bin/bash
# 2-GPU validation of the NVLink paths + multi-GPU kernel table (push vs pull gradients). Run with gpurun --gpus 2.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sharded_ebc_nccl_gpu.py -x -q 2>&1 | tail -4 > gpurun_out/g2_pytest.log; cat gpurun_out/g2_pytest.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $T --master-port 29521 tools/step_profile.py --gpus 2 > gpurun_out/g2_profile_push.md 2>&1
TRB_GRAD_PUSH=0 timeout 300 $T --master-port 29522 tools/step_profile.py --gpus 2 > gpurun_out/g2_profile_pull.md 2>&1
timeout 300 $T --master-port 29523 bench.py --gpus 2 --steps 30 --warmup 5 2>&1 | tail -1 > gpurun_out/bench2_push.json
TRB_GRAD_PUSH=0 timeout 300 $T --master-port 29524 bench.py --gpus 2 --steps 30 --warmup 5 2>&1 | tail -1 > gpurun_out/bench2_pull.json
for f in gpurun_out/bench2_push.json gpurun_out/bench2_pull.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[1], d["value"], d["ms_per_step"], d["e2e"]["value"], d["clocks"])
except Exception as e:
    print(sys.argv[1], "FAILED", e, open(sys.argv[1]).read()[-400:])
PY
done
grep -v "^$" gpurun_out/g2_profile_push.md | tail -32
