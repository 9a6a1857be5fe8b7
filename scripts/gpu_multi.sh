#!/bin/bash
# usage: scripts/gpu_multi.sh N TAG [bench args...] — torchrun bench on N GPUs of one box
N=$1; TAG=$2; shift 2
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus $N --steps 20 --warmup 5 "$@" > gpurun_out/${TAG}_n${N}.json 2> gpurun_out/${TAG}_n${N}.err
rc=$?
grep stage_us gpurun_out/${TAG}_n${N}.err > gpurun_out/${TAG}_n${N}_stages.json
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${TAG}_n${N}.json"))
    print("${TAG} N=$N", "value", round(d["value"]/1e6,2), "ms", round(d["ms_per_step"],4), "e2e ms", round((d.get("e2e") or {}).get("ms_per_step",0),4), "sec", (d.get("secondary") or {}).get("ms_per_step"), "parity", (d.get("parity_check") or {}).get("ok"))
except Exception as e:
    print("${TAG} N=$N failed", e, open("gpurun_out/${TAG}_n${N}.err").read()[-1500:])
PY
exit $rc
