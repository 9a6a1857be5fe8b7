#!/bin/bash
# usage: scripts/gpu_step.sh TAG [bench args...]  — plain bench run, then (after it exited 0) the ncu
# launch list of the same command; outputs under gpurun_out/TAG_*
TAG=$1; shift
python bench.py --steps 20 --warmup 5 "$@" > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err || { tail -5 gpurun_out/${TAG}_bench.err; exit 1; }
python - <<PY
import json; d=json.load(open("gpurun_out/${TAG}_bench.json"))
print("${TAG}", "value", round(d["value"]/1e6,2), "M/s ms", round(d["ms_per_step"],4), "e2e ms", round(d["e2e"]["ms_per_step"],4), "sec", d.get("secondary",{}).get("ms_per_step"), (d.get("secondary") or {}).get("e2e",{}).get("ms_per_step"), "parity", (d.get("parity_check") or {}).get("ok"))
PY
if [ -n "$NCU" ]; then
  ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 3 --warmup 3 --step-only "$@" > gpurun_out/${TAG}_ncu.log 2>&1
  python scripts/summarize_launches.py gpurun_out/${TAG}_launches.csv > gpurun_out/${TAG}_launches.md 2>/dev/null; head -30 gpurun_out/${TAG}_launches.md
fi
