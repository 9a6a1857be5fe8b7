# round-1f call K: final launch lists (default workload, top-k)
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_cfg2_r1f.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_cfg2_r1f.log 2>&1
echo "ncu cfg2 exit $?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_topk_r1f.csv python bench.py --steps 2 --warmup 3 --workload yago-complex-d256-topk > gpurun_out/ncu_topk_r1f.log 2>&1
echo "ncu topk exit $?"
