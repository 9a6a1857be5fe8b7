cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --workload yago-complex-d256-topk > gpurun_out/plain_topk.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_topk.csv python bench.py --steps 2 --warmup 3 --workload yago-complex-d256-topk > gpurun_out/ncu_topk.log 2>&1
echo "ncu topk exit $?"
