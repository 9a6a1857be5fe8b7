cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python bench.py --steps 5 --warmup 3 --workload yago-complex-d256-topk > gpurun_out/bench_topk.json 2> gpurun_out/bench_topk.err; echo "bench exit $?"; tail -5 gpurun_out/bench_topk.err; cat gpurun_out/bench_topk.json
