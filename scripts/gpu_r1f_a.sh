# round-1f call A: new top-k merge + table-operand cache tests, top-k bench, ncu --set full of the L1 pair kernels
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu --timeout 300 -k "topk_merge or table_operand" > gpurun_out/t_newkernels.log 2>&1; echo "exit $? new kernels"; tail -15 gpurun_out/t_newkernels.log
timeout 600 python -m pytest tests/test_gpu_bess.py -q -m gpu --timeout 300 -k "topk" > gpurun_out/t_topk.log 2>&1; echo "exit $? topk"; tail -8 gpurun_out/t_topk.log
timeout 300 python bench.py --steps 10 --warmup 3 --workload yago-complex-d256-topk > gpurun_out/bench_topk2.json 2> gpurun_out/bench_topk2.err; echo "bench topk exit $?"; tail -3 gpurun_out/bench_topk2.err; cat gpurun_out/bench_topk2.json
BESS_TOPK_MERGE_V1=1 timeout 300 python bench.py --steps 10 --warmup 3 --workload yago-complex-d256-topk > gpurun_out/bench_topk2_v1.json 2>/dev/null; cat gpurun_out/bench_topk2_v1.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pair_ -c 3 -f -o gpurun_out/prof_pair python bench.py --steps 1 --warmup 3 --no-cpu-baseline --workload wikikg2-transe-l1-d256-bf16 > gpurun_out/ncu_pair.log 2>&1
echo "ncu pair exit $?"
