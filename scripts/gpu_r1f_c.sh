# round-1f call C: FP32 pipe micro-benchmark, AllScores pipeline tests, per-triple prefetch kernel, ncu of the pair kernels
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
./scripts/ubench/fp32_pipes > gpurun_out/ubench_fp32.log 2>&1; cat gpurun_out/ubench_fp32.log
timeout 900 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_kernels.py -q -m gpu --timeout 300 > gpurun_out/t_pipe.log 2>&1; echo "exit $? pipeline+kernels"; tail -30 gpurun_out/t_pipe.log
timeout 600 python -m pytest tests/test_gpu_bess.py -q -m gpu --timeout 300 -x > gpurun_out/t_bess.log 2>&1; echo "exit $? bess"; tail -5 gpurun_out/t_bess.log
for w in wikikg2-rotate-d512-scoremoving wikikg2-pairre-d512-scoremoving; do
  timeout 600 python bench.py --steps 10 --warmup 3 --workload $w > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "bench $w exit $?"; tail -3 gpurun_out/bench_$w.err; cat gpurun_out/bench_$w.json
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pair_ -c 3 -f -o gpurun_out/prof_pair2 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --workload wikikg2-transe-l1-d256-bf16 > gpurun_out/ncu_pair2.log 2>&1
echo "ncu pair exit $?"
