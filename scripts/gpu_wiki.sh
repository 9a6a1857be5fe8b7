cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_bess.py -q -m gpu --timeout 300 -k "shared or training or forward" 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 --workload wikikg2-transe-l1-d256-bf16 --no-cpu-baseline > gpurun_out/bench_wiki.json 2> gpurun_out/bench_wiki.err; echo "bench exit $?"; tail -3 gpurun_out/bench_wiki.err; cut -c1-400 gpurun_out/bench_wiki.json
