# round-1f call H: device feed tests + full GPU suite + bench lines with the traffic field
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_device_feed.py -q -m gpu --timeout 300 > gpurun_out/t_feed.log 2>&1; echo "exit $? device feed"; tail -30 gpurun_out/t_feed.log
timeout 1200 python -m pytest tests -q -m gpu --timeout 600 > gpurun_out/t_all6.log 2>&1; echo "exit $? all gpu tests"; tail -5 gpurun_out/t_all6.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --workload wikikg2-transe-l1-d256-bf16 > gpurun_out/bench_wiki6.json 2> gpurun_out/bench_wiki6.err; echo "bench wiki exit $?"; tail -2 gpurun_out/bench_wiki6.err; cut -c1-300 gpurun_out/bench_wiki6.json
