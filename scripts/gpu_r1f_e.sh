# round-1f call E: double-buffered pair_fwd, occupancy hints on the L1 backward kernels
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --timeout 600 -x > gpurun_out/t_all3.log 2>&1; echo "exit $? all gpu tests"; tail -4 gpurun_out/t_all3.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --workload wikikg2-transe-l1-d256-bf16 > gpurun_out/bench_wiki4.json 2> gpurun_out/bench_wiki4.err; echo "bench wiki exit $?"; tail -3 gpurun_out/bench_wiki4.err; cat gpurun_out/bench_wiki4.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pair_ -c 3 -f -o gpurun_out/prof_pair4 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --workload wikikg2-transe-l1-d256-bf16 > gpurun_out/ncu_pair4.log 2>&1
echo "ncu pair exit $?"
