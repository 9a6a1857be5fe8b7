# round-1f call D: 512-thread pair_fwd, 3-instruction L1 backward (half tables), per-triple prefetch gating
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --timeout 600 -x > gpurun_out/t_all2.log 2>&1; echo "exit $? all gpu tests"; tail -8 gpurun_out/t_all2.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --workload wikikg2-transe-l1-d256-bf16 > gpurun_out/bench_wiki3.json 2> gpurun_out/bench_wiki3.err; echo "bench wiki exit $?"; tail -3 gpurun_out/bench_wiki3.err; cat gpurun_out/bench_wiki3.json
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --workload biokg-transe-l2-d128-fp32 > gpurun_out/bench_l2.json 2> gpurun_out/bench_l2.err; echo "bench l2 exit $?"; tail -3 gpurun_out/bench_l2.err; cat gpurun_out/bench_l2.json
for w in wikikg2-rotate-d512-scoremoving wikikg2-pairre-d512-scoremoving; do
  timeout 600 python bench.py --steps 10 --warmup 3 --workload $w > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "bench $w exit $?"; tail -3 gpurun_out/bench_$w.err; cat gpurun_out/bench_$w.json
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_wiki3.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --workload wikikg2-transe-l1-d256-bf16 > gpurun_out/ncu_wiki3.log 2>&1
echo "ncu launches exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pair_ -c 3 -f -o gpurun_out/prof_pair3 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --workload wikikg2-transe-l1-d256-bf16 > gpurun_out/ncu_pair3.log 2>&1
echo "ncu pair exit $?"
