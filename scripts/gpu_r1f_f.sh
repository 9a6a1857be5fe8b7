# round-1f call F: 7-resident bwd_q, unroll-8 fwd, vectorised segment / relation reduce
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --timeout 600 -x > gpurun_out/t_all4.log 2>&1; echo "exit $? all gpu tests"; tail -4 gpurun_out/t_all4.log
for i in 1 2; do
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --workload wikikg2-transe-l1-d256-bf16 > gpurun_out/bench_wiki5_$i.json 2> gpurun_out/bench_wiki5.err; echo "bench wiki exit $?"; tail -3 gpurun_out/bench_wiki5.err; cut -c1-420 gpurun_out/bench_wiki5_$i.json
done
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_cfg2c.json 2> gpurun_out/bench_cfg2c.err; echo "bench cfg2 exit $?"; tail -3 gpurun_out/bench_cfg2c.err; cat gpurun_out/bench_cfg2c.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_wiki5.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --workload wikikg2-transe-l1-d256-bf16 > gpurun_out/ncu_wiki5.log 2>&1
echo "ncu launches exit $?"
