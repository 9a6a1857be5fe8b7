# round-1f call B: all GPU tests with the pipelined pair_fwd / FMA-pipe L1 backward / per-triple row kernel,
# cfg-4 and cfg-5 bench lines
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --timeout 600 -x > gpurun_out/t_all.log 2>&1; echo "exit $? all gpu tests"; tail -15 gpurun_out/t_all.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --workload wikikg2-transe-l1-d256-bf16 > gpurun_out/bench_wiki2.json 2> gpurun_out/bench_wiki2.err; echo "bench wiki exit $?"; tail -3 gpurun_out/bench_wiki2.err; cat gpurun_out/bench_wiki2.json
for w in wikikg2-rotate-d512-scoremoving wikikg2-pairre-d512-scoremoving; do
  timeout 600 python bench.py --steps 10 --warmup 3 --workload $w > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "bench $w exit $?"; tail -3 gpurun_out/bench_$w.err; cat gpurun_out/bench_$w.json
done
BESS_PERTRIPLE_V1=1 timeout 600 python bench.py --steps 10 --warmup 3 --workload wikikg2-rotate-d512-scoremoving > gpurun_out/bench_sm_rotate_v1.json 2>/dev/null; cat gpurun_out/bench_sm_rotate_v1.json
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_cfg2b.json 2> gpurun_out/bench_cfg2b.err; echo "bench cfg2 exit $?"; cat gpurun_out/bench_cfg2b.json
