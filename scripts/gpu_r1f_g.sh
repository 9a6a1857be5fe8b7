# round-1f call G: TripleRE + full GPU suite + smoke + default bench line
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --timeout 600 > gpurun_out/t_all5.log 2>&1; echo "exit $? all gpu tests"; tail -12 gpurun_out/t_all5.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke2.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke2.log
timeout 400 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench exit $?"; tail -2 gpurun_out/bench_default.err; cat gpurun_out/bench_default.json
