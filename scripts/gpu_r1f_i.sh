# round-1f call I: norm-expanded L2 on the tensor cores
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --timeout 600 > gpurun_out/t_all7.log 2>&1; echo "exit $? all gpu tests"; grep -E "passed|failed|Error" gpurun_out/t_all7.log | tail -5; grep -E "^FAILED|^ERROR" gpurun_out/t_all7.log | head -20
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --workload biokg-transe-l2-d128-fp32 > gpurun_out/bench_l2b.json 2> gpurun_out/bench_l2b.err; echo "bench l2 exit $?"; tail -2 gpurun_out/bench_l2b.err; cut -c1-300 gpurun_out/bench_l2b.json
