# round-1f call J: L2 tensor-core path gated by pass size; A/B at a large L2 shape
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --timeout 600 > gpurun_out/t_all8.log 2>&1; echo "exit $? all gpu tests"; grep -E "passed|failed" gpurun_out/t_all8.log | tail -3; grep -E "^FAILED|^ERROR" gpurun_out/t_all8.log | head
for tc in 1 0; do
BESS_L2_TC=$tc timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --workload biokg-transe-l2-d128-fp32 --shard-bs 16384 --negatives 2048 > gpurun_out/bench_l2big_tc$tc.json 2> gpurun_out/bench_l2big_tc$tc.err; echo "bench l2 big tc=$tc exit $?"; tail -2 gpurun_out/bench_l2big_tc$tc.err; cut -c1-260 gpurun_out/bench_l2big_tc$tc.json
done
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --workload biokg-transe-l2-d128-fp32 > gpurun_out/bench_l2c.json 2>/dev/null; cut -c1-260 gpurun_out/bench_l2c.json
