cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
BESSKGE_GEMM_CLUSTER=22 timeout 300 python -m pytest tests/test_gpu_gemm.py -x -q -m gpu --timeout 120 > gpurun_out/test_gpu_gemm_2sm.log 2>&1
echo "exit $? gemm tests 2sm"; tail -25 gpurun_out/test_gpu_gemm_2sm.log
BESSKGE_GEMM_CLUSTER=22 timeout 200 python scripts/gemm_bench.py > gpurun_out/gemm_bench_2sm.log 2>&1; echo "exit $? gemm bench 2sm"; cat gpurun_out/gemm_bench_2sm.log | tail -12
