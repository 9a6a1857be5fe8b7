set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm.py -x -q -m gpu --timeout 300 > gpurun_out/test_gpu_gemm.log 2>&1
echo "exit $? gemm tests"; tail -40 gpurun_out/test_gpu_gemm.log
timeout 300 python scripts/gemm_bench.py > gpurun_out/gemm_bench.log 2>&1; echo "exit $? gemm bench"; tail -20 gpurun_out/gemm_bench.log
