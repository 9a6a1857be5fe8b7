set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm.py -q -m gpu --timeout 300 > gpurun_out/test_gpu_gemm.log 2>&1
echo "exit $? gemm tests"; tail -5 gpurun_out/test_gpu_gemm.log
for cs in 2 1 4; do
BESSKGE_GEMM_CLUSTER=$cs timeout 300 python scripts/gemm_bench.py > gpurun_out/gemm_bench_cs$cs.log 2>&1; echo "exit $? gemm bench cs=$cs"; head -3 gpurun_out/gemm_bench_cs$cs.log
done
BESSKGE_GEMM_CLUSTER=4 timeout 600 python -m pytest tests/test_gpu_gemm.py -q -m gpu --timeout 300 2>&1 | tail -3
