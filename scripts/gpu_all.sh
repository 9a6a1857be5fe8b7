set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for f in tests/test_gpu_gemm.py tests/test_gpu_kernels.py tests/test_gpu_bess.py tests/test_gpu_fullsize.py; do
  timeout 900 python -m pytest $f -q -m gpu --timeout 600 > gpurun_out/$(basename $f .py).log 2>&1
  echo "exit $? for $f"; tail -15 gpurun_out/$(basename $f .py).log
done
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
