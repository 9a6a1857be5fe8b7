set -x
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for f in tests/test_gpu_kernels.py tests/test_gpu_bess.py; do
  timeout 900 python -m pytest $f -q -m gpu --timeout 600 > gpurun_out/$(basename $f .py).log 2>&1
  echo "exit $? for $f"; tail -30 gpurun_out/$(basename $f .py).log
done
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -5 gpurun_out/smoke.log
