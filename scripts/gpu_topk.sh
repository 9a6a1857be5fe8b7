cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_bess.py -q -m gpu --timeout 300 -k "topk" > gpurun_out/test_topk.log 2>&1
echo "exit $?"; tail -60 gpurun_out/test_topk.log
