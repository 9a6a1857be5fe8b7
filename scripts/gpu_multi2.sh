cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${1:-2}
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 tests/dist_gpu_worker.py > gpurun_out/dist_parity_$N.log 2>&1; echo "dist parity exit $?"; grep -E "ok|==|Error|error" gpurun_out/dist_parity_$N.log | tail -12
bash scripts/gpu_multi_bench.sh $N
