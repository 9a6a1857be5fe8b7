cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${1:-8}
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench exit $?"; tail -5 gpurun_out/bench_n$N.err; cat gpurun_out/bench_n$N.json
timeout 180 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 tests/dist_gpu_worker.py > gpurun_out/dist_parity_$N.log 2>&1; echo "dist parity exit $?"; tail -4 gpurun_out/dist_parity_$N.log
