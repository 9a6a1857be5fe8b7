cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${1:-8}
WL=${2:-biokg-distmult-d256-fp32}
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --workload $WL > gpurun_out/bench_${WL}_n$N.json 2> gpurun_out/bench_${WL}_n$N.err; echo "bench exit $?"; tail -3 gpurun_out/bench_${WL}_n$N.err; cat gpurun_out/bench_${WL}_n$N.json
