set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=2
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 tests/dist_gpu_worker.py > gpurun_out/dist_parity_r1g_$N.log 2>&1; echo "dist parity exit $?"; grep -c "ok\|== local" gpurun_out/dist_parity_r1g_$N.log; tail -3 gpurun_out/dist_parity_r1g_$N.log
for WL in biokg-distmult-d256-fp32 wikikg2-transe-l1-d256-bf16; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --workload $WL > gpurun_out/bench_r1g_${WL}_n$N.json 2> gpurun_out/bench_r1g_${WL}_n$N.err; echo "bench $WL exit $?"; tail -2 gpurun_out/bench_r1g_${WL}_n$N.err; cut -c1-330 gpurun_out/bench_r1g_${WL}_n$N.json
done
