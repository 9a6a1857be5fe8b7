set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=4
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_r1g_default_n$N.json 2> gpurun_out/bench_r1g_default_n$N.err; echo "bench exit $?"; tail -2 gpurun_out/bench_r1g_default_n$N.err; cut -c1-330 gpurun_out/bench_r1g_default_n$N.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29553 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/bench_r1g_ref_n$N.json 2> gpurun_out/bench_r1g_ref_n$N.err; echo "ref exit $?"; cut -c1-200 gpurun_out/bench_r1g_ref_n$N.json
