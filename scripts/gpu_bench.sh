set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
python bench.py --steps 10 --warmup 3 --workload wikikg2-transe-l1-d256-bf16 --no-cpu-baseline > gpurun_out/bench_wiki.json 2> gpurun_out/bench_wiki.err; echo "bench exit $?"; tail -3 gpurun_out/bench_wiki.err; cat gpurun_out/bench_wiki.json
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu.log
