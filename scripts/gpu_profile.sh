set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu.log 2>&1
echo "ncu launches exit $?"; tail -3 gpurun_out/ncu.log
python scripts/gemm_bench.py --one > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_tc -c 4 -f -o gpurun_out/prof_gemm python scripts/gemm_bench.py --one > gpurun_out/ncu2.log 2>&1
echo "ncu full exit $?"; tail -3 gpurun_out/ncu2.log; ls -la gpurun_out
