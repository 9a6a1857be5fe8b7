cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu.log 2>&1
echo "ncu launches exit $?"
ncu --set full --clock-control none --import-source on -k regex:loss_kernel -c 2 -f -o gpurun_out/prof_loss python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu3.log 2>&1
echo "ncu loss exit $?"
