set -x
cd $GRAFT_REPO_ROOT
timeout 600 python bench.py > gpurun_out/bench_r02.json 2> gpurun_out/bench_r02.err; echo bench rc=$?
timeout 600 python bench.py --impl reference > gpurun_out/bench_r02_reference.json 2> gpurun_out/bench_r02_reference.err; echo ref rc=$?
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_bench_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_launch.log 2>&1; echo launches rc=$?
timeout 900 ncu --set full --clock-control none -o /tmp/r02_step -f python scripts/one_step.py 2 > gpurun_out/ncu_full.log 2>&1; echo full rc=$?
ncu -i /tmp/r02_step.ncu-rep --page raw --csv > gpurun_out/r02_step_raw.csv 2>/dev/null
ls -la gpurun_out | tail -8
