timeout 600 python bench.py --steps 2 --warmup 3 --streams 1 --no-e2e --no-cpu-baseline > gpurun_out/plain_r01b.log 2>&1 || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01b.csv python bench.py --steps 2 --warmup 3 --streams 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launch_r01b.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --launch-skip 24 -c 6 -f -o gpurun_out/prof_r01b_full python bench.py --steps 2 --warmup 3 --streams 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full_r01b.log 2>&1
tail -2 gpurun_out/ncu_full_r01b.log
