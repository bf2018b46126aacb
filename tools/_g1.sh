timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_all.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_all.log
tail -4 gpurun_out/pytest_all.log
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_default.err
timeout 600 python bench.py --workload stress --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/bench_stress.json 2> gpurun_out/bench_stress.err; echo "stress rc=$?"; tail -2 gpurun_out/bench_stress.err
timeout 600 python bench.py --workload cfg5 --steps 50 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg5.json 2> gpurun_out/bench_cfg5.err; echo "cfg5 rc=$?"; tail -2 gpurun_out/bench_cfg5.err
