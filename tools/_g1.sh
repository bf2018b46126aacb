timeout 900 python -m pytest tests/test_gpu_draw.py -x -q > gpurun_out/pytest_draw.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_draw.log
tail -12 gpurun_out/pytest_draw.log
timeout 600 python bench.py --no-cpu-baseline --no-e2e --steps 30 > gpurun_out/bench_sum.json 2> gpurun_out/bench_sum.err; echo "cfg2 rc=$?"; tail -3 gpurun_out/bench_sum.err
