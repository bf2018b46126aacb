timeout 900 python -m pytest tests/test_gpu_summary.py -x -q > gpurun_out/pytest_summary.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_summary.log
tail -40 gpurun_out/pytest_summary.log
