timeout 900 python -m pytest tests/test_gpu_draw.py -x -q > gpurun_out/pytest_draw.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_draw.log
tail -30 gpurun_out/pytest_draw.log
