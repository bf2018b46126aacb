timeout 900 python -m pytest tests/test_gpu_tf_published_vectors.py -x -q > gpurun_out/pytest_tfv.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_tfv.log
tail -30 gpurun_out/pytest_tfv.log
