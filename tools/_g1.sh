timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_col2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_col2.log
tail -5 gpurun_out/pytest_col2.log
for r in 48 56 64 72; do for kb in 0 48; do
  MASKLAB_B200_LIB=$PWD/variants/lib_r$r.so MLP_ROI_WINDOW_KB=$kb timeout 300 python bench.py --steps 100 --warmup 5 --streams 1 --no-e2e --no-cpu-baseline 2>gpurun_out/roi_r${r}_kb${kb}.err | python tools/stage_times.py r${r}_kb$kb | tee -a gpurun_out/roi_col_sweep2.txt
done; done
for kb in 0 48; do
  MLP_ROI_WINDOW_KB=$kb timeout 300 python bench.py --workload stress --steps 30 --warmup 3 --streams 1 --no-e2e --no-cpu-baseline 2>gpurun_out/roi_s_kb${kb}.err | python tools/stage_times.py stress_kb$kb | tee -a gpurun_out/roi_col_sweep2.txt
done
