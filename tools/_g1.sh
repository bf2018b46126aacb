for cfg in "48 32" "48 36" "56 40" "56 48" "56 0"; do set -- $cfg; r=$1; kb=$2
  MASKLAB_B200_LIB=$PWD/variants/lib_r$r.so MLP_ROI_WINDOW_KB=$kb timeout 300 python bench.py --steps 100 --warmup 5 --streams 1 --no-e2e --no-cpu-baseline 2>gpurun_out/roi_r${r}_kb${kb}.err | python tools/stage_times.py r${r}_kb$kb | tee -a gpurun_out/roi_col_sweep3.txt
  MASKLAB_B200_LIB=$PWD/variants/lib_r$r.so MLP_ROI_WINDOW_KB=$kb timeout 300 python bench.py --workload stress --steps 30 --warmup 3 --streams 1 --no-e2e --no-cpu-baseline 2>gpurun_out/roi_s_kb${kb}.err | python tools/stage_times.py stress_r${r}_kb$kb | tee -a gpurun_out/roi_col_sweep3.txt
done
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "ref rc=$?"
