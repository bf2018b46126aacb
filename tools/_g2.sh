for kb in 0 48; do
MLP_ROI_WINDOW_KB=$kb timeout 600 ncu --set full --clock-control none --import-source on -k regex:roi_align --launch-skip 3 -c 1 -f -o gpurun_out/roi_col_kb$kb python bench.py --steps 3 --warmup 3 --streams 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_roi_kb$kb.log 2>&1
tail -3 gpurun_out/ncu_roi_kb$kb.log
done
