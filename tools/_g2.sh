timeout 600 ncu --set full --clock-control none --import-source on -k regex:box_summary --launch-skip 3 -c 1 -f -o gpurun_out/box_summary python bench.py --steps 2 --warmup 3 --streams 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_box.log 2>&1
tail -2 gpurun_out/ncu_box.log
