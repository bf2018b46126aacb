#!/bin/bash
# The round's evidence on ONE B200 (run under gpurun from the repo root): full GPU suite, debug-bounds build, the driver's
# bench command and the reference arm, the other workloads, ncu --set full of the six kernels of the path (+ planar tail,
# + the serving-tail kernels), and the ncu launch list of a short bench run.  Outputs land in gpurun_out/fin_*; the
# summaries kept under profiles/ are made from them with tools/ncu_summary.py.
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/fin_pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/fin_pytest_gpu.log
tail -4 gpurun_out/fin_pytest_gpu.log
bash tools/run_debug_bounds.sh > gpurun_out/fin_debug_bounds.txt 2>&1; tail -4 gpurun_out/fin_debug_bounds.txt
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/fin_bench_reference.json 2> gpurun_out/fin_bench_reference.err; echo ref rc=$?
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/fin_bench_n1.json 2> gpurun_out/fin_bench_n1.err; echo bench rc=$?
timeout 600 python bench.py --workload cfg1 --steps 200 --warmup 5 --no-summary > gpurun_out/fin_bench_cfg1.json 2> gpurun_out/fin_bench_cfg1.err; echo rc=$?
timeout 600 python bench.py --workload stress --steps 30 --warmup 3 --no-cpu-baseline --no-summary --no-e2e > gpurun_out/fin_bench_stress.json 2> gpurun_out/fin_bench_stress.err; echo rc=$?
timeout 600 python bench.py --workload cfg5 --steps 100 --warmup 5 --no-cpu-baseline --no-summary --no-e2e > gpurun_out/fin_bench_cfg5.json 2> gpurun_out/fin_bench_cfg5.err; echo rc=$?
python tools/prof_step.py cfg2 1 0 > gpurun_out/fin_prof_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -s 6 -c 6 -o gpurun_out/fin_prof python tools/prof_step.py cfg2 1 0 > gpurun_out/fin_ncu_full.log 2>&1
python tools/prof_serving.py cfg2 1 > gpurun_out/fin_prof_plain4.log 2>&1 && ncu --set full --clock-control none -k regex:"box_summary|box_plan|draw_tiles|road_rows|jpeg_dct|jpeg_enc|jpeg_place|jpeg_stuff|crack_cols|pack_tiles|box_lines" -s 9 -c 11 -o gpurun_out/fin_prof_serving python tools/prof_serving.py cfg2 1 > gpurun_out/fin_ncu_serving.log 2>&1
python tools/prof_step.py cfg2 1 0 1 > gpurun_out/fin_prof_plain3.log 2>&1 && ncu --set full --clock-control none -k regex:"tail_prep" -s 1 -c 1 -o gpurun_out/fin_prof_planar python tools/prof_step.py cfg2 1 0 1 > gpurun_out/fin_ncu_planar.log 2>&1
python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/fin_bench_short.json 2> gpurun_out/fin_bench_short.err && ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/fin_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/fin_ncu_launches.log 2>&1
ls -la gpurun_out/fin_*
