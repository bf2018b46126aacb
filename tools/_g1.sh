timeout 900 python bench.py --no-cpu-baseline --steps 50 > gpurun_out/bench_serv.json 2> gpurun_out/bench_serv.err; echo "rc=$?"; tail -3 gpurun_out/bench_serv.err
