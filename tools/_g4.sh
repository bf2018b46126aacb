for n in 4 8; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --gpus $n --steps 200 --warmup 5 --no-summary --no-cpu-baseline > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err; echo "n$n rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n bench.py --gpus $n --workload cfg5 --frames 1024 --no-e2e --no-cpu-baseline --no-summary > gpurun_out/bench_stream_n$n.json 2> gpurun_out/bench_stream_n$n.err; echo "stream n$n rc=$?"
done
