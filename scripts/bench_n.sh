N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.log; echo "rc=$?"; tail -c 3000 gpurun_out/bench_n$N.json; tail -5 gpurun_out/bench_n$N.log
