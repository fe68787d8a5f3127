set -x
mkdir -p gpurun_out
nvidia-smi -L
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multigpu_parity.py > gpurun_out/mg_parity_n2.json 2> gpurun_out/mg_parity_n2.err; echo "parity rc=$?"; cat gpurun_out/mg_parity_n2.json; tail -5 gpurun_out/mg_parity_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload wholebody --steps 1 --warmup 1 > gpurun_out/wholebody_n2.json 2> gpurun_out/wholebody_n2.err; echo "wb rc=$?"; cat gpurun_out/wholebody_n2.json; tail -5 gpurun_out/wholebody_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench2 rc=$?"; cat gpurun_out/bench_n2.json; tail -5 gpurun_out/bench_n2.err
python benchmarks/kernel_bench.py --only vote,dice > gpurun_out/kb_c.log 2>&1; cat gpurun_out/kb_c.log
python benchmarks/kernel_bench.py --shape wholebody --only vote,dice > gpurun_out/kb_c_wb.log 2>&1; cat gpurun_out/kb_c_wb.log
python benchmarks/kernel_bench.py --shape brats --only extract > gpurun_out/kb_c_brats.log 2>&1; cat gpurun_out/kb_c_brats.log
