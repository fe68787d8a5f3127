set -x
mkdir -p gpurun_out
nvidia-smi -L | wc -l
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_g.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_g.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tests/multigpu_parity.py > gpurun_out/mg_parity_n8.json 2> gpurun_out/mg_parity_n8.err; echo "parity rc=$?"; cat gpurun_out/mg_parity_n8.json; grep -i "error\|Traceback" gpurun_out/mg_parity_n8.err | head -5
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --workload wholebody --steps 2 --warmup 1 > gpurun_out/wholebody_n8.json 2> gpurun_out/wholebody_n8.err; echo "wb8 rc=$?"; cat gpurun_out/wholebody_n8.json; grep -i "error\|Traceback" gpurun_out/wholebody_n8.err | head -5
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --workload wholebody --steps 1 --warmup 1 --block-dims 1x1x8 > gpurun_out/wholebody_n8_slab.json 2> gpurun_out/wholebody_n8_slab.err; echo "wb8slab rc=$?"; cat gpurun_out/wholebody_n8_slab.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 4 --workload wholebody --steps 1 --warmup 1 > gpurun_out/wholebody_n4.json 2> gpurun_out/wholebody_n4.err; echo "wb4 rc=$?"; cat gpurun_out/wholebody_n4.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus 8 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; echo "bench8 rc=$?"; cat gpurun_out/bench_n8.json
