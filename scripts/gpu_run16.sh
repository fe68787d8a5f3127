set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_n.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_n.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multigpu_parity.py > gpurun_out/mg_parity_n2b.json 2> gpurun_out/mg_parity_n2b.err; echo "parity rc=$?"; cat gpurun_out/mg_parity_n2b.json; grep -i "error\|Traceback" gpurun_out/mg_parity_n2b.err | head -5
for h in nccl p2p; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload wholebody --steps 1 --warmup 1 --halo $h > gpurun_out/wholebody_n2_$h.json 2> gpurun_out/wholebody_n2_$h.err; echo "wb $h rc=$?"; python -c "
import json
d=json.loads(open('gpurun_out/wholebody_n2_$h.json').read().strip().splitlines()[-1]); print('$h', d['ms_per_step'], d['config']['partition'], d['breakdown_ms_per_step_by_rank']['halo'])"; grep -i "error\|Traceback" gpurun_out/wholebody_n2_$h.err | head -3
done
