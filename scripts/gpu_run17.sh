set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$TR --nproc-per-node 8 --master-port 29511 tests/multigpu_parity.py > gpurun_out/mg_parity_n8b.json 2> gpurun_out/mg_parity_n8b.err; echo "parity rc=$?"; cat gpurun_out/mg_parity_n8b.json; grep -i "error\|Traceback" gpurun_out/mg_parity_n8b.err | head -5
for h in p2p nccl; do
$TR --nproc-per-node 8 --master-port 29512 bench.py --gpus 8 --workload wholebody --steps 2 --warmup 1 --halo $h > gpurun_out/wholebody_n8_$h.json 2> gpurun_out/wholebody_n8_$h.err; echo "wb8 $h rc=$?"; cat gpurun_out/wholebody_n8_$h.json; grep -i "error\|Traceback" gpurun_out/wholebody_n8_$h.err | head -3
done
$TR --nproc-per-node 4 --master-port 29513 bench.py --gpus 4 --workload wholebody --steps 1 --warmup 1 --halo p2p > gpurun_out/wholebody_n4_p2p.json 2> gpurun_out/wholebody_n4_p2p.err; echo "wb4 rc=$?"; cat gpurun_out/wholebody_n4_p2p.json
$TR --nproc-per-node 8 --master-port 29514 bench.py --gpus 8 --workload brats --steps 2 --warmup 1 > gpurun_out/bench_brats_n8.json 2> gpurun_out/bench_brats_n8.err; echo "brats8 rc=$?"; cat gpurun_out/bench_brats_n8.json; grep -i "error\|Traceback" gpurun_out/bench_brats_n8.err | head -3
$TR --nproc-per-node 8 --master-port 29515 bench.py --gpus 8 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_n8b.json 2> gpurun_out/bench_n8b.err; echo "bench8 rc=$?"; cat gpurun_out/bench_n8b.json
