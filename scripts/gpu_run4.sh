set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_c.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_c.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multigpu_parity.py > gpurun_out/mg_parity_n2.json 2> gpurun_out/mg_parity_n2.err; echo "parity rc=$?"; cat gpurun_out/mg_parity_n2.json; grep -v "^\s*$" gpurun_out/mg_parity_n2.err | grep -i "error\|Traceback" | head -5
python benchmarks/kernel_bench.py --only extract,dice,vote,resample > gpurun_out/kb_d.log 2>&1; cat gpurun_out/kb_d.log
python benchmarks/kernel_bench.py --shape brats --only extract,dice,vote,finalize,resample > gpurun_out/kb_d_brats.log 2>&1; cat gpurun_out/kb_d_brats.log
python benchmarks/kernel_bench.py --shape wholebody --only vote,dice > gpurun_out/kb_d_wb.log 2>&1; cat gpurun_out/kb_d_wb.log
