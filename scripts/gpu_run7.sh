set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_f.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_f.log
python benchmarks/kernel_bench.py > gpurun_out/kb_g.log 2>&1; cat gpurun_out/kb_g.log
python benchmarks/kernel_bench.py --shape brats > gpurun_out/kb_g_brats.log 2>&1; cat gpurun_out/kb_g_brats.log
