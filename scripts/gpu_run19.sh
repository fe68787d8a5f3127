set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_o.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_o.log
python benchmarks/kernel_bench.py --only hausdorff > gpurun_out/kb_o.log 2>&1; cat gpurun_out/kb_o.log
