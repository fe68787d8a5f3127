set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_p.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_p.log
python benchmarks/kernel_bench.py --only hausdorff,loss > gpurun_out/kb_p.log 2>&1; cat gpurun_out/kb_p.log
