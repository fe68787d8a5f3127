set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_m.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_m.log
python benchmarks/kernel_bench.py --only mirror > gpurun_out/kb_m.log 2>&1; cat gpurun_out/kb_m.log
