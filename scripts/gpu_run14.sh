set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_l.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_l.log
python benchmarks/kernel_bench.py --only accumulate,resample > gpurun_out/kb_l.log 2>&1; cat gpurun_out/kb_l.log
python benchmarks/kernel_bench.py --shape brats --only accumulate,resample > gpurun_out/kb_l_brats.log 2>&1; cat gpurun_out/kb_l_brats.log
