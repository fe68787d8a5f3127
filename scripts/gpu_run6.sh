set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_e.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_e.log
python benchmarks/kernel_bench.py --only resample,intensity > gpurun_out/kb_f.log 2>&1; cat gpurun_out/kb_f.log
python benchmarks/kernel_bench.py --shape brats --only resample,intensity > gpurun_out/kb_f_brats.log 2>&1; cat gpurun_out/kb_f_brats.log
