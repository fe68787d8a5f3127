set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_d.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_d.log
python benchmarks/kernel_bench.py --only extract,resample,dice > gpurun_out/kb_e.log 2>&1; cat gpurun_out/kb_e.log
python benchmarks/kernel_bench.py --shape brats --only extract,finalize,resample > gpurun_out/kb_e_brats.log 2>&1; cat gpurun_out/kb_e_brats.log
python bench.py --workload wholebody --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/wholebody_n1.json 2> gpurun_out/wholebody_n1.err; echo "wb1 rc=$?"; cat gpurun_out/wholebody_n1.json; tail -3 gpurun_out/wholebody_n1.err
