set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_k.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_k.log
python benchmarks/kernel_bench.py --only accumulate,mirror > gpurun_out/kb_k.log 2>&1; cat gpurun_out/kb_k.log
python benchmarks/kernel_bench.py --shape brats --only accumulate > gpurun_out/kb_k_brats.log 2>&1; cat gpurun_out/kb_k_brats.log
python bench.py --workload brats --steps 1 --warmup 1 > gpurun_out/bench_brats_n1.json 2> gpurun_out/bench_brats_n1.err; echo "brats rc=$?"; cat gpurun_out/bench_brats_n1.json; tail -3 gpurun_out/bench_brats_n1.err
KB="python benchmarks/kernel_bench.py --reps 1"
$KB --only resample > gpurun_out/plain_resample.log 2>&1 && ncu --set full --clock-control none -k regex:resample_ -s 3 -c 1 -f -o gpurun_out/r1d_resample $KB --only resample > gpurun_out/ncu_resample.log 2>&1; echo "ncu rc=$?"
