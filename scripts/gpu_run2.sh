set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_b.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_b.log
python benchmarks/kernel_bench.py > gpurun_out/kb_b.log 2>&1; echo "kb rc=$?"; cat gpurun_out/kb_b.log
python benchmarks/kernel_bench.py --shape brats --only vote,dice,finalize,extract > gpurun_out/kb_b_brats.log 2>&1; cat gpurun_out/kb_b_brats.log
KR='regex:accumulate_kernel|finalize_kernel|vote_|dice_kernel|extract_'
for k in finalize vote dice extract; do
  python benchmarks/kernel_bench.py --only $k --reps 1 > gpurun_out/plain_$k.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k "$KR" -s 3 -c 2 -f -o gpurun_out/r1b_$k python benchmarks/kernel_bench.py --only $k --reps 1 > gpurun_out/ncu_$k.log 2>&1
  echo "ncu $k rc=$?"
done
python bench.py --steps 2 --warmup 3 > gpurun_out/bench_b.json 2> gpurun_out/bench_b.err; echo "bench rc=$?"; cat gpurun_out/bench_b.json
