set -x
mkdir -p gpurun_out
python __graft_entry__.py smoke > gpurun_out/smoke_a.log 2>&1; echo "smoke rc=$?"
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_a.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_a.log
python benchmarks/kernel_bench.py > gpurun_out/kb_a.log 2>&1; echo "kb rc=$?"; cat gpurun_out/kb_a.log
KR='regex:accumulate_kernel|finalize_kernel|vote_kernel|dice_kernel'
for k in accumulate finalize vote dice; do
  python benchmarks/kernel_bench.py --only $k --reps 1 > gpurun_out/plain_$k.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k "$KR" -s 3 -c 2 -f -o gpurun_out/r1a_$k python benchmarks/kernel_bench.py --only $k --reps 1 > gpurun_out/ncu_$k.log 2>&1
  echo "ncu $k rc=$?"
done
ls -la gpurun_out
