mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_stitch.py -m gpu -q -x -k "rows_kernel or fused_labels" 2>&1 | tail -3
timeout 300 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -x -k "cfg4" 2>&1 | tail -2
for tpc in 1 2 4 8; do echo "== TPC $tpc"; MSS_ROWS_TPC=$tpc timeout 120 python benchmarks/kernel_bench.py --shape brats --only accumulate --reps 10 2>&1 | grep -i "fused->labels"; done
timeout 120 python benchmarks/kernel_bench.py --shape brats_w156 --only accumulate --reps 10 2>&1 | grep -i "fused->labels"
