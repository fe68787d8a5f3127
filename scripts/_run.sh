mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_stitch.py -m gpu -q -x 2>&1 | tail -2
timeout 300 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -x -k "cfg4" 2>&1 | tail -1
for st in 2 3 4; do echo "== stages $st"; MSS_ROWS_STAGES=$st timeout 120 python benchmarks/kernel_bench.py --shape brats --only accumulate --reps 10 2>&1 | grep -i "fused"; done
MSS_ROWS_STAGES=2 timeout 300 python -m pytest tests/test_gpu_stitch.py -m gpu -q -x -k "rows_kernel" 2>&1 | tail -1
