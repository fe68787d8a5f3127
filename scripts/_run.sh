mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_stitch.py -m gpu -q -x 2>&1 | tail -1
timeout 300 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -x -k "cfg4" 2>&1 | tail -1
timeout 120 python benchmarks/kernel_bench.py --shape brats --only accumulate --reps 12 2>&1 | grep -i "fused"
timeout 120 python benchmarks/kernel_bench.py --shape brats_w156 --only accumulate --reps 12 2>&1 | grep -i "fused"
timeout 120 python benchmarks/kernel_bench.py --shape btcv_k3 --only accumulate --reps 7 2>&1 | grep -i "fused"
