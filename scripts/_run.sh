mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x -k "hausdorff or edt or hd95 or mask_edges" 2>&1 | tail -2
python benchmarks/kernel_bench.py --only hausdorff,hausdorff_api --reps 8 2>&1 | grep -i "hausdorff"
python benchmarks/kernel_bench.py --shape brats --only hausdorff,hausdorff_api --reps 8 2>&1 | grep -i "hausdorff"
