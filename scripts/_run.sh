mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x -k "hausdorff or edt or hd95 or mask_edges" 2>&1 | tail -2
python benchmarks/kernel_bench.py --only hausdorff --reps 10 2>&1 | grep -i "mask_edges"
for n in 1 2 4 8; do echo "== streams $n"; MSS_HD_STREAMS=$n python benchmarks/kernel_bench.py --only hausdorff_api --reps 6 2>&1 | grep -i "hausdorff"; done
MSS_HD_STREAMS=8 python benchmarks/kernel_bench.py --shape brats --only hausdorff_api --reps 6 2>&1 | grep -i "hausdorff"
