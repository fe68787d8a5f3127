mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x -k "hausdorff or edt or hd95 or mask_edges" 2>&1 | tail -3
python benchmarks/kernel_bench.py --only hausdorff,hausdorff_api --reps 6 2>&1 | grep -i "hausdorff"
python benchmarks/kernel_bench.py --shape brats --only hausdorff,hausdorff_api --reps 6 2>&1 | grep -i "hausdorff"
python benchmarks/kernel_bench.py --only resample --reps 3 > /dev/null 2>&1 && timeout 300 ncu --set full --clock-control none --import-source on -k regex:resample_stream -c 2 -f -o gpurun_out/r2_ncu_resample python benchmarks/kernel_bench.py --only resample --reps 1 > gpurun_out/ncu_resample.log 2>&1; echo "ncu rc=$?"
