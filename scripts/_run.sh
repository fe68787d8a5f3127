mkdir -p gpurun_out
for rpt in 1 2 4; do for st in 4 8; do echo "== rpt $rpt stages $st"; MSS_ROWS_RPT=$rpt MSS_ROWS_STAGES=$st timeout 120 python benchmarks/kernel_bench.py --shape brats --only accumulate --reps 10 2>&1 | grep -i "fused->labels"; done; done
for tpc in 1 4; do echo "== rpt 2 tpc $tpc";  MSS_ROWS_TPC=$tpc timeout 120 python benchmarks/kernel_bench.py --shape brats --only accumulate --reps 10 2>&1 | grep -i "fused->labels"; done
timeout 300 python -m pytest tests/test_gpu_stitch.py -m gpu -q -x -k "rows_kernel or fused_labels" 2>&1 | tail -1
MSS_ROWS_RPT=4 timeout 300 python -m pytest tests/test_gpu_stitch.py -m gpu -q -x -k "rows_kernel" 2>&1 | tail -1
timeout 300 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -x -k "cfg4" 2>&1 | tail -1
