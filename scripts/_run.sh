mkdir -p gpurun_out
for tpc in 1 2 8 16; do echo "== TPC $tpc"; MSS_ROWS_TPC=$tpc timeout 200 python benchmarks/kernel_bench.py --only accumulate --reps 7 2>&1 | grep -i "fused->labels"; done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:accumulate_rows -c 1 -f -o gpurun_out/r2_ncu_acc_rows_k14 python benchmarks/kernel_bench.py --only accumulate --reps 1 > gpurun_out/ncu_acc_rows14.log 2>&1; echo "ncu rc=$?"
