mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x -k "resample or opt_in" 2>&1 | tail -2
for pf in 0 2 4 8; do echo "== PF=$pf"; MSS_RESAMPLE_PF=$pf python benchmarks/kernel_bench.py --only resample --reps 20 2>&1 | grep -i resample; done
echo "== CB=8 PF=4"; MSS_RESAMPLE_CB=8 python benchmarks/kernel_bench.py --only resample --reps 20 2>&1 | grep -i resample
MSS_RESAMPLE_CB=8 python -m pytest tests -m gpu -q -x -k "resample" 2>&1 | tail -1
MSS_RESAMPLE_CB=4 MSS_RESAMPLE_PF=1 python -m pytest tests -m gpu -q -x -k "resample" 2>&1 | tail -1
