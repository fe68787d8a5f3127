set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_j.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_j.log
python benchmarks/kernel_bench.py --only resample,mirror > gpurun_out/kb_j.log 2>&1; cat gpurun_out/kb_j.log
python benchmarks/kernel_bench.py --shape brats --only resample > gpurun_out/kb_j_brats.log 2>&1; cat gpurun_out/kb_j_brats.log
# launch list of the bench command: one step = 50 backbone calls x 540 launches + 2 extract + 1 accumulate = 27003 launches;
# the window covers the second half of the timed step (its accumulate included) and the first extract of the next one
BC="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
$BC > gpurun_out/plain_bench.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 40504 -c 13503 --csv --log-file gpurun_out/launches_bench.csv $BC > gpurun_out/ncu_bench.log 2>&1
echo "ncu bench rc=$?"
python scripts/summarise_launches.py gpurun_out/launches_bench.csv --out gpurun_out/launches_bench_summary.md | tail -8
gzip -f gpurun_out/launches_bench.csv; ls -la gpurun_out/launches_bench*
