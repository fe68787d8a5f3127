set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_i.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_i.log
python benchmarks/kernel_bench.py --shape brats --only accumulate > gpurun_out/kb_i_brats.log 2>&1; cat gpurun_out/kb_i_brats.log
python __graft_entry__.py smoke
# launch list of the bench command, bounded: skip the warm-up step, profile about one step's worth of launches
BC="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
$BC > gpurun_out/plain_bench.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 13000 -c 13500 --csv --log-file gpurun_out/launches_bench.csv $BC > gpurun_out/ncu_bench.log 2>&1
echo "ncu bench rc=$?"
python scripts/summarise_launches.py gpurun_out/launches_bench.csv --out gpurun_out/launches_bench_summary.md | tail -30
gzip -f gpurun_out/launches_bench.csv; ls -la gpurun_out/launches_bench*
