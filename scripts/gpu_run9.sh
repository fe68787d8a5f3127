set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_h.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_h.log
python benchmarks/kernel_bench.py --only resample,intensity > gpurun_out/kb_h.log 2>&1; cat gpurun_out/kb_h.log
python benchmarks/kernel_bench.py --shape brats --only resample,intensity,accumulate > gpurun_out/kb_h_brats.log 2>&1; cat gpurun_out/kb_h_brats.log
for sb in 8 16; do python bench.py --sw-batch $sb --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_sb$sb.json 2> gpurun_out/bench_sb$sb.err; echo "sb$sb rc=$?"; python -c "
import json,sys
d=json.loads(open('gpurun_out/bench_sb$sb.json').read().strip().splitlines()[-1]); print($sb, d['ms_per_step'], d['value'], d['breakdown_ms_per_step'])"; done
