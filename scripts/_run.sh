mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -x -k "dice" 2>&1 | tail -1
python benchmarks/kernel_bench.py --only dice --reps 15 2>&1 | grep -i dice
python benchmarks/kernel_bench.py --shape brats --only dice --reps 15 2>&1 | grep -i dice
python bench.py --no-legs --no-cpu-baseline --steps 1 --warmup 1 > /dev/null 2>&1 && timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:accumulate|extract|finalize|importance|gaussian" --csv --log-file gpurun_out/launches_mss.csv python bench.py --no-legs --no-cpu-baseline --steps 1 --warmup 1 > gpurun_out/ncu_launches_mss.log 2>&1; echo "launches rc=$?"
