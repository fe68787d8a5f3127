set -x
mkdir -p gpurun_out
KB="python benchmarks/kernel_bench.py --reps 1"
prof() { # name, regex, skip, count, extra args...
  local name=$1 rx=$2 sk=$3 ct=$4; shift 4
  $KB "$@" > gpurun_out/plain_$name.log 2>&1 && \
  ncu --set full --clock-control none -k "regex:$rx" -s $sk -c $ct -f -o gpurun_out/r1c_$name $KB "$@" > gpurun_out/ncu_$name.log 2>&1
  echo "ncu $name rc=$?"
}
prof accumulate accumulate_kernel 3 2 --only accumulate
prof finalize finalize_ 3 1 --only finalize
prof vote vote_ 3 1 --only vote
prof dice dice_kernel 3 2 --only dice
prof extract_tma extract_tma 3 1 --only extract --extract-sizes 300
prof extract_shifted extract_shifted 3 1 --only extract --extract-sizes 300
prof halo halo_add 3 1 --only halo
prof intensity intensity_kernel 3 1 --only intensity
prof resample resample_ 3 1 --only resample
ls -la gpurun_out/*.ncu-rep
# launch list of the bench command (one warm-up step + one timed + one e2e step keep the list at ~40k launches)
BC="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
$BC > gpurun_out/plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench.csv $BC > gpurun_out/ncu_bench.log 2>&1
echo "ncu bench rc=$?"
python scripts/summarise_launches.py gpurun_out/launches_bench.csv --out gpurun_out/launches_bench_summary.md | tail -25
gzip -f gpurun_out/launches_bench.csv; ls -la gpurun_out/launches_bench*
