set -x
mkdir -p gpurun_out
KB="python benchmarks/kernel_bench.py --reps 1"
prof() { local name=$1 rx=$2 sk=$3 ct=$4; shift 4
  $KB "$@" > gpurun_out/plain_$name.log 2>&1 && \
  timeout 200 ncu --set full --clock-control none -k "regex:$rx" -s $sk -c $ct -f -o gpurun_out/r1e_$name $KB "$@" > gpurun_out/ncu_$name.log 2>&1
  echo "ncu $name rc=$?"; }
prof resample resample_ 3 1 --only resample
prof mirror mirror_kernel 3 2 --only mirror
prof hausdorff 'mask_edges|edt_pass' 4 4 --only hausdorff
prof loss dice_ce 3 1 --only loss
ls -la gpurun_out/r1e_*
