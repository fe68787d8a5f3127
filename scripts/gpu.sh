#!/bin/bash
# One parameterised driver for the GPU box (replaces the numbered one-off scripts of round 1).
#   scripts/gpu.sh tests [pytest args]       -> gpurun_out/pytest_gpu.log
#   scripts/gpu.sh bench [bench.py args]     -> gpurun_out/bench.json / bench.log
#   scripts/gpu.sh kbench [kernel_bench args]-> gpurun_out/kbench.json
#   scripts/gpu.sh launches [bench.py args]  -> ncu launch list gpurun_out/launches.csv
#   scripts/gpu.sh ncu NAME REGEX SKIP COUNT -- <command>  -> gpurun_out/NAME.ncu-rep (ncu --set full)
set -x
mkdir -p gpurun_out
what=$1; shift
case "$what" in
  tests)    python -m pytest tests -m gpu -q -x "$@" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log ;;
  bench)    python bench.py "$@" > gpurun_out/bench.json 2> gpurun_out/bench.log; echo "bench rc=$?"; cat gpurun_out/bench.json ;;
  kbench)   python benchmarks/kernel_bench.py "$@" > gpurun_out/kbench.log 2>&1; echo "kbench rc=$?"; tail -40 gpurun_out/kbench.log ;;
  launches) # a step of the default workload is ~27 000 launches and ncu costs ~25 ms per profiled launch: profile a WINDOW
            # (LAUNCH_SKIP / LAUNCH_COUNT), e.g. the second half of the timed step incl. its accumulate launch
            python bench.py "$@" > /dev/null 2>&1 && timeout 1100 ncu --metrics gpu__time_duration.sum --clock-control none \
              -s ${LAUNCH_SKIP:-40504} -c ${LAUNCH_COUNT:-13503} --csv --log-file gpurun_out/launches.csv python bench.py "$@" \
              > gpurun_out/ncu_launches.log 2>&1; echo "launches rc=$?" ;;
  ncu)      name=$1; rx=$2; sk=$3; ct=$4; shift 5
            "$@" > gpurun_out/plain_$name.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on \
              -k "regex:$rx" -s $sk -c $ct -f -o gpurun_out/$name "$@" > gpurun_out/ncu_$name.log 2>&1; echo "ncu $name rc=$?" ;;
  *) echo "unknown: $what"; exit 2 ;;
esac
