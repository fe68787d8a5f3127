set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_final.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_final.log
python __graft_entry__.py smoke
python bench.py > gpurun_out/bench_final_n1.json 2> gpurun_out/bench_final_n1.err; echo "bench rc=$?"; cat gpurun_out/bench_final_n1.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_final_ref.json 2> gpurun_out/bench_final_ref.err; echo "ref rc=$?"; cat gpurun_out/bench_final_ref.json
python bench.py --workload wholebody --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/wholebody_n1_sb8.json 2> gpurun_out/wholebody_n1_sb8.err; echo "wb1 rc=$?"; cat gpurun_out/wholebody_n1_sb8.json
