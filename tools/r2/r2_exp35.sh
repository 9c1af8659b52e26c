#!/bin/bash
# Round-2 GPU call 35 (one B200): the final tree -- full GPU suite, smoke, launch list of the bench command, N = 1 bench line.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
O=gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q --durations=8 ) > $O/r2e35_gpu_tests.log 2>&1
echo "pytest rc=$?"; tail -14 $O/r2e35_gpu_tests.log
timeout 300 python __graft_entry__.py --smoke > $O/r2e35_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/r2e35_smoke.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_final_launches_iters100.csv \
    python bench.py --steps 2 --warmup 1 --iters-per-step 100 --no-cpu-baseline --no-variants > $O/r2_final_launches_bench.json 2> $O/r2_final_launches_bench.log
echo "ncu launches rc=$?"
timeout 600 python bench.py > $O/r2_final_bench_n1.json 2> $O/r2_final_bench_n1.log; echo "bench rc=$?"
cat $O/r2_final_bench_n1.json
