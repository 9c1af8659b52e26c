#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
O=gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --durations=8 ) > $O/r2e17_gpu_tests.log 2>&1
echo "pytest rc=$?"; tail -22 $O/r2e17_gpu_tests.log
python __graft_entry__.py --smoke > $O/r2e17_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/r2e17_smoke.log
SWEEP_ITERS=64 SWEEP_ROUND=8 timeout 600 ncu --set full --clock-control none --import-source on -k regex:mf_sgd_user_runs -s 6 -c 1 -f -o $O/r2_cell \
    python tools/dsgd_stability_map.py nfcell8 444 none:0 > $O/r2_ncu_cell.log 2>&1
echo "ncu cell rc=$?"
python bench.py --steps 4 --warmup 3 > $O/bench_r2_n1_final.json 2> $O/bench_r2_n1_final.log
echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_r2_reference.json 2> $O/bench_r2_reference.log
echo "reference arm rc=$?"
