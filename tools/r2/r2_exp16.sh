#!/bin/bash
# Round-2 GPU call 16 (one B200): full GPU suite (incl. the at-size parity tests), ncu captures of the final kernels,
# launch list of the bench command, full-size batched predict at k = 128 / 50 / 300.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
O=gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q --durations=12 ) > $O/r2e16_gpu_tests.log 2>&1
echo "pytest rc=$?"; tail -25 $O/r2e16_gpu_tests.log
SWEEP_ITERS=64 timeout 600 ncu --set full --clock-control none --import-source on -k regex:mf_sgd_user_rounds -s 3 -c 1 -f -o $O/r2_rounds \
    python tools/sweep_sgd.py fused > $O/r2_ncu_rounds.log 2>&1
echo "ncu rounds rc=$?"
SWEEP_ITERS=64 SWEEP_ROUND=8 timeout 600 ncu --set full --clock-control none --import-source on -k regex:mf_sgd_user_runs -s 20 -c 1 -f -o $O/r2_cell \
    python tools/dsgd_stability_map.py nfcell8 444 none:0 > $O/r2_ncu_cell.log 2>&1
echo "ncu cell rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches_iters100.csv \
    python bench.py --steps 2 --warmup 1 --iters-per-step 100 --no-cpu-baseline --no-variants > $O/r2_launches_bench.json 2> $O/r2_launches_bench.log
echo "ncu launches rc=$?"
for k in 128 50 300; do
  PREDICT_K=$k timeout 600 python tests/predict_full_size.py >> $O/r2_predict.jsonl 2>> $O/r2_predict.err
  echo "predict k=$k rc=$?"
done
cat $O/r2_predict.jsonl
