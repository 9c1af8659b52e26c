#!/bin/bash
# Round-2 GPU call 3 (one B200): GPU tests on the new layout / defaults, item placement A/B on the single-GPU
# kernel, DSGD block stand-ins with the partition's row placement.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2e3_gpu_tests.log 2>&1
echo "pytest rc=$?"; tail -5 $O/r2e3_gpu_tests.log
SWEEP_ITERS=200 python tools/sweep_sgd.py "placement0_ibs1;CU2B_PLACEMENT=0;CU2B_IB_STRIDE=1" "placement0_ibs64;CU2B_PLACEMENT=0" \
    "placement1_ibs1;CU2B_IB_STRIDE=1" "placement1_ibs8;CU2B_IB_STRIDE=8" "placement1_ibs64" "placement1_ibs64_occ8;CU2B_TUNE_MINB=8" \
    "placement1_ibs64_pf;CU2B_TUNE_PF=1" "placement1_ibs64_occ4;CU2B_TUNE_MINB=4" > $O/r2_sweep_placement.jsonl 2> $O/r2_sweep_placement.err
echo "sweep rc=$?"; cat $O/r2_sweep_placement.jsonl
SWEEP_WORKLOAD=ml20m SWEEP_K=64 SWEEP_ITERS=200 python tools/sweep_sgd.py "ml20m_placement0;CU2B_PLACEMENT=0;CU2B_IB_STRIDE=1" "ml20m_placement1" >> $O/r2_sweep_placement.jsonl 2>> $O/r2_sweep_placement.err
M="none:0,bias:0.25,rows:0.5"
for wl in nfblock8 nfblock4 nfblock2; do
  python tools/dsgd_stability_map.py $wl 592,1184 $M >> $O/r2_stability_placed.jsonl 2>> $O/r2_stability_placed.err
  echo "map $wl rc=$?"
done
cat $O/r2_stability_placed.jsonl
