#!/bin/bash
# Round-2 GPU call 1 (one B200): L2 row micro-benchmarks, item-bias layout A/B on the single-GPU kernel
# and on the DSGD block stand-ins, per-L2-slice counters, and the GPU test suite on the new layout.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/r2e1_smi.txt 2>&1
( time ./tools/micro/l2_rows > $O/r2_l2_rows_micro.jsonl ) 2> $O/r2_l2_rows_micro.err
echo "micro rc=$?"
# single-GPU kernel: bias stride A/B (+ occupancy / look-ahead on the padded layout)
SWEEP_ITERS=200 python tools/sweep_sgd.py "ibs1;CU2B_IB_STRIDE=1" "ibs8;CU2B_IB_STRIDE=8" "ibs32;CU2B_IB_STRIDE=32" "ibs64;CU2B_IB_STRIDE=64" \
    "ibs64_occ8;CU2B_IB_STRIDE=64;CU2B_TUNE_MINB=8" "ibs64_pf;CU2B_IB_STRIDE=64;CU2B_TUNE_PF=1" "ibs64_occ8_pf;CU2B_IB_STRIDE=64;CU2B_TUNE_MINB=8;CU2B_TUNE_PF=1" \
    > $O/r2_sweep_ibs.jsonl 2> $O/r2_sweep_ibs.err
echo "sweep_ibs rc=$?"
# DSGD block stand-ins: default / bias-thinned / row+bias-thinned kernel at forced grids, dense vs padded bias
for wl in nfblock8 nfblock4; do
  for ibs in 1 64; do
    SWEEP_WORKLOAD=$wl CU2B_IB_STRIDE=$ibs python tools/dsgd_thin_sweep.py 296 592 1184 >> $O/r2_dsgd_thin_sweep.jsonl 2>> $O/r2_dsgd_thin_sweep.err
    echo "thin sweep $wl ibs=$ibs rc=$?"
  done
done
# per-slice counters of one steady round of the single-GPU kernel, dense vs padded bias
for ibs in 1 64; do
  CU2B_IB_STRIDE=$ibs SWEEP_ITERS=64 timeout 600 ncu --metrics lts__d_atomic_input_cycles_active.sum,lts__t_sectors_srcunit_tex_op_red.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_requests_srcunit_tex.sum,lts__t_sectors.sum,gpu__time_duration.sum \
      --print-metric-instances values --clock-control none -k regex:mf_sgd_user_rounds -s 3 -c 1 \
      --log-file $O/r2_slices_ibs$ibs.txt python tools/sweep_sgd.py fused > /dev/null 2> $O/r2_slices_ibs$ibs.err
  echo "ncu slices ibs=$ibs rc=$?"
done
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2e1_gpu_tests.log 2>&1
echo "pytest rc=$?"
tail -3 $O/r2e1_gpu_tests.log
