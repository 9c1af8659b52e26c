#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/r2_cell_sweep.jsonl
: > $O
for g in 148 296 444 592 888 1184; do
  SWEEP_ITERS=512 SWEEP_ROUND=8 python tools/dsgd_stability_map.py nfcell8 $g none:0,rows:0.5 2>/dev/null >> $O
done
SWEEP_ITERS=512 SWEEP_ROUND=16 python tools/dsgd_stability_map.py nfcell8 444,888 none:0 2>/dev/null | sed 's/"workload": "nfcell8"/"workload": "nfcell8_round16"/' >> $O
for g in 296 592 888 1184; do
  SWEEP_ITERS=512 SWEEP_ROUND=16 python tools/dsgd_stability_map.py nfcell4 $g none:0 2>/dev/null >> $O
done
cat $O
