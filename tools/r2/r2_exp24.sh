#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
for c in 1 2 4 8; do
  for g in 444 888 1184; do
    CU2B_DSGD_CLAIM=$c SWEEP_ITERS=512 SWEEP_ROUND=8 python tools/dsgd_stability_map.py nfcell8 $g none:0 2>/dev/null | sed "s/^/claim $c /"
  done
done
for c in 4 8; do
  CU2B_DSGD_CLAIM=$c SWEEP_ITERS=512 SWEEP_ROUND=16 python tools/dsgd_stability_map.py nfcell4 888 none:0 2>/dev/null | sed "s/^/claim $c /"
done
