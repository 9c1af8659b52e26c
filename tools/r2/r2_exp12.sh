#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
for r in 8 16 32; do
  for g in 444 888; do
    SWEEP_ITERS=512 SWEEP_ROUND=$r python tools/dsgd_stability_map.py nfcell8 $g none:0,rows:0.5 2>/dev/null | sed "s/^/round $r /"
  done
done
for r in 16 32; do
  SWEEP_ITERS=512 SWEEP_ROUND=$r python tools/dsgd_stability_map.py nfcell4 888 none:0,rows:0.5 2>/dev/null | sed "s/^/round $r /"
done
