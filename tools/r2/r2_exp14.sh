#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
for g in 444 888; do
  SWEEP_ITERS=256 SWEEP_ROUND=8 python tools/dsgd_stability_map.py nfblock8 $g none:0 2>/dev/null | sed "s/^/round 8 /"
  SWEEP_ITERS=256 SWEEP_ROUND=64 python tools/dsgd_stability_map.py nfblock8 $g none:0 2>/dev/null | sed "s/^/round 64 /"
done
