#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
for r in 32 64 96 128 256; do
  SWEEP_ROUND=$r python tools/dsgd_stability_map.py nfblock8 1184 none:0 2>/dev/null | sed "s/^/round $r /"
done
