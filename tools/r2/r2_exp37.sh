#!/bin/bash
# Round-2 GPU call 37 (one B200): loss kernel, occupancy against ratings in flight (k = 128 layout only).
cd "${GRAFT_REPO_ROOT:-.}"
for v in default u1c5 u2c4 u1c3; do
  CU2B_LOSS_VARIANT=$v python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-variants 2>/dev/null | python -c "import sys,json; d=json.load(sys.stdin); print('$v', d['value']/1e9, d['breakdown_ms_per_step'], d['test_rmse'][-1])"
done
