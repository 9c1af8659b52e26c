#!/bin/bash
# Round-2 GPU call 36 (one B200): loss kernel with U ratings in flight per lane group (contiguous runs): parity tests and A/B of U = 1 / 2 / 4.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "loss or residual or metric or config1 or reproducible" 2>&1 | tail -3
for u in 1 2 4; do
  CU2B_LOSS_UNROLL=$u python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-variants 2>/dev/null | python -c "import sys,json; d=json.load(sys.stdin); print('U=$u', d['value']/1e9, d['breakdown_ms_per_step'], d['test_rmse'][-1], d['e2e']['ms_per_step'])"
done
