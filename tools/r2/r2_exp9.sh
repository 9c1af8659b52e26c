#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 600 bash tools/run_bench_n.sh 2 r2_n2_fused --steps 3 --warmup 2
CU2B_DSGD_ROUND=128 timeout 600 bash tools/run_bench_n.sh 2 r2_n2_fused_round128 --steps 3 --warmup 2
tail -3 gpurun_out/bench_r2_n2_fused_round128.log
