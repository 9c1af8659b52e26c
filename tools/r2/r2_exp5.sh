#!/bin/bash
# Round-2 GPU call 5 (two B200): DSGD at N = 2 with the fused sub-epoch kernel, and unfused for A/B.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2e5_topo.txt 2>&1
timeout 600 bash tools/run_bench_n.sh 2 r2_n2_fused --steps 3 --warmup 2
tail -5 gpurun_out/bench_r2_n2_fused.log
CU2B_DSGD_FUSED=0 timeout 600 bash tools/run_bench_n.sh 2 r2_n2_unfused --steps 3 --warmup 2
tail -3 gpurun_out/bench_r2_n2_unfused.log
