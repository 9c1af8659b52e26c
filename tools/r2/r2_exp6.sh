#!/bin/bash
# Round-2 GPU call 6 (one B200): the single-GPU bench line with the new roofline / e2e, and a host-side trace of
# the single-call path (where do create / destroy spend their time?).
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
python bench.py --steps 4 --warmup 3 > gpurun_out/bench_r2_n1.json 2> gpurun_out/bench_r2_n1.log
echo "bench rc=$?"; tail -3 gpurun_out/bench_r2_n1.log
python tools/trace_e2e.py > gpurun_out/r2_trace_e2e.txt 2>&1
echo "trace rc=$?"; tail -40 gpurun_out/r2_trace_e2e.txt
