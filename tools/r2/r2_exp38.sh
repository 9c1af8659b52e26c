#!/bin/bash
# Round-2 GPU call 38 (one B200): the final tree after the loss-kernel change -- full GPU suite, smoke, N = 1 bench line.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
O=gpurun_out
( time timeout 600 python -m pytest tests -m gpu -q --durations=8 ) > $O/r2e38_gpu_tests.log 2>&1
echo "pytest rc=$?"; tail -14 $O/r2e38_gpu_tests.log
timeout 200 python __graft_entry__.py --smoke > $O/r2e38_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/r2e38_smoke.log
timeout 400 python bench.py > $O/r2_final2_bench_n1.json 2> $O/r2_final2_bench_n1.log; echo "bench rc=$?"
python -c "import json; d=json.load(open('$O/r2_final2_bench_n1.json')); print(d['value']/1e9, d['roofline']['frac'], d['e2e']['value']/1e9, d['e2e']['ms_all_steps'], d['breakdown_ms_per_step'], d['test_rmse'])"
