#!/bin/bash
# Round-2 GPU call 10 (eight B200): DSGD at N = 8 (default, unfused, row thinning, longer rounds) and N = 4.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
run() { # n tag env...
  local n=$1 tag=$2; shift 2
  env "$@" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29513 \
      bench.py --gpus $n --steps 4 --warmup 4 > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.log
  echo "$tag rc=$? $(wc -c < gpurun_out/bench_${tag}.json) bytes"
}
run 8 r2_n8_default A=1
run 4 r2_n4_default A=1
run 8 r2_n8_rows05 CU2B_DSGD_THIN=0.5
run 8 r2_n8_unfused CU2B_DSGD_FUSED=0
run 8 r2_n8_round32 CU2B_DSGD_ROUND=32
run 8 r2_n8_rows025 CU2B_DSGD_THIN=0.25
tail -3 gpurun_out/bench_r2_n8_default.log
