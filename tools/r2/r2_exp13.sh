#!/bin/bash
# Round-2 GPU call 13 (eight B200): round length at N = 8 / 4 with the right-sized grid and single-tile claims.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
run() { # n tag env...
  local n=$1 tag=$2; shift 2
  env "$@" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29513 \
      bench.py --gpus $n --steps 4 --warmup 4 > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.log
  echo "$tag rc=$? $(wc -c < gpurun_out/bench_${tag}.json) bytes"
}
run 8 r2b_n8_round64 A=1
run 8 r2b_n8_round128 CU2B_DSGD_ROUND=128
run 8 r2b_n8_round256 CU2B_DSGD_ROUND=256
run 8 r2b_n8_round128_rows05 CU2B_DSGD_ROUND=128 CU2B_DSGD_THIN=0.5
run 4 r2b_n4_round128 CU2B_DSGD_ROUND=128
