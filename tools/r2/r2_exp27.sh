#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
run() { # n tag env...
  local n=$1 tag=$2; shift 2
  env "$@" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29513 \
      bench.py --gpus $n --steps 4 --warmup 4 > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.log
  echo "$tag rc=$? $(wc -c < gpurun_out/bench_${tag}.json) bytes"
}
run 8 r2d_n8_r1_c4 CU2B_DSGD_RANGES=1 CU2B_DSGD_CLAIM=4
run 8 r2d_n8_r64_c1 CU2B_DSGD_RANGES=64 CU2B_DSGD_CLAIM=1
run 8 r2d_n8_r16_c2 CU2B_DSGD_RANGES=16 CU2B_DSGD_CLAIM=2
